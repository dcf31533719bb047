// exp(x) for x <= 0 (the RBF kernel's exp(-r), r >= 0), hand-rolled for the FP64 pipe: 12 FP64 instructions where
// libdevice's exp() issues about 23, which is what bounds the covariance kernels on B200 (DFMA and DMMA share one pipe,
// 64 lanes/clk/SM).  No overflow, NaN or denormal branches are needed because x <= 0 and the result lies in [0, 1].
//
//   k = rint(x * 64/ln2),  k = 64 m + j (0 <= j < 64),  t = x - k ln2/64 (Cody-Waite, |t| <= ln2/128 = 0.0054)
//   exp(x) = 2^m * T_j * e^t,   T_j = 2^(j/64) as a double-double (hi, lo) from a 1 KB table,
//   e^t - 1 = t + t^2 (1/2 + t (1/6 + t (1/24 + t (1/120 + t/720))))        truncation t^7/5040 < 3e-20
//   result = 2^m * (T_hi + (T_lo + T_hi (e^t - 1)))                          one final rounding: < 0.51 ulp
// Results below 2^-1021 (x < -707.7) are flushed to zero (numpy returns denormals there; the difference is < 2.3e-308);
// the argument is screened on its high word first, so that huge |x| (tiny lengthscales) cannot wrap the integer k.
//
// The same source compiles as plain C (fma() from libm) so that a CPU test can check the polynomial against mpmath
// bit for bit: every operation below is a correctly rounded IEEE operation, written out explicitly (no contraction).
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define GPM_EXP_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define GPM_EXP_HD static inline
#endif

typedef struct { double hi, lo; } gpm_exp_pair;

#if defined(__CUDACC__)
static __device__ const gpm_exp_pair gpm_exp_tab_dev[64] = {
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.02c9a3e778061p+0, -0x1.19083535b085dp-56},
    {0x1.059b0d3158574p+0, 0x1.d73e2a475b465p-55},
    {0x1.0874518759bc8p+0, 0x1.186be4bb284ffp-57},
    {0x1.0b5586cf9890fp+0, 0x1.8a62e4adc610bp-54},
    {0x1.0e3ec32d3d1a2p+0, 0x1.03a1727c57b53p-59},
    {0x1.11301d0125b51p+0, -0x1.6c51039449b3ap-54},
    {0x1.1429aaea92de0p+0, -0x1.32fbf9af1369ep-54},
    {0x1.172b83c7d517bp+0, -0x1.19041b9d78a76p-55},
    {0x1.1a35beb6fcb75p+0, 0x1.e5b4c7b4968e4p-55},
    {0x1.1d4873168b9aap+0, 0x1.e016e00a2643cp-54},
    {0x1.2063b88628cd6p+0, 0x1.dc775814a8495p-55},
    {0x1.2387a6e756238p+0, 0x1.9b07eb6c70573p-54},
    {0x1.26b4565e27cddp+0, 0x1.2bd339940e9d9p-55},
    {0x1.29e9df51fdee1p+0, 0x1.612e8afad1255p-55},
    {0x1.2d285a6e4030bp+0, 0x1.0024754db41d5p-54},
    {0x1.306fe0a31b715p+0, 0x1.6f46ad23182e4p-55},
    {0x1.33c08b26416ffp+0, 0x1.32721843659a6p-54},
    {0x1.371a7373aa9cbp+0, -0x1.63aeabf42eae2p-54},
    {0x1.3a7db34e59ff7p+0, -0x1.5e436d661f5e3p-56},
    {0x1.3dea64c123422p+0, 0x1.ada0911f09ebcp-55},
    {0x1.4160a21f72e2ap+0, -0x1.ef3691c309278p-58},
    {0x1.44e086061892dp+0, 0x1.89b7a04ef80d0p-59},
    {0x1.486a2b5c13cd0p+0, 0x1.3c1a3b69062f0p-56},
    {0x1.4bfdad5362a27p+0, 0x1.d4397afec42e2p-56},
    {0x1.4f9b2769d2ca7p+0, -0x1.4b309d25957e3p-54},
    {0x1.5342b569d4f82p+0, -0x1.07abe1db13cadp-55},
    {0x1.56f4736b527dap+0, 0x1.9bb2c011d93adp-54},
    {0x1.5ab07dd485429p+0, 0x1.6324c054647adp-54},
    {0x1.5e76f15ad2148p+0, 0x1.ba6f93080e65ep-54},
    {0x1.6247eb03a5585p+0, -0x1.383c17e40b497p-54},
    {0x1.6623882552225p+0, -0x1.bb60987591c34p-54},
    {0x1.6a09e667f3bcdp+0, -0x1.bdd3413b26456p-54},
    {0x1.6dfb23c651a2fp+0, -0x1.bbe3a683c88abp-57},
    {0x1.71f75e8ec5f74p+0, -0x1.16e4786887a99p-55},
    {0x1.75feb564267c9p+0, -0x1.0245957316dd3p-54},
    {0x1.7a11473eb0187p+0, -0x1.41577ee04992fp-55},
    {0x1.7e2f336cf4e62p+0, 0x1.05d02ba15797ep-56},
    {0x1.82589994cce13p+0, -0x1.d4c1dd41532d8p-54},
    {0x1.868d99b4492edp+0, -0x1.fc6f89bd4f6bap-54},
    {0x1.8ace5422aa0dbp+0, 0x1.6e9f156864b27p-54},
    {0x1.8f1ae99157736p+0, 0x1.5cc13a2e3976cp-55},
    {0x1.93737b0cdc5e5p+0, -0x1.75fc781b57ebcp-57},
    {0x1.97d829fde4e50p+0, -0x1.d185b7c1b85d1p-54},
    {0x1.9c49182a3f090p+0, 0x1.c7c46b071f2bep-56},
    {0x1.a0c667b5de565p+0, -0x1.359495d1cd533p-54},
    {0x1.a5503b23e255dp+0, -0x1.d2f6edb8d41e1p-54},
    {0x1.a9e6b5579fdbfp+0, 0x1.0fac90ef7fd31p-54},
    {0x1.ae89f995ad3adp+0, 0x1.7a1cd345dcc81p-54},
    {0x1.b33a2b84f15fbp+0, -0x1.2805e3084d708p-57},
    {0x1.b7f76f2fb5e47p+0, -0x1.5584f7e54ac3bp-56},
    {0x1.bcc1e904bc1d2p+0, 0x1.23dd07a2d9e84p-55},
    {0x1.c199bdd85529cp+0, 0x1.11065895048ddp-55},
    {0x1.c67f12e57d14bp+0, 0x1.2884dff483cadp-54},
    {0x1.cb720dcef9069p+0, 0x1.503cbd1e949dbp-56},
    {0x1.d072d4a07897cp+0, -0x1.cbc3743797a9cp-54},
    {0x1.d5818dcfba487p+0, 0x1.2ed02d75b3707p-55},
    {0x1.da9e603db3285p+0, 0x1.c2300696db532p-54},
    {0x1.dfc97337b9b5fp+0, -0x1.1a5cd4f184b5cp-54},
    {0x1.e502ee78b3ff6p+0, 0x1.39e8980a9cc8fp-55},
    {0x1.ea4afa2a490dap+0, -0x1.e9c23179c2893p-54},
    {0x1.efa1bee615a27p+0, 0x1.dc7f486a4b6b0p-54},
    {0x1.f50765b6e4540p+0, 0x1.9d3e12dd8a18bp-54},
    {0x1.fa7c1819e90d8p+0, 0x1.74853f3a5931ep-55},
};
#endif
static const gpm_exp_pair gpm_exp_tab_host[64] = {
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.02c9a3e778061p+0, -0x1.19083535b085dp-56},
    {0x1.059b0d3158574p+0, 0x1.d73e2a475b465p-55},
    {0x1.0874518759bc8p+0, 0x1.186be4bb284ffp-57},
    {0x1.0b5586cf9890fp+0, 0x1.8a62e4adc610bp-54},
    {0x1.0e3ec32d3d1a2p+0, 0x1.03a1727c57b53p-59},
    {0x1.11301d0125b51p+0, -0x1.6c51039449b3ap-54},
    {0x1.1429aaea92de0p+0, -0x1.32fbf9af1369ep-54},
    {0x1.172b83c7d517bp+0, -0x1.19041b9d78a76p-55},
    {0x1.1a35beb6fcb75p+0, 0x1.e5b4c7b4968e4p-55},
    {0x1.1d4873168b9aap+0, 0x1.e016e00a2643cp-54},
    {0x1.2063b88628cd6p+0, 0x1.dc775814a8495p-55},
    {0x1.2387a6e756238p+0, 0x1.9b07eb6c70573p-54},
    {0x1.26b4565e27cddp+0, 0x1.2bd339940e9d9p-55},
    {0x1.29e9df51fdee1p+0, 0x1.612e8afad1255p-55},
    {0x1.2d285a6e4030bp+0, 0x1.0024754db41d5p-54},
    {0x1.306fe0a31b715p+0, 0x1.6f46ad23182e4p-55},
    {0x1.33c08b26416ffp+0, 0x1.32721843659a6p-54},
    {0x1.371a7373aa9cbp+0, -0x1.63aeabf42eae2p-54},
    {0x1.3a7db34e59ff7p+0, -0x1.5e436d661f5e3p-56},
    {0x1.3dea64c123422p+0, 0x1.ada0911f09ebcp-55},
    {0x1.4160a21f72e2ap+0, -0x1.ef3691c309278p-58},
    {0x1.44e086061892dp+0, 0x1.89b7a04ef80d0p-59},
    {0x1.486a2b5c13cd0p+0, 0x1.3c1a3b69062f0p-56},
    {0x1.4bfdad5362a27p+0, 0x1.d4397afec42e2p-56},
    {0x1.4f9b2769d2ca7p+0, -0x1.4b309d25957e3p-54},
    {0x1.5342b569d4f82p+0, -0x1.07abe1db13cadp-55},
    {0x1.56f4736b527dap+0, 0x1.9bb2c011d93adp-54},
    {0x1.5ab07dd485429p+0, 0x1.6324c054647adp-54},
    {0x1.5e76f15ad2148p+0, 0x1.ba6f93080e65ep-54},
    {0x1.6247eb03a5585p+0, -0x1.383c17e40b497p-54},
    {0x1.6623882552225p+0, -0x1.bb60987591c34p-54},
    {0x1.6a09e667f3bcdp+0, -0x1.bdd3413b26456p-54},
    {0x1.6dfb23c651a2fp+0, -0x1.bbe3a683c88abp-57},
    {0x1.71f75e8ec5f74p+0, -0x1.16e4786887a99p-55},
    {0x1.75feb564267c9p+0, -0x1.0245957316dd3p-54},
    {0x1.7a11473eb0187p+0, -0x1.41577ee04992fp-55},
    {0x1.7e2f336cf4e62p+0, 0x1.05d02ba15797ep-56},
    {0x1.82589994cce13p+0, -0x1.d4c1dd41532d8p-54},
    {0x1.868d99b4492edp+0, -0x1.fc6f89bd4f6bap-54},
    {0x1.8ace5422aa0dbp+0, 0x1.6e9f156864b27p-54},
    {0x1.8f1ae99157736p+0, 0x1.5cc13a2e3976cp-55},
    {0x1.93737b0cdc5e5p+0, -0x1.75fc781b57ebcp-57},
    {0x1.97d829fde4e50p+0, -0x1.d185b7c1b85d1p-54},
    {0x1.9c49182a3f090p+0, 0x1.c7c46b071f2bep-56},
    {0x1.a0c667b5de565p+0, -0x1.359495d1cd533p-54},
    {0x1.a5503b23e255dp+0, -0x1.d2f6edb8d41e1p-54},
    {0x1.a9e6b5579fdbfp+0, 0x1.0fac90ef7fd31p-54},
    {0x1.ae89f995ad3adp+0, 0x1.7a1cd345dcc81p-54},
    {0x1.b33a2b84f15fbp+0, -0x1.2805e3084d708p-57},
    {0x1.b7f76f2fb5e47p+0, -0x1.5584f7e54ac3bp-56},
    {0x1.bcc1e904bc1d2p+0, 0x1.23dd07a2d9e84p-55},
    {0x1.c199bdd85529cp+0, 0x1.11065895048ddp-55},
    {0x1.c67f12e57d14bp+0, 0x1.2884dff483cadp-54},
    {0x1.cb720dcef9069p+0, 0x1.503cbd1e949dbp-56},
    {0x1.d072d4a07897cp+0, -0x1.cbc3743797a9cp-54},
    {0x1.d5818dcfba487p+0, 0x1.2ed02d75b3707p-55},
    {0x1.da9e603db3285p+0, 0x1.c2300696db532p-54},
    {0x1.dfc97337b9b5fp+0, -0x1.1a5cd4f184b5cp-54},
    {0x1.e502ee78b3ff6p+0, 0x1.39e8980a9cc8fp-55},
    {0x1.ea4afa2a490dap+0, -0x1.e9c23179c2893p-54},
    {0x1.efa1bee615a27p+0, 0x1.dc7f486a4b6b0p-54},
    {0x1.f50765b6e4540p+0, 0x1.9d3e12dd8a18bp-54},
    {0x1.fa7c1819e90d8p+0, 0x1.74853f3a5931ep-55},
};

// x <= 0 (values above 0 up to a few ulps of rounding noise are fine); returns exp(x), correctly signed zero for -inf
GPM_EXP_HD double gpm_exp_neg(double x) {
  const double INV = 0x1.71547652b82fep+6;             // 64 / ln2
  const double C1 = 0x1.62e42fef80000p-7;             // ln2 / 64, high part (19 trailing zero bits: k * C1 is exact for |k| < 2^17)
  const double C2 = 0x1.1cf79abc9e3b4p-42;            // ln2 / 64 - C1
  const double MAGIC = 0x1.8p+52;                        // 1.5 * 2^52: adding it rounds to an integer held in the low word
#if defined(__CUDA_ARCH__)
  const bool under = (unsigned)__double2hiint(x) >= 0xC0862400u;       // x <= -708.5 (or -inf): underflow (branch-free)
  const double km = __fma_rn(x, INV, MAGIC);
  const double kf = __dadd_rn(km, -MAGIC);
  const int k = __double2loint(km);
  double t = __fma_rn(-kf, C1, x);
  t = __fma_rn(-kf, C2, t);
  double q = __fma_rn(0x1.6c16c16c16c17p-10, t, 0x1.1111111111111p-7);     // 1/720, 1/120
  q = __fma_rn(q, t, 0x1.5555555555555p-5);                                 // 1/24
  q = __fma_rn(q, t, 0x1.5555555555555p-3);                                 // 1/6
  q = __fma_rn(q, t, 0.5);
  const double t2 = __dmul_rn(t, t);
  const double p = __fma_rn(t2, q, t);                                      // e^t - 1
  const gpm_exp_pair T = gpm_exp_tab_dev[k & 63];
  double r = __fma_rn(T.hi, p, T.lo);
  r = __dadd_rn(T.hi, r);
  const int m = k >> 6;
  const double scaled = __hiloint2double(__double2hiint(r) + (m << 20), __double2loint(r));
  return (under || m < -1021) ? 0.0 : scaled;
#else
  {
    int64_t xb;
    memcpy(&xb, &x, 8);
    if ((uint32_t)((uint64_t)xb >> 32) >= 0xC0862400u) return 0.0;
  }
  const double km = fma(x, INV, MAGIC);
  const double kf = km - MAGIC;
  int64_t bits;
  memcpy(&bits, &km, 8);
  const int k = (int)(int32_t)(uint32_t)(bits & 0xffffffffu);
  double t = fma(-kf, C1, x);
  t = fma(-kf, C2, t);
  double q = fma(0x1.6c16c16c16c17p-10, t, 0x1.1111111111111p-7);
  q = fma(q, t, 0x1.5555555555555p-5);
  q = fma(q, t, 0x1.5555555555555p-3);
  q = fma(q, t, 0.5);
  const double t2 = t * t;
  const double p = fma(t2, q, t);
  const gpm_exp_pair T = gpm_exp_tab_host[k & 63];
  double r = fma(T.hi, p, T.lo);
  r = T.hi + r;
  const int m = k >> 6;
  if (m < -1021) return 0.0;
  int64_t rb;
  memcpy(&rb, &r, 8);
  rb += (int64_t)m << 52;
  memcpy(&r, &rb, 8);
  return r;
#endif
}

// exp(-d2 / 2) for d2 >= 0: the form the RBF kernel needs.  Same algorithm with the factor -1/2 folded into the
// constants (scaling by a power of two is exact, so the reduced argument is the same number): u = -2 t,
//   e^t - 1 = u * (-1/2 + u (1/8 + u (-1/48 + u (1/384 + u (-1/3840 + u/46080)))))
// One FP64 instruction fewer than gpm_exp_neg(-0.5 * d2): 11 in all.
GPM_EXP_HD double gpm_exp_neg_half(double d2) {
  const double NINV2 = -0x1.71547652b82fep+5;          // -32 / ln2
  const double D1 = 0x1.62e42fef80000p-6;              // 2 * C1
  const double D2 = 0x1.1cf79abc9e3b4p-41;             // 2 * C2
  const double MAGIC = 0x1.8p+52;
  const double Q6 = 0x1.6c16c16c16c17p-16, Q5 = -0x1.1111111111111p-12, Q4 = 0x1.5555555555555p-9,
               Q3 = -0x1.5555555555555p-6, Q2 = 0.125;
#if defined(__CUDA_ARCH__)
  const bool under = (unsigned)__double2hiint(d2) >= 0x40962400u;       // d2 >= 1417 (or +inf): underflow
  const double km = __fma_rn(d2, NINV2, MAGIC);
  const double kf = __dadd_rn(km, -MAGIC);
  const int k = __double2loint(km);
  double u = __fma_rn(kf, D1, d2);
  u = __fma_rn(kf, D2, u);
  double q = __fma_rn(Q6, u, Q5);
  q = __fma_rn(q, u, Q4);
  q = __fma_rn(q, u, Q3);
  q = __fma_rn(q, u, Q2);
  q = __fma_rn(q, u, -0.5);
  const double p = __dmul_rn(u, q);                                         // e^t - 1
  const gpm_exp_pair T = gpm_exp_tab_dev[k & 63];
  double r = __fma_rn(T.hi, p, T.lo);
  r = __dadd_rn(T.hi, r);
  const int m = k >> 6;
  const double scaled = __hiloint2double(__double2hiint(r) + (m << 20), __double2loint(r));
  return (under || m < -1021) ? 0.0 : scaled;
#else
  int64_t xb;
  memcpy(&xb, &d2, 8);
  const int under = (uint32_t)((uint64_t)xb >> 32) >= 0x40962400u;
  const double km = fma(d2, NINV2, MAGIC);
  const double kf = km - MAGIC;
  int64_t bits;
  memcpy(&bits, &km, 8);
  const int k = (int)(int32_t)(uint32_t)(bits & 0xffffffffu);
  double u = fma(kf, D1, d2);
  u = fma(kf, D2, u);
  double q = fma(Q6, u, Q5);
  q = fma(q, u, Q4);
  q = fma(q, u, Q3);
  q = fma(q, u, Q2);
  q = fma(q, u, -0.5);
  const double p = u * q;
  const gpm_exp_pair T = gpm_exp_tab_host[k & 63];
  double r = fma(T.hi, p, T.lo);
  r = T.hi + r;
  const int m = k >> 6;
  if (under || m < -1021) return 0.0;
  int64_t rb;
  memcpy(&rb, &r, 8);
  rb += (int64_t)m << 52;
  memcpy(&r, &rb, 8);
  return r;
#endif
}
