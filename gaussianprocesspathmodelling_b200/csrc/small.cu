// Batched fits of SHORT paths (N <= 112, the reference's regime: GPmap.py:189 resamples every trajectory to 33
// points): ONE CTA PER PATH, the whole fit in shared memory.
//
//   covariance (lower triangle, packed)  ->  Cholesky (left-looking, one thread per row)  ->  forward and backward
//   substitution  ->  alpha, log marginal likelihood
//
// Nothing but X, Y (in) and alpha, lml, info (out) touches HBM: 8 N (D + 2R) + 8R bytes per path, where the tiled
// pipeline of batched.cu writes and re-reads a padded 128 x 128 block three times.  A path of N = 33 needs 4.5 KB of
// shared memory, so 16 CTAs share an SM and hide each other's barrier and pivot latencies; N = 112 needs 52 KB.
// Measured (B200): N=33 41.9 M fits/s (tiled pipeline 5.0 M), N=64 16.8 M, N=96 6.0 M (4.0 M).
// FP64 CUDA cores only: a 33 x 33 factorisation is 12 kflop, less than the 561 exponentials of its covariance.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace gpm {

constexpr int SMALL_MAX_N = 112;     // measured crossover with the tiled pipeline (8192 paths: N=96 6.0 vs 4.0 M fits/s, N=112 3.8 vs 3.8, N=128 2.8 vs 3.9)

__host__ __device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

// TWO: a path a few rows longer than a multiple of 32 (the reference's N = 33 is the case that matters) does not get a
// nearly empty extra warp -- which would cost a full warp's issue slots in every column, N = 33 ran at 0.56 of the N = 32
// rate -- but T = 32 floor(N / 32) threads, and thread t < N - T also owns row T + t as its SECOND row.  A row t is idle
// in the columns j > t of the factorisation (and after step t of the forward substitution), which is exactly when its
// thread works on row T + t in the same pass (same trip count: the dot products of column j all have length j), so only
// the first N - T columns / steps need a second pass (for the threads whose two rows are both live).
template <int D, int RR, bool TWO>
__global__ void __launch_bounds__(128)
fit_small_kernel(const double* __restrict__ Xb, const double* __restrict__ Yb, int N, int R, Theta th,
                 const double* __restrict__ theta_dev, int theta_stride, double* __restrict__ alphab,
                 double* __restrict__ lmlb, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* Kp = sm;                          // packed lower triangle, row i at i(i+1)/2
  double* xs = Kp + tri(N);                 // [N][3] coordinates / lengthscale
  double* dinv = xs + 3 * N;                // [N] 1 / L_ii
  double* piv = dinv + N;                   // [N] pivots d_j (broadcast from the diagonal row to the rows below)
  double* vs = piv + N;                     // [N][RR] the solved entry of the current substitution step
  double* red = vs + N * RR;                // [4][RR + 1] cross-warp reduction
  const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
  // a one-warp CTA (N <= 44) orders its shared-memory traffic with __syncwarp(): no trip to the barrier unit
  const bool one_warp = nthr == 32;
#define SM_SYNC() do { if (one_warp) __syncwarp(); else __syncthreads(); } while (0)
  const long long b = blockIdx.x;
  if (theta_dev) {                          // per-path hyper-parameters
    const double* t = theta_dev + b * theta_stride;
#pragma unroll
    for (int d = 0; d < D; d++) th.l[d] = t[d];
    th.sf2 = t[D];
    th.sn2 = t[D + 1];
  }
  const double* X = Xb + b * N * D;
  const double* Y = Yb + b * N * R;
  for (int e = tid; e < N * D; e += nthr) {
    const int i = e / D, d = e % D;
    xs[i * 3 + d] = X[e] / (d == 0 ? th.l[0] : (d == 1 ? th.l[1] : th.l[2]));
  }
  const int i = tid;                        // the row this thread owns
  const int T = nthr;                       // TWO: rows T .. N-1 are the second rows of threads 0 .. e2-1
  const int e2 = TWO ? N - T : 0;
  const bool has2 = TWO && tid < e2;
  const int i2 = T + tid;
  double y[RR], y0[RR], y2[RR], y02[RR];     // y2 / y02: the second row (TWO only; dead otherwise)
#pragma unroll
  for (int r = 0; r < RR; r++) y0[r] = y[r] = (i < N && r < R) ? Y[i * R + r] : 0.0;
  if constexpr (TWO) {
#pragma unroll
    for (int r = 0; r < RR; r++) y02[r] = y2[r] = (has2 && r < R) ? Y[i2 * R + r] : 0.0;
  }
  SM_SYNC();

  // ---- covariance: the N(N+1)/2 entries of the lower triangle dealt evenly to the threads (same arithmetic as
  //      cov_kernel) ----
  for (int e = tid; e < tri(N); e += nthr) {
    int r = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while (tri(r + 1) <= e) r++;
    while (tri(r) > e) r--;
    const int c = e - tri(r);
    const double a[3] = {xs[r * 3], xs[r * 3 + 1], D == 3 ? xs[r * 3 + 2] : 0.0};
    const double bq[3] = {xs[c * 3], xs[c * 3 + 1], D == 3 ? xs[c * 3 + 2] : 0.0};
    double v = rbf<D>(a, bq, th.sf2);
    if (c == r) v += th.sn2;
    Kp[e] = v;
  }
  SM_SYNC();

  // ---- Cholesky, left-looking: column j of L from the finished columns 0..j-1 (two barriers per column) ----
  int bad = 0;
  // K[row][j] - sum_{k < j} L[row][k] L[j][k]
  auto col_entry = [&](int row, int j) -> double {
    const double* ri = Kp + tri(row);
    const double* rj = Kp + tri(j);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = 0;
    for (; k + 7 < j; k += 8) {               // sixteen independent shared loads in flight per step
      const double a0 = ri[k], a1 = ri[k + 1], a2 = ri[k + 2], a3 = ri[k + 3];
      const double a4 = ri[k + 4], a5 = ri[k + 5], a6 = ri[k + 6], a7 = ri[k + 7];
      const double b0 = rj[k], b1 = rj[k + 1], b2 = rj[k + 2], b3 = rj[k + 3];
      const double b4 = rj[k + 4], b5 = rj[k + 5], b6 = rj[k + 6], b7 = rj[k + 7];
      s0 = fma(a0, b0, s0); s1 = fma(a1, b1, s1); s2 = fma(a2, b2, s2); s3 = fma(a3, b3, s3);
      s0 = fma(a4, b4, s0); s1 = fma(a5, b5, s1); s2 = fma(a6, b6, s2); s3 = fma(a7, b7, s3);
    }
    for (; k < j; k++) s0 = fma(ri[k], rj[k], s0);
    return ri[j] - ((s0 + s1) + (s2 + s3));
  };
  for (int j = 0; j < N; j++) {
    // the row of this pass: the thread's own row while it is live (i >= j), its second row afterwards (TWO)
    const int row = (i >= j) ? i : (has2 ? i2 : N);
    const bool act = row >= j && row < N;
    const bool both = TWO && j < e2 && has2 && i >= j;   // first e2 columns: both rows of the thread are live (second pass)
    double s = 0.0, sb = 0.0;
    if (act) {
      s = col_entry(row, j);
      if (row == j) piv[j] = s;              // the pivot d_j; the rows below keep their unnormalised entry in a register
    }
    if (TWO && j < e2) {                      // block-uniform
      if (both) sb = col_entry(i2, j);       // i2 >= 32 > j: never the diagonal
    }
    SM_SYNC();
    if (act || both) {
      double d = piv[j];
      if (!(d > 0.0) || !(d < 1.0e300)) { if (!bad) bad = j + 1; d = 1.0; }
      const double rinv = rsqrt(d);          // as potf2: 1 ulp, a fifth of the latency of sqrt + divide
      if (act) {
        if (row == j) { dinv[j] = rinv; Kp[tri(j) + j] = d * rinv; }
        else Kp[tri(row) + j] = s * rinv;
      }
      if (both) Kp[tri(i2) + j] = sb * rinv;
    }
    SM_SYNC();
  }
  // the thread that owns row N-1 takes part in every column (with one row or the other), so it saw the first bad pivot
  if (TWO ? (has2 && i2 == N - 1) : (i == N - 1)) info[b] = bad;

  // ---- forward substitution L z = y (column-oriented: one barrier per step) ----
  //      TWO: y is the thread's own row until that row is solved (step i), then its second row (swapped in: y2 keeps z_i);
  //      while both are live (steps j < i < e2) the second row is updated in y2
  for (int j = 0; j < N; j++) {
    if (TWO ? (j < T ? i == j : (has2 && i2 == j)) : (i == j)) {
#pragma unroll
      for (int r = 0; r < RR; r++) { y[r] *= dinv[j]; vs[j * RR + r] = y[r]; }
      if constexpr (TWO) {
        if (j < T && has2) {
#pragma unroll
          for (int r = 0; r < RR; r++) { const double t = y[r]; y[r] = y2[r]; y2[r] = t; }
        }
      }
    }
    SM_SYNC();
    const int row = (i > j) ? i : (has2 ? i2 : 0);
    if (row > j && row < N) {
      const double l = Kp[tri(row) + j];
#pragma unroll
      for (int r = 0; r < RR; r++) y[r] = fma(-l, vs[j * RR + r], y[r]);
    }
    if (TWO && j + 1 < e2) {                  // block-uniform
      if (has2 && i > j) {
        const double l = Kp[tri(i2) + j];
#pragma unroll
        for (int r = 0; r < RR; r++) y2[r] = fma(-l, vs[j * RR + r], y2[r]);
      }
    }
  }
  SM_SYNC();
  if constexpr (TWO) {                        // back to y = own row (z_i), y2 = second row (z_i2)
    if (has2) {
#pragma unroll
      for (int r = 0; r < RR; r++) { const double t = y[r]; y[r] = y2[r]; y2[r] = t; }
    }
  }
  // ---- backward substitution L^T alpha = z ----
  for (int j = N - 1; j >= 0; j--) {
    if (TWO && j >= T) {
      if (has2 && i2 == j) {
#pragma unroll
        for (int r = 0; r < RR; r++) { y2[r] *= dinv[j]; vs[j * RR + r] = y2[r]; }
      }
    } else if (i == j) {
#pragma unroll
      for (int r = 0; r < RR; r++) { y[r] *= dinv[j]; vs[j * RR + r] = y[r]; }
    }
    SM_SYNC();
    if (i < j) {
      const double l = Kp[tri(j) + i];
#pragma unroll
      for (int r = 0; r < RR; r++) y[r] = fma(-l, vs[j * RR + r], y[r]);
    }
    if (TWO && j > T) {                       // block-uniform: second rows below row j
      if (has2 && i2 < j) {
        const double l = Kp[tri(j) + i2];
#pragma unroll
        for (int r = 0; r < RR; r++) y2[r] = fma(-l, vs[j * RR + r], y2[r]);
      }
    }
  }
  if (i < N) {
    double* al = alphab + b * N * R + i * R;
#pragma unroll
    for (int r = 0; r < RR; r++) if (r < R) al[r] = y[r];
  }
  if (has2) {
    double* al = alphab + b * N * R + i2 * R;
#pragma unroll
    for (int r = 0; r < RR; r++) if (r < R) al[r] = y2[r];
  }
  if (lmlb == nullptr) return;

  // ---- lml[r] = -1/2 y^T alpha - sum_i log L_ii - N/2 log(2 pi); reductions in a fixed order ----
  double part[RR + 1];
#pragma unroll
  for (int r = 0; r < RR; r++) part[r] = (i < N) ? y0[r] * y[r] : 0.0;
  part[RR] = (i < N) ? log(Kp[tri(i) + i]) : 0.0;
  if constexpr (TWO) {
    if (has2) {
#pragma unroll
      for (int r = 0; r < RR; r++) part[r] = fma(y02[r], y2[r], part[r]);
      part[RR] += log(Kp[tri(i2) + i2]);
    }
  }
#pragma unroll
  for (int r = 0; r <= RR; r++) {
    const double sv = warp_sum(part[r]);
    if (lane == 0) red[warp * (RR + 1) + r] = sv;
  }
  SM_SYNC();
  if (tid < R) {
    double q = 0.0, ld = 0.0;
    for (int w = 0; w < nwarp; w++) { q += red[w * (RR + 1) + tid]; ld += red[w * (RR + 1) + RR]; }
    lmlb[b * R + tid] = -0.5 * q - ld - 0.5 * (double)N * 1.8378770664093453;   // log(2 pi)
  }
}

#undef SM_SYNC

bool fit_small_supported(long long N) { return N >= 1 && N <= SMALL_MAX_N; }

static size_t small_smem(int N, int RR) {
  return (size_t)(tri(N) + 3 * N + 2 * N + N * RR + 4 * (RR + 1)) * sizeof(double);
}

template <int D, int RR>
static int launch_small_dr(const double* Xb, const double* Yb, int B, int N, int R, const Theta& th,
                           const double* theta_dev, int theta_stride, double* alpha, double* lml, int* info,
                           cudaStream_t stream, int two_max) {
  const size_t smem = small_smem(N, RR);
  // 32 < N <= 32 + small_two_max: the rows beyond 32 ride as second rows of the first threads instead of opening a second warp.
  // Measured (B200, 16384 paths, M fits/s, second rows vs one thread per row): N=33 65.8 vs 49.5, 36 55.5 vs 46.8, 40 47.0 vs
  // 41.1, 44 38.1 vs 35.6, 48 30.4 vs 30.8, 52 24.4 vs 26.3; beyond one warp it loses (the second rows all sit in warp 0 and
  // the other warps wait for it at the barriers: N=68 11.6 vs 12.3, 72 10.1 vs 11.1, 100 3.8 vs 4.1), so T = 32 only.
  const bool two = N > 32 && N < 64 && N - 32 <= two_max;
  if (smem > 48 * 1024) {
    GPM_CUDA(cudaFuncSetAttribute(fit_small_kernel<D, RR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GPM_CUDA(cudaFuncSetAttribute(fit_small_kernel<D, RR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (two) {
    fit_small_kernel<D, RR, true><<<B, N / 32 * 32, smem, stream>>>(Xb, Yb, N, R, th, theta_dev, theta_stride, alpha, lml, info);
  } else {
    const int threads = (N + 31) / 32 * 32;
    fit_small_kernel<D, RR, false><<<B, threads, smem, stream>>>(Xb, Yb, N, R, th, theta_dev, theta_stride, alpha, lml, info);
  }
  GPM_LAUNCH_CHECK();
  return 0;
}

template <int D>
static int launch_small_d(const double* Xb, const double* Yb, int B, int N, int R, const Theta& th,
                          const double* theta_dev, int theta_stride, double* alpha, double* lml, int* info,
                          cudaStream_t stream, int two_max) {
  if (R <= 1) return launch_small_dr<D, 1>(Xb, Yb, B, N, R, th, theta_dev, theta_stride, alpha, lml, info, stream, two_max);
  if (R <= 2) return launch_small_dr<D, 2>(Xb, Yb, B, N, R, th, theta_dev, theta_stride, alpha, lml, info, stream, two_max);
  if (R <= 4) return launch_small_dr<D, 4>(Xb, Yb, B, N, R, th, theta_dev, theta_stride, alpha, lml, info, stream, two_max);
  return launch_small_dr<D, 8>(Xb, Yb, B, N, R, th, theta_dev, theta_stride, alpha, lml, info, stream, two_max);
}

// B paths of N <= SMALL_MAX_N samples each, one CTA per path.  theta_dev: optional per-path hyper-parameters (device).
int launch_fit_small(const double* Xb, const double* Yb, long long B, long long N, int D, int R, const Theta& th,
                     const double* theta_dev, int theta_stride, double* alpha, double* lml, int* info,
                     cudaStream_t stream, int two_max) {
  return D == 2 ? launch_small_d<2>(Xb, Yb, (int)B, (int)N, R, th, theta_dev, theta_stride, alpha, lml, info, stream, two_max)
                : launch_small_d<3>(Xb, Yb, (int)B, (int)N, R, th, theta_dev, theta_stride, alpha, lml, info, stream, two_max);
}

}  // namespace gpm
