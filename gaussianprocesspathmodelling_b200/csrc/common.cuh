// Shared device/host helpers for libgpmap_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gpmap_b200.h"
#include "exp_neg.cuh"

#define NB GPM_NB            // factorisation block / GEMM tile edge (128)
#define SLAB_K 16            // contraction depth of one TMA slab: 16 doubles = 128 B = one swizzle row
#define SLAB_BYTES (NB * SLAB_K * 8)   // 16 KB per operand slab

namespace gpm {

// ----------------------------------------------------------------------------------------------
// host-side error plumbing
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(long long n);          // process-wide count of kernels this library launched
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define GPM_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t _e = (call);                                                   \
    if (_e != cudaSuccess) return gpm::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define GPM_LAUNCH_CHECK()                                                     \
  do {                                                                         \
    gpm::count_launch(1);                                                      \
    cudaError_t _e = cudaPeekAtLastError();                                    \
    if (_e != cudaSuccess) return gpm::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define GPM_ARG(cond, idx)                                                     \
  do {                                                                         \
    if (!(cond)) { gpm::set_error("argument %d invalid: %s", (idx), #cond); return -(idx); } \
  } while (0)

struct Theta {             // kernel-parameter form of theta (host pointer is read at call time)
  double l[3];             // lengthscales; coordinates are divided by l_d (as the oracle's X / ls)
  double sf2, sn2;
  int D;
};

inline int make_theta(const double* theta, int D, Theta* out) {
  if (!theta || (D != 2 && D != 3)) return 1;
  out->D = D;
  for (int d = 0; d < 3; d++) out->l[d] = 1.0;
  for (int d = 0; d < D; d++) {
    if (!(theta[d] > 0.0)) return 1;
    out->l[d] = theta[d];
  }
  out->sf2 = theta[D];
  out->sn2 = theta[D + 1];
  if (!(out->sf2 > 0.0) || !(out->sn2 >= 0.0)) return 1;
  return 0;
}

// Debug / comparison switches.  Read ONCE from the environment (GPM_<NAME>) when the handle is created and
// changed afterwards only through gpm_set_option(); the hot path never calls getenv().
struct Options {
  int no_lookahead = 0;      // potrf: no look-ahead stream
  int tpc_wide = 4, tpc_narrow = 8;                    // potrf: tiles per CTA of the look-ahead updates
  int wide_min = 36, wide4_min = 64, wide8_min = 96;   // potrf: outer panel width thresholds (36: N = 4096, 32 block columns, is faster all-narrow: 2.00 vs 2.15 ms; N = 6144 keeps its wide start: 4.16 vs 4.30 ms; tools/potrf_sweep3.sh)
  int no_separable = 0;      // grid queries through the pointwise kernels
  int no_small_fused = 0;    // short paths through the tiled pipeline
  int small_two_max = 12;    // short paths, 32 < N <= 32 + this: the rows beyond 32 ride as second rows of the first threads (0 = one thread per row always)
  int no_small_tiles = 0;    // no latency tile kernel
  int no_split_column = 0;   // look-ahead Cholesky: update the next panel's whole column before its diagonal block (no split)
  int half_stages = 0;       // half-tile kernel: cap on the ring depth of the plain-store launches (0 = auto: 4, or 3 with more than 4 right-hand sides)
  int half_warps = 8;        // half-tile kernel: warps per CTA (8: 32 x 32 warp tiles, four warps per scheduler; 4: 64 x 32)
  int no_half_tiles = 0;     // batched launches through the 128 x 128 tile kernel instead of the half-tile one
  int no_fused_fwd = 0;      // batched fits: separate forward substitution
  int batch_width = 4;       // potrf, whole-batch launches: outer panel width in block columns (left-looking inside the panel; 4096 x 512: 10.90 / 10.69 / 10.66 ms at 1 / 2 / 4, 1024 x 1024: 15.27 / 15.12 / 15.18 ms at 2 / 4 / 8)
  int no_scratch_factor = 0; // batched fits: potf2 stores the whole lower triangle of L_kk although only its diagonal is read afterwards
  int no_fused_mean = 0;     // predict: separate mean kernel
  int var_steps = 0;         // variance: one launch per block column
  int solve_steps = 0;       // solves: one launch per block step
  int grad_sweep = 0;        // gradient: inverse by a triangular sweep
  int no_path_fused = 0;     // batched fits of 112 < N <= 1024: always the tiled batched pipeline
  int path_fused = 1;        // ... 0 = never one CTA per path, 1 = by batch size (batched.cu: use_path_fit), 2 = always
  int no_fused_solve = 0;    // single-matrix fit: separate forward substitution
};

struct gpm_handle_impl {
  int device;
  Options opt;
  int sm_count;
  cudaStream_t aux;                 // high-priority helper stream for the look-ahead panel
  cudaEvent_t* ev;                  // event pool (cudaEventDisableTiming)
  int n_ev;
  PFN_cuTensorMapEncodeTiled_v12000 encode;
  int* flags;                       // 2 x n_flags device ints: block-published flags of the chained solves
  int n_flags;                      //   (cleared on the stream at the start of every solve: graph-replay safe); a third
                                    //   array of n_flags publishes the per-block LML shares of the backward pass
  double* lml_part;                 // n_flags x 9 doubles: per-block shares of the log marginal likelihood
  bool scratch_factor;              // set by gpm_fit_batched around its factorisation: L is scratch there (only alpha and the LML leave),
                                    //   so potf2 stores just the diagonal 8 x 8 tiles of L_kk (the log-determinant reads the diagonal)
  bool gemm_attr, potf2_attr, gemm_small_attr, gemm_strip_attr, pathfit_attr, gemm_half_attr;       // opt-in shared-memory sizes set for this handle's device (function attributes are per device)
};

// Entry points run on the handle's device whatever the caller's current device is, and restore it on return.
struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int device) : dev(device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// 2-D row-major float64 tensor map with a [rows_box x 16] box and 128-byte swizzle.
int make_tmap(gpm_handle_impl* h, CUtensorMap* map, const double* base, int64_t rows, int64_t cols,
              int64_t ld, int rows_box);

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// device primitives
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// generic-proxy writes -> async-proxy (TMA) reads ordering
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// TMA: 2-D tile global -> shared, completion on an mbarrier.  c0 = column (inner), c1 = row.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// lane = 4*g + q:  a = A[g][q],  b = B[q][g],  c0,c1 = C[g][2q], C[g][2q+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// Byte offset, inside a [rows x 16 doubles] slab written by TMA with CU_TENSOR_MAP_SWIZZLE_128B,
// of the element a lane reads for k4-step t (0..3):  row = (8-aligned tile row) + g, and the
// contraction index  kidx(t, q) = 2t + 8(q>>1) + (q&1).  Both operands use the same kidx, so the
// permutation of the contraction order is harmless; it makes the 16 lanes of a half-warp hit 16
// distinct 8-byte bank pairs (chunks {t, t+4} XOR g cover all eight 16-byte chunks).
__device__ __forceinline__ uint32_t frag_off(int g, int q, int t) {
  int chunk = (t + 4 * (q >> 1)) ^ g;
  return (uint32_t)(g * 128 + chunk * 16 + (q & 1) * 8);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// RBF kernel value from pre-scaled coordinates (x/l): sf2 * exp(-0.5 * |a-b|^2).  The squared
// distance is summed unfused and in dimension order, as scipy's cdist 'sqeuclidean' does, so K
// differs from the oracle's only by the exp() implementation (gpm_exp_neg: < 0.51 ulp, 12 FP64-pipe instructions).
template <int D>
__device__ __forceinline__ double rbf(const double* a, const double* b, double sf2) {
  double dx = a[0] - b[0], dy = a[1] - b[1];
  double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  if (D == 3) {
    double dz = a[2] - b[2];
    d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
  }
  return sf2 * gpm_exp_neg_half(d2);
}

// the same with a caller-provided copy of the exp table (shared memory, gpm_exp_stage_table): bitwise the same value
template <int D>
__device__ __forceinline__ double rbf_t(const double* a, const double* b, double sf2, const gpm_exp_pair* tab) {
  double dx = a[0] - b[0], dy = a[1] - b[1];
  double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  if (D == 3) {
    double dz = a[2] - b[2];
    d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
  }
  return sf2 * gpm_exp_neg_half_t(d2, tab);
}

// coordinates of grid point m (matches numpy.linspace: start + i*step, last point = stop exactly)
__device__ __forceinline__ void grid_point(const gpm_grid_t& g, int64_t m, double& x, double& y) {
  int64_t iy = m / g.gx;
  int ix = (int)(m - iy * g.gx);
  double sx = g.gx > 1 ? (g.x1 - g.x0) / (double)(g.gx - 1) : 0.0;
  double sy = g.gy > 1 ? (g.y1 - g.y0) / (double)(g.gy - 1) : 0.0;
  x = (ix == g.gx - 1 && g.gx > 1) ? g.x1 : __dadd_rn(__dmul_rn((double)ix, sx), g.x0);
  y = (iy == g.gy - 1 && g.gy > 1) ? g.y1 : __dadd_rn(__dmul_rn((double)iy, sy), g.y0);
}
#endif  // __CUDACC__

}  // namespace gpm
