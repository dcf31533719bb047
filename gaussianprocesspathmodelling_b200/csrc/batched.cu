// Batched per-path fits: B independent paths of equal length N.
//   N <= 112 (the reference's 33-point trajectories): one CTA per path, the whole fit in shared memory (small.cu).
//   112 < N <= 1024 (BASELINE config 3, N = 512): one CTA per path on the tensor cores, two paths per SM (pathfit.cu).
//   longer paths (and the option no_path_fused): all paths advance together through batched launches (grid.y = path) of the same kernels the
//   single-matrix path uses: covariance (lower tiles) -> blocked Cholesky -> blocked solves -> LML.
#include <stdlib.h>

#include "gemm.cuh"

namespace gpm {

int launch_cov(const double* X, long long N, int D, const Theta& th, double* K, long long ldk,
               int lower_only, int batch, long long batch_x, long long batch_k, cudaStream_t stream,
               const double* theta_dev, int theta_stride);
int potrf_blocked(gpm_handle_impl* h, double* K, long long N, long long ldk, double* invD, int* info,
                  int batch, long long batch_rows, cudaStream_t s0, double* rhs_r, double* rhs_z, int R,
                  long long batch_rhs_rows);
bool solve_paths_supported(long long N);
int solve_blocked(const double* L, long long N, long long ldl, const double* invD, double* alpha, int R,
                  int batch, long long batch_l, long long batch_inv, long long batch_z, cudaStream_t stream);
int solve_paths(const double* L, long long N, long long ldl, const double* invD, const double* Y,
                double* alpha, double* lml, int R, int batch, long long batch_l, long long batch_inv,
                long long batch_y, cudaStream_t stream, const double* zfwd);
int launch_lml(const double* L, long long N, long long ldl, const double* Y, const double* alpha, int R,
               double* lml, int batch, long long batch_l, long long batch_y, cudaStream_t stream);

bool fit_small_supported(long long N);
int launch_fit_small(const double* Xb, const double* Yb, long long B, long long N, int D, int R, const Theta& th,
                     const double* theta_dev, int theta_stride, double* alpha, double* lml, int* info,
                     cudaStream_t stream, int two_max);

bool path_fit_supported(long long N, int R);
size_t path_fit_workspace_bytes(const gpm_handle_impl* h, long long B, long long N);
int launch_path_fit(gpm_handle_impl* h, const double* Xb, const double* Yb, long long B, long long N, int D, int R,
                    const Theta& th, const double* theta_host, int theta_stride, double* alpha, double* lml,
                    int* info, void* ws, cudaStream_t st);

static inline long long round_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }

// 112 < N <= 1024: which of the two pipelines.  Measured on B200 (tools/ab_batched.py; 4096 x 512: tiled 12.3 ms, one
// CTA per path 12.5 - 12.8 ms; 256 x 512: 0.99 vs 0.97 ms): the whole-batch tiled launches win by 2 - 6 % once the
// batch is several waves of CTAs, the one-CTA-per-path kernel wins for a batch of at most one wave (2 CTAs per SM)
// and is the only one whose scratch does not grow with the batch (per CTA, not per path: 0.7 GB instead of 8.6 GB at
// 4096 x 512), so it also takes the batches whose tiled workspace would pass 32 GB.  Option path_fused: 0 = never,
// 1 = this rule, 2 = always.
static bool use_path_fit(const gpm_handle_impl* h, long long B, long long N, int R) {
  if (!path_fit_supported(N, R) || h->opt.no_path_fused || h->opt.path_fused == 0) return false;
  if (h->opt.path_fused >= 2) return true;
  const long long np = round_up_ll(N, NB);
  return B <= 2ll * h->sm_count || B > 65535 || (double)B * (double)(np * np + np * NB) * 8.0 > 32.0 * (1ull << 30);
}

}  // namespace gpm

using namespace gpm;

extern "C" size_t gpm_fit_batched_workspace_bytes(gpm_handle_t handle, int64_t B, int64_t N) {
  if (!handle || B <= 0 || N <= 0) return 0;
  const gpm_handle_impl* h = reinterpret_cast<const gpm_handle_impl*>(handle);
  if (fit_small_supported(N) && !h->opt.no_small_fused) return (size_t)B * 8 * sizeof(double);       // one CTA per path: only per-path theta is staged
  if (use_path_fit(h, B, N, 1)) return path_fit_workspace_bytes(h, B, N);   // per-CTA scratch, not per path
  const long long np = round_up_ll(N, NB);
  return (size_t)B * (size_t)(np * np + np * NB + 8 + np * 8) * sizeof(double);   // + per-path theta + z = L^{-1} Y
}

extern "C" int gpm_fit_batched(gpm_handle_t handle, const double* Xb, const double* Yb, int64_t B, int64_t N,
                               int32_t D, int32_t R, const double* theta, int64_t theta_stride,
                               double* alpha, double* lml, int32_t* info, void* ws, gpm_stream_t stream) {
  GPM_ARG(handle != nullptr, 1);
  GPM_ARG(Xb != nullptr, 2);
  GPM_ARG(Yb != nullptr, 3);
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  const bool small = fit_small_supported(N) && !h->opt.no_small_fused;
  const bool fused = !small && use_path_fit(h, B, N, R);
  GPM_ARG(B > 0 && (B <= 65535 || ((small || fused) && B < (1ll << 31))), 4);   // gridDim.y / gridDim.x
  GPM_ARG(N > 0 && (small || fused || B * ((N + NB - 1) / NB * NB) < (1ll << 31)), 5);   // TMA row coordinates of the tiled pipeline
  Theta th;
  GPM_ARG(R >= 1 && R <= 8, 7);
  GPM_ARG(theta_stride == 0 || theta_stride == D + 2, 9);
  GPM_ARG(theta != nullptr && (D == 2 || D == 3), 8);
  for (int64_t b = 0; b < (theta_stride ? B : 1); b++) GPM_ARG(make_theta(theta + b * theta_stride, D, &th) == 0, 8);
  GPM_ARG(alpha != nullptr && alpha != Yb, 10);
  GPM_ARG(info != nullptr, 12);
  GPM_ARG(ws != nullptr && ((uintptr_t)ws & 15) == 0, 13);
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (small) {
    // short paths (the reference's 33-sample trajectories): one CTA per path, the whole fit in shared memory
    const double* tdev = nullptr;
    if (theta_stride) {
      GPM_CUDA(cudaMemcpyAsync(ws, theta, (size_t)B * (D + 2) * sizeof(double), cudaMemcpyHostToDevice, st));
      tdev = reinterpret_cast<const double*>(ws);
    }
    return launch_fit_small(Xb, Yb, B, N, D, R, th, tdev, (int)theta_stride, alpha, lml, info, st, h->opt.small_two_max);
  }
  if (fused)   // 112 < N <= 1024: one CTA per path, two paths in flight per SM, K never materialised (pathfit.cu)
    return launch_path_fit(h, Xb, Yb, B, N, D, R, th, theta_stride ? theta : nullptr, (int)theta_stride, alpha, lml, info, ws, st);
  const long long np = round_up_ll(N, NB);
  const int nblk = (int)(np / NB);
  double* Kb = reinterpret_cast<double*>(ws);            // B stacked np x np matrices, ld = np
  double* invD = Kb + (long long)B * np * np;            // B x nblk x NB x NB
  int rc;
  const double* theta_dev = nullptr;
  if (theta_stride) {                                      // per-path theta: staged after the inverse blocks
    double* tdev = invD + (long long)B * nblk * NB * NB;
    GPM_CUDA(cudaMemcpyAsync(tdev, theta, (size_t)B * (D + 2) * sizeof(double), cudaMemcpyHostToDevice, st));
    theta_dev = tdev;
  }
  // the upper 64 x 64 quadrant of a diagonal block has no reader in this pipeline (potf2 loads the packed lower triangle, the
  // diagonal-tile updates compute and store only sub-tiles on / below the diagonal): not generated
  const int lower_mode = h->opt.no_scratch_factor ? 1 : 2;
  if ((rc = launch_cov(Xb, N, D, th, Kb, np, lower_mode, (int)B, N * D, np * np, st, theta_dev, (int)theta_stride))) return rc;
  // Paths short enough for the one-CTA-per-path solve: the forward substitution rides along with the
  // factorisation (alpha holds the running residual, zf receives z = L^{-1} Y), and the solve kernel only
  // runs the backward pass.
  if (solve_paths_supported(N) && !h->opt.no_fused_fwd) {
    double* zf = invD + (long long)B * nblk * NB * NB + (long long)B * 8;
    GPM_CUDA(cudaMemcpyAsync(alpha, Yb, (size_t)B * N * R * sizeof(double), cudaMemcpyDeviceToDevice, st));
    h->scratch_factor = !h->opt.no_scratch_factor;      // nothing below reads L_kk beyond its diagonal (the backward pass uses inv(L_kk))
    rc = potrf_blocked(h, Kb, N, np, invD, info, (int)B, np, st, alpha, zf, R, N);
    h->scratch_factor = false;
    if (rc) return rc;
    return solve_paths(Kb, N, np, invD, Yb, alpha, lml, R, (int)B, np * np, (long long)nblk * NB * NB, N * R, st, zf);
  }
  if ((rc = potrf_blocked(h, Kb, N, np, invD, info, (int)B, np, st, nullptr, nullptr, 0, 0))) return rc;
  rc = solve_paths(Kb, N, np, invD, Yb, alpha, lml, R, (int)B, np * np, (long long)nblk * NB * NB, N * R, st, nullptr);
  if (rc >= 0) return rc;
  GPM_CUDA(cudaMemcpyAsync(alpha, Yb, (size_t)B * N * R * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if ((rc = solve_blocked(Kb, N, np, invD, alpha, R, (int)B, np * np, (long long)nblk * NB * NB, N * R, st))) return rc;
  if (lml) return launch_lml(Kb, N, np, Yb, alpha, R, lml, (int)B, np * np, N * R, st);
  return 0;
}
