// gpm_fit: the whole single-matrix fit in one call -- covariance (lower tiles) -> blocked Cholesky with the forward
// substitution riding along (potf2 forms z_k = inv(L_kk) r_k, the panel-solve epilogues update r_i -= L_ik z_k with
// the tile still in registers) -> backward substitution over the chained solve kernel -> log marginal likelihood.
// Against gpm_cov + gpm_potrf + gpm_solve_lml this removes one full pass over L (8 N^2 bytes and, more to the point,
// half of the latency chain of block hand-offs that bounds the solve: DESIGN.md section 4.4).
#include "gemm.cuh"

namespace gpm {

int launch_cov(const double* X, long long N, int D, const Theta& th, double* K, long long ldk,
               int lower_only, int batch, long long batch_x, long long batch_k, cudaStream_t stream,
               const double* theta_dev, int theta_stride);
int potrf_blocked(gpm_handle_impl* h, double* K, long long N, long long ldk, double* invD, int* info,
                  int batch, long long batch_rows, cudaStream_t s0, double* rhs_r, double* rhs_z, int R,
                  long long batch_rhs_rows);
int solve_chain(gpm_handle_impl* h, const double* L, long long N, long long ldl, const double* invD,
                double* alpha, int R, cudaStream_t stream, int first_dir, const double* Y, double* lml);
int solve_blocked(const double* L, long long N, long long ldl, const double* invD, double* alpha, int R,
                  int batch, long long batch_l, long long batch_inv, long long batch_z, cudaStream_t stream);
int launch_lml(const double* L, long long N, long long ldl, const double* Y, const double* alpha, int R,
               double* lml, int batch, long long batch_l, long long batch_y, cudaStream_t stream);

}  // namespace gpm

using namespace gpm;

extern "C" int gpm_fit(gpm_handle_t handle, const double* X, int64_t N, int32_t D, const double* theta,
                       const double* Y, int32_t R, double* K, int64_t ldk, void* ws, double* alpha, double* lml,
                       int32_t* info, gpm_stream_t stream) {
  GPM_ARG(handle != nullptr, 1);
  GPM_ARG(X != nullptr, 2);
  GPM_ARG(N > 0 && N <= (1 << 20), 3);
  GPM_ARG(D == 2 || D == 3, 4);
  Theta th;
  GPM_ARG(theta != nullptr && make_theta(theta, D, &th) == 0, 5);
  GPM_ARG(Y != nullptr, 6);
  GPM_ARG(R >= 1 && R <= 8, 7);
  GPM_ARG(K != nullptr && ((uintptr_t)K & 15) == 0, 8);
  GPM_ARG(ldk >= N && (ldk & 1) == 0, 9);
  GPM_ARG(ws != nullptr && ((uintptr_t)ws & 15) == 0, 10);
  GPM_ARG(alpha != nullptr && alpha != Y, 11);
  GPM_ARG(info != nullptr, 13);
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  double* invD = reinterpret_cast<double*>(ws);
  int rc;
  if ((rc = launch_cov(X, N, D, th, K, ldk, 1, 1, 0, 0, st, nullptr, 0))) return rc;
  GPM_CUDA(cudaMemcpyAsync(alpha, Y, (size_t)N * R * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (h->opt.no_fused_solve || h->opt.solve_steps) {
    if ((rc = potrf_blocked(h, K, N, ldk, invD, info, 1, 0, st, nullptr, nullptr, 0, 0))) return rc;
    rc = h->opt.solve_steps ? solve_blocked(K, N, ldk, invD, alpha, R, 1, 0, 0, 0, st)
                            : solve_chain(h, K, N, ldk, invD, alpha, R, st, 0, Y, lml);
  } else {
    // alpha enters as the running residual and is overwritten block by block with z = L^{-1} Y
    if ((rc = potrf_blocked(h, K, N, ldk, invD, info, 1, 0, st, alpha, alpha, R, N))) return rc;
    rc = solve_chain(h, K, N, ldk, invD, alpha, R, st, 1, Y, lml);    // the LML comes out of the backward pass
  }
  if (rc) return rc;
  if (lml && h->opt.solve_steps) return launch_lml(K, N, ldk, Y, alpha, R, lml, 1, 0, 0, st);
  return 0;
}
