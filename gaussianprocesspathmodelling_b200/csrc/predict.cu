// Steps 4+5: posterior mean (fused, K* never stored) and variance on query points.
//
// Variance: W = K*^T L^{-T}, i.e. solve W L^T = K*^T for the query-major cross-covariance
// K*^T (Mc x N, one row per query point), by blocked forward substitution along the columns:
//     W[:,k] = ( K*^T[:,k] - W[:,0:k] L[k,0:k]^T ) inv(L_kk)^T
// Both products are NT tile GEMMs on the FP64 tensor cores (gemm.cu); the second one's epilogue
// accumulates the row sums of squares, so var[m] = signal_var - sum_i W[m,i]^2 needs no extra pass.
// Queries are processed in chunks whose W fits the caller's workspace; chunk sizes are multiples
// of (SM count x 128) rows so that every GEMM launch is a whole number of waves.
#include <stdlib.h>

#include "gemm.cuh"

namespace gpm {

int launch_cross_cov_t(const gpm_handle_impl* h, const double* X, long long N, int D, const Theta& th, const double* Xs,
                       const gpm_grid_t* grid, long long m0, long long M, double* KsT, long long ldks,
                       long long ncols_pad, cudaStream_t stream);
int launch_cross_cov_mean(const double* X, long long N, int D, const Theta& th, const double* alpha, int R,
                          const double* Xs, const gpm_grid_t* grid, long long m0, long long M, double* KsT,
                          long long ldks, long long ncols_pad, double* mu, cudaStream_t stream);
int launch_predict_mean(const gpm_handle_impl* h, const double* X, long long N, int D, const Theta& th, const double* alpha, int R,
                        const double* Xs, const gpm_grid_t* grid, long long m0, long long M, double* mu,
                        cudaStream_t stream);
bool grid_separable_enabled(const gpm_handle_impl* h, const gpm_grid_t* grid, long long m0, long long M);

__global__ void var_finalize_kernel(const double* __restrict__ rowsq, long long M, double base,
                                    double* __restrict__ var) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) var[m] = fmax(base - rowsq[m], 0.0);      // cancellation next to training points can give -1e-16 * base
}

static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// rows of W per chunk for a workspace of ws_bytes (0 = default 16 GiB cap)
static long long chunk_rows(int sm_count, long long N, long long M, size_t ws_bytes) {
  const long long npad = round_up(N, NB);
  const long long per_row = npad * 8 + 8;
  const long long cap_bytes = ws_bytes ? (long long)ws_bytes : (16ll << 30);
  long long rows = cap_bytes / per_row / NB * NB;
  const long long need = round_up(M, NB);
  if (rows >= need) return need;                       // everything fits: one chunk
  const long long wave = (long long)sm_count * NB;     // otherwise whole waves per chunk
  if (rows >= wave) rows = rows / wave * wave;
  return rows;
}

}  // namespace gpm

using namespace gpm;

extern "C" size_t gpm_predict_workspace_bytes(gpm_handle_t handle, int64_t N, int64_t M) {
  if (!handle || N <= 0 || M <= 0) return 0;
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  const long long rows = chunk_rows(h->sm_count, N, M, 0);
  return (size_t)(rows * (round_up(N, NB) * 8 + 8));
}

extern "C" int gpm_predict(gpm_handle_t handle, const double* X, int64_t N, int32_t D, const double* theta,
                           const double* L, int64_t ldl, const void* potrf_ws, const double* alpha,
                           int32_t R, const double* Xs, const gpm_grid_t* grid, int64_t m0, int64_t m1,
                           double* mu, double* var, void* ws, size_t ws_bytes, int32_t flags,
                           gpm_stream_t stream) {
  GPM_ARG(handle != nullptr, 1);
  GPM_ARG(X != nullptr, 2);
  GPM_ARG(N > 0, 3);
  Theta th;
  GPM_ARG(make_theta(theta, D, &th) == 0, 5);
  GPM_ARG(Xs != nullptr || (grid != nullptr && grid->gx > 0 && grid->gy > 0), 11);
  GPM_ARG(m0 >= 0 && m1 >= m0, 13);
  if (!Xs) GPM_ARG(m1 <= (int64_t)grid->gx * grid->gy, 14);
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long M = m1 - m0;
  if (M == 0) return 0;
  int rc;
  bool fuse_mean = false;
  if (flags & GPM_PREDICT_MEAN) {
    GPM_ARG(alpha != nullptr, 9);
    GPM_ARG(R >= 1 && R <= 8, 10);
    GPM_ARG(mu != nullptr, 15);
    // with the variance requested the mean is fused into the cross-covariance pass below
    // (grid queries: the separable mean kernel + the separable cross-covariance are cheaper than the fused pass)
    fuse_mean = (flags & GPM_PREDICT_VAR) && R <= 8 && !h->opt.no_fused_mean &&
                !(Xs == nullptr && grid_separable_enabled(h, grid, m0, M));
    if (!fuse_mean && (rc = launch_predict_mean(h, X, N, D, th, alpha, R, Xs, grid, m0, M, mu, st))) return rc;
  }
  if (!(flags & GPM_PREDICT_VAR)) return 0;
  GPM_ARG(L != nullptr && ((uintptr_t)L & 15) == 0, 6);
  GPM_ARG(ldl >= N && (ldl & 1) == 0, 7);
  GPM_ARG(potrf_ws != nullptr, 8);
  GPM_ARG(var != nullptr, 16);
  GPM_ARG(ws != nullptr && ((uintptr_t)ws & 15) == 0, 17);
  const long long npad = round_up(N, NB);
  const int nblk = (int)(npad / NB);
  const long long rows = chunk_rows(h->sm_count, N, M, ws_bytes);
  GPM_ARG(rows >= NB, 18);
  double* W = reinterpret_cast<double*>(ws);
  double* rowsq = W + rows * npad;
  CUtensorMap mapW, mapL, mapInv;
  if ((rc = make_tmap(h, &mapW, W, rows, npad, npad, NB))) return rc;
  if ((rc = make_tmap(h, &mapL, L, N, N, ldl, NB))) return rc;
  if ((rc = make_tmap(h, &mapInv, reinterpret_cast<const double*>(potrf_ws), (long long)nblk * NB, NB, NB, NB))) return rc;
  const double base = th.sf2 + ((flags & GPM_PREDICT_ADD_NOISE) ? th.sn2 : 0.0);

  for (long long c0 = 0; c0 < M; c0 += rows) {
    const long long mc = (M - c0) < rows ? (M - c0) : rows;
    const int tiles = (int)((mc + NB - 1) / NB);
    GPM_CUDA(cudaMemsetAsync(rowsq, 0, (size_t)mc * sizeof(double), st));
    if (fuse_mean)
      rc = launch_cross_cov_mean(X, N, D, th, alpha, R, Xs, grid, m0 + c0, mc, W, npad, npad, mu + c0 * R, st);
    else
      rc = launch_cross_cov_t(h, X, N, D, th, Xs, grid, m0 + c0, mc, W, npad, npad, st);
    if (rc) return rc;
    const bool per_step = h->opt.var_steps != 0;        // debugging: one launch per block column
    GemmArgs a = {};
    a.C = W; a.ldc = npad;
    a.tiles_m = tiles; a.tiles_n = 1; a.tri = 0;
    a.a_row0 = 0; a.b_tile_rows = 0;
    a.c_row0 = 0; a.c_rows_end = mc;
    if (!per_step) {
      // the whole blocked forward substitution in one persistent launch
      a.c_cols_end = npad; a.rowsq = rowsq; a.sweep_nblk = nblk; a.epi = EPI_STORE; a.klen = NB;
      if ((rc = launch_gemm(h, mapW, mapL, mapW, a, 1, st, &mapInv))) return rc;
    } else {
      for (int k = 0; k < nblk; k++) {
        a.b_row0 = k * NB;
        a.c_col0 = (long long)k * NB;
        a.c_cols_end = (long long)(k + 1) * NB;
        if (k > 0) {                       // W[:,k] -= W[:,0:k] L[k,0:k]^T
          a.a_col0 = 0; a.b_col0 = 0; a.klen = k * NB; a.epi = EPI_SUB; a.rowsq = nullptr; a.tri_b = 0;
          if ((rc = launch_gemm(h, mapW, mapL, mapW, a, 1, st))) return rc;
        }
        a.a_col0 = k * NB; a.b_col0 = 0; a.klen = NB; a.epi = EPI_STORE; a.rowsq = rowsq; a.tri_b = 1;   // W[:,k] *= inv(L_kk)^T
        if ((rc = launch_gemm(h, mapW, mapInv, mapW, a, 1, st))) return rc;
      }
    }
    var_finalize_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, st>>>(rowsq, mc, base, var + c0);
    GPM_LAUNCH_CHECK();
  }
  return 0;
}
