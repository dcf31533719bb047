// exact gradient of the log marginal likelihood w.r.t. the log hyper-parameters,
//   d lml_r / d log(theta_j) = 0.5 * alpha_r^T dK_j alpha_r - 0.5 * tr(K^{-1} dK_j).
// K^{-1} = L^{-T} L^{-1}:  W = L^{-T} by the fused tensor-core sweep on an identity right-hand side
// (row block t is zero left of block column t, so its sweep starts there: N^3/3 flops), then
// K^{-1} = W W^T as a lower-triangular tile SYRK that contracts from the first non-zero block
// (another N^3/3).  One fused pass over K^{-1} then evaluates the kernel derivatives on the fly
// (distance + exp, like the covariance kernel) and reduces all D+2 traces and quadratic forms.
#include <stdlib.h>

#include <algorithm>

#include "gemm.cuh"

namespace gpm {

constexpr int GT = 64;              // reduction tile edge
constexpr int GMAXR = 8;
constexpr int GACC = 4 * (1 + GMAXR);   // (D+1 <= 4 kernel terms) x (trace + R quadratic forms)

__global__ void identity_kernel(double* __restrict__ W, long long N, long long ld, long long rows) {
  const long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x;
  const long long m = blockIdx.x;
  if (i < ld && m < rows) W[m * ld + i] = (i == m && m < N) ? 1.0 : 0.0;
}

// X (lower) and U (upper) start as the inverted diagonal blocks and their transposes.
__global__ void __launch_bounds__(256)
place_diag_kernel(const double* __restrict__ invD, double* __restrict__ Xm, double* __restrict__ Um, long long ld) {
  __shared__ double t[32][33];
  const int k = blockIdx.z, bi = blockIdx.y * 32, bj = blockIdx.x * 32;
  const double* D = invD + (long long)k * NB * NB;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long o = (long long)k * NB;
  for (int r = ty; r < 32; r += 8) {
    const double v = D[(bi + r) * NB + bj + tx];
    t[r][tx] = v;
    Xm[(o + bi + r) * ld + o + bj + tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) Um[(o + bj + r) * ld + o + bi + tx] = t[tx][r];
}

// partial[cta][j*(1+R) + 0]     = sum_ab w * Kinv_ab * dKf_j(a,b)            (j < D: lengthscale d, j = D: signal_var)
// partial[cta][j*(1+R) + 1 + r] = sum_ab w * alpha_ra alpha_rb * dKf_j(a,b)
// partial[cta][GACC + 0]        = sum_a Kinv_aa,   partial[cta][GACC + 1 + r] = sum_a alpha_ra^2     (noise term)
template <int D>
__global__ void __launch_bounds__(256)
lml_grad_partial_kernel(const double* __restrict__ X, long long N, Theta th, const double* __restrict__ Kinv,
                        long long ldk, const double* __restrict__ alpha, int R, double* __restrict__ partial) {
  __shared__ double xi[GT][3], xj[GT][3];
  __shared__ double ai[GT][GMAXR], aj[GT][GMAXR];
  __shared__ double red[8][GACC + 1 + GMAXR];
  const int b = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= b) ti++;
  while (ti * (ti + 1) / 2 > b) ti--;
  const int tj = b - ti * (ti + 1) / 2;
  const long long i0 = (long long)ti * GT, j0 = (long long)tj * GT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 2 * GT * D; e += 256) {
    const int which = e / (GT * D), r = (e % (GT * D)) / D, d = e % D;
    const long long gr = (which ? j0 : i0) + r;
    const double v = (gr < N) ? X[gr * D + d] / th.l[d] : 0.0;
    if (which) xj[r][d] = v; else xi[r][d] = v;
  }
  for (int e = tid; e < 2 * GT * GMAXR; e += 256) {
    const int which = e / (GT * GMAXR), r = (e % (GT * GMAXR)) / GMAXR, k = e % GMAXR;
    const long long gr = (which ? j0 : i0) + r;
    const double v = (gr < N && k < R) ? alpha[gr * R + k] : 0.0;
    if (which) aj[r][k] = v; else ai[r][k] = v;
  }
  __syncthreads();
  const double w = (ti == tj) ? 1.0 : 2.0;          // off-diagonal tiles stand for their mirror image too
  double acc[GACC + 1 + GMAXR];
#pragma unroll
  for (int k = 0; k < GACC + 1 + GMAXR; k++) acc[k] = 0.0;
  const int c = tid & 63;
#pragma unroll 4
  for (int p = 0; p < 16; p++) {
    const int r = (tid >> 6) + 4 * p;
    const long long ga = i0 + r, gb = j0 + c;
    if (ga < N && gb < N) {
      const double kinv = Kinv[ga * ldk + gb];
      double dd[3];
      double d2 = 0.0;
#pragma unroll
      for (int d = 0; d < D; d++) { const double t = xi[r][d] - xj[c][d]; dd[d] = t * t; d2 += dd[d]; }
      const double kf = w * th.sf2 * gpm_exp_neg_half(d2);
      double aa[GMAXR];
#pragma unroll
      for (int k = 0; k < GMAXR; k++) aa[k] = ai[r][k] * aj[c][k];
#pragma unroll
      for (int j = 0; j <= D; j++) {
        const double dk = (j < D) ? kf * dd[j] : kf;
        acc[j * (1 + GMAXR)] = fma(kinv, dk, acc[j * (1 + GMAXR)]);
#pragma unroll
        for (int k = 0; k < GMAXR; k++) acc[j * (1 + GMAXR) + 1 + k] = fma(aa[k], dk, acc[j * (1 + GMAXR) + 1 + k]);
      }
      if (ga == gb) {
        acc[GACC] += kinv;
#pragma unroll
        for (int k = 0; k < GMAXR; k++) acc[GACC + 1 + k] += aa[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < GACC + 1 + GMAXR; k++) acc[k] = warp_sum(acc[k]);
  if (lane == 0) for (int k = 0; k < GACC + 1 + GMAXR; k++) red[warp][k] = acc[k];
  __syncthreads();
  if (tid < GACC + 1 + GMAXR) {
    double s = 0.0;
    for (int wv = 0; wv < 8; wv++) s += red[wv][tid];
    partial[(long long)b * (GACC + 1 + GMAXR) + tid] = s;
  }
}

// grad[r][j] = 0.5 * (Q[r][j] - T[j]); fixed-order sum over the per-tile partials (deterministic)
__global__ void __launch_bounds__(64)
lml_grad_final_kernel(const double* __restrict__ partial, int ntiles, int D, int R, double sn2,
                      double* __restrict__ grad) {
  __shared__ double tot[GACC + 1 + GMAXR];
  const int tid = threadIdx.x;
  if (tid < GACC + 1 + GMAXR) {
    double s = 0.0;
    for (int t = 0; t < ntiles; t++) s += partial[(long long)t * (GACC + 1 + GMAXR) + tid];
    tot[tid] = s;
  }
  __syncthreads();
  if (tid < R * (D + 2)) {
    const int r = tid / (D + 2), j = tid % (D + 2);
    double q, tr;
    if (j <= D) { q = tot[j * (1 + GMAXR) + 1 + r]; tr = tot[j * (1 + GMAXR)]; }
    else { q = sn2 * tot[GACC + 1 + r]; tr = sn2 * tot[GACC]; }
    grad[r * (D + 2) + j] = 0.5 * (q - tr);
  }
}

static inline long long round_up_g(long long a, long long b) { return (a + b - 1) / b * b; }

}  // namespace gpm

using namespace gpm;

extern "C" size_t gpm_lml_grad_workspace_bytes(int64_t N) {
  if (N <= 0) return 0;
  const long long np = round_up_g(N, NB);
  const long long t = (N + GT - 1) / GT;
  return (size_t)(3 * np * np + np + t * (t + 1) / 2 * (GACC + 1 + GMAXR)) * sizeof(double);
}

extern "C" int gpm_lml_grad(gpm_handle_t handle, const double* X, int64_t N, int32_t D, const double* theta,
                            const double* L, int64_t ldl, const void* potrf_ws, const double* alpha, int32_t R,
                            double* grad, void* ws, size_t ws_bytes, gpm_stream_t stream) {
  GPM_ARG(handle != nullptr, 1);
  GPM_ARG(X != nullptr, 2);
  GPM_ARG(N > 0, 3);
  Theta th;
  GPM_ARG(make_theta(theta, D, &th) == 0, 5);
  GPM_ARG(L != nullptr && ((uintptr_t)L & 15) == 0, 6);
  GPM_ARG(ldl >= N && (ldl & 1) == 0, 7);
  GPM_ARG(potrf_ws != nullptr, 8);
  GPM_ARG(alpha != nullptr, 9);
  GPM_ARG(R >= 1 && R <= GMAXR, 10);
  GPM_ARG(grad != nullptr, 11);
  GPM_ARG(ws != nullptr && ((uintptr_t)ws & 15) == 0, 12);
  GPM_ARG(ws_bytes >= gpm_lml_grad_workspace_bytes(N), 13);
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long np = round_up_g(N, NB);
  const int nblk = (int)(np / NB);
  double* W = reinterpret_cast<double*>(ws);
  double* Kinv = W + np * np;              // also holds X = L^{-1} during the recursion
  double* rowsq = Kinv + 2 * np * np;      // [Kinv | T] then rowsq, partial
  double* partial = rowsq + np;
  CUtensorMap mapW, mapL, mapInv, mapKi;
  int rc;
  if ((rc = make_tmap(h, &mapW, W, np, np, np, NB))) return rc;
  if ((rc = make_tmap(h, &mapL, L, N, N, ldl, NB))) return rc;
  if ((rc = make_tmap(h, &mapInv, reinterpret_cast<const double*>(potrf_ws), (long long)nblk * NB, NB, NB, NB))) return rc;
  if ((rc = make_tmap(h, &mapKi, Kinv, np, np, np, NB))) return rc;

  if (h->opt.grad_sweep) {
    // reference path: W = I, then W <- W L^{-T} by the triangular sweep (row block 0 is a long serial chain)
    dim3 gi((unsigned)np, (unsigned)((np + 255) / 256));
    identity_kernel<<<gi, 256, 0, st>>>(W, N, np, np);
    GPM_LAUNCH_CHECK();
    GPM_CUDA(cudaMemsetAsync(rowsq, 0, (size_t)np * sizeof(double), st));
    GemmArgs a = {};
    a.C = W; a.ldc = np; a.rowsq = rowsq;
    a.tiles_m = nblk; a.tiles_n = 1; a.tri = 0;
    a.c_row0 = 0; a.c_rows_end = np; a.c_cols_end = np;
    a.sweep_nblk = nblk; a.sweep_tri = 1; a.epi = EPI_STORE; a.klen = NB;
    if ((rc = launch_gemm(h, mapW, mapL, mapW, a, 1, st, &mapInv))) return rc;
  } else {
    // Recursive doubling: for blocks of size ns = 128, 256, ... and each pair (1, 2) of adjacent blocks
    //   T   = U11 L21^T            X21 = -X22 T^T            U12 = -T X22^T
    // with X = L^{-1} (lower) and U = L^{-T} (upper) both kept, so that every product is an NT GEMM with
    // the contraction index contiguous; triangular operands skip their zero halves per tile.  All pairs of a
    // level run in one batched launch: N^3/2 flops in fully parallel launches.  U ends up in W.
    double* Xm = Kinv;                       // K^{-1} is formed afterwards, so its storage holds X until then
    double* Tm = Xm + np * np;               // (workspace sized for it)
    CUtensorMap mapX, mapT;
    if ((rc = make_tmap(h, &mapX, Xm, np, np, np, NB))) return rc;
    if ((rc = make_tmap(h, &mapT, Tm, np, np, np, NB))) return rc;
    GPM_CUDA(cudaMemsetAsync(W, 0, (size_t)np * np * sizeof(double), st));
    GPM_CUDA(cudaMemsetAsync(Xm, 0, (size_t)np * np * sizeof(double), st));
    place_diag_kernel<<<dim3(4, 4, nblk), 256, 0, st>>>(reinterpret_cast<const double*>(potrf_ws), Xm, W, np);
    GPM_LAUNCH_CHECK();
    for (long long ns = NB; ns < np; ns *= 2) {
      const long long nb_s = (np + ns - 1) / ns;
      const int pairs = (int)(nb_s / 2);
      const long long n2_last = std::min(ns, np - ((long long)(2 * pairs - 1) * ns));
      const int full = (n2_last == ns) ? pairs : pairs - 1;
      for (int pass = 0; pass < 2; pass++) {          // pass 0: the full-size pairs (batched); pass 1: a ragged last pair
        const int batch = pass == 0 ? full : (pairs - full);
        if (batch <= 0) continue;
        const long long o1 = pass == 0 ? 0 : (long long)(2 * full) * ns, o2 = o1 + ns;
        const long long n2 = pass == 0 ? ns : n2_last;
        const int t1 = (int)(ns / NB), t2 = (int)(n2 / NB);
        GemmArgs g = {};
        g.batch_a_rows = g.batch_b_rows = g.batch_c_rows = 2 * ns; g.batch_cols = 2 * ns;
        g.b_tile_rows = NB; g.tri = 0; g.rowsq = nullptr;
        // T[a][c] = sum_i U11[a][i] L21[c][i]          (U11 upper: contract from the tile's own block on)
        g.C = Tm; g.ldc = np; g.tiles_m = t1; g.tiles_n = t2;
        g.a_row0 = (int)o1; g.a_col0 = (int)o1; g.b_row0 = (int)o2; g.b_col0 = (int)o1; g.klen = (int)ns;
        g.c_row0 = o1; g.c_col0 = o2; g.c_rows_end = np; g.c_cols_end = np;
        g.epi = EPI_STORE; g.kstart_mode = 1; g.kend_mode = 0;
        if ((rc = launch_gemm(h, mapW, mapL, mapT, g, batch, st))) return rc;
        // X21[b][a] = -sum_c X22[b][c] T[a][c]         (X22 lower: contract up to the tile's own block)
        g.C = Xm; g.tiles_m = t2; g.tiles_n = t1;
        g.a_row0 = (int)o2; g.a_col0 = (int)o2; g.b_row0 = (int)o1; g.b_col0 = (int)o2; g.klen = (int)n2;
        g.c_row0 = o2; g.c_col0 = o1;
        g.epi = EPI_NEG; g.kstart_mode = 0; g.kend_mode = 1;
        if ((rc = launch_gemm(h, mapX, mapT, mapX, g, batch, st))) return rc;
        // U12[a][b] = -sum_c T[a][c] X22[b][c]
        g.C = W; g.tiles_m = t1; g.tiles_n = t2;
        g.a_row0 = (int)o1; g.a_col0 = (int)o2; g.b_row0 = (int)o2; g.b_col0 = (int)o2; g.klen = (int)n2;
        g.c_row0 = o1; g.c_col0 = o2;
        g.epi = EPI_NEG; g.kstart_mode = 0; g.kend_mode = 2;
        if ((rc = launch_gemm(h, mapT, mapX, mapW, g, batch, st))) return rc;
      }
    }
  }
  // K^{-1} = W W^T on the tiles on/below the diagonal, contracting from the first non-zero block column
  GemmArgs s = {};
  s.C = Kinv; s.ldc = np; s.rowsq = nullptr;
  s.tri = 1; s.tiles_m = nblk; s.tiles_n = nblk; s.kstart_mode = 1;
  s.a_row0 = 0; s.b_row0 = 0; s.a_col0 = 0; s.b_col0 = 0; s.b_tile_rows = NB; s.klen = (int)np;
  s.c_row0 = 0; s.c_col0 = 0; s.c_rows_end = np; s.c_cols_end = np;
  s.epi = EPI_STORE;
  if ((rc = launch_gemm(h, mapW, mapW, mapKi, s, 1, st))) return rc;
  // fused reduction over the lower triangle of K^{-1}
  const int t64 = (int)((N + GT - 1) / GT);
  const int ntiles = t64 * (t64 + 1) / 2;
  if (D == 2) lml_grad_partial_kernel<2><<<ntiles, 256, 0, st>>>(X, N, th, Kinv, np, alpha, R, partial);
  else lml_grad_partial_kernel<3><<<ntiles, 256, 0, st>>>(X, N, th, Kinv, np, alpha, R, partial);
  GPM_LAUNCH_CHECK();
  lml_grad_final_kernel<<<1, 64, 0, st>>>(partial, ntiles, D, R, th.sn2, grad);
  GPM_LAUNCH_CHECK();
  return 0;
}
