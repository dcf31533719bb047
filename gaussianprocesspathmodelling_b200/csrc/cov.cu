// Step 1 / 4 kernels: fused distance+exp covariance K(X,X)+sn2*I (symmetric-tile trick: each
// off-diagonal 64x64 tile is evaluated once and stored twice, direct and transposed, both
// coalesced), the materialised query-major cross-covariance, and the fused posterior mean.
// CUDA cores only: the input dimension is 2-3, there is no contraction to give a tensor core.
#include <stdlib.h>

#include "common.cuh"

namespace gpm {

constexpr int CT = 64;            // covariance tile edge
constexpr int CT_LD = CT + 1;     // padded smem pitch

template <int D>
__global__ void __launch_bounds__(256)
cov_kernel(const double* __restrict__ X, long long N, Theta th, double* __restrict__ K,
           long long ldk, int lower_only, long long batch_x, long long batch_k,
           const double* __restrict__ theta_dev, int theta_stride) {
  __shared__ double xi[CT][3], xj[CT][3];
  __shared__ double tile[CT][CT_LD];
  __shared__ gpm_exp_pair etab[GPM_EXP_J];
  gpm_exp_stage_table(etab, threadIdx.x, 256);
  if (theta_dev) {                                  // per-path hyper-parameters: [l_1..l_D, sf2, sn2] per batch index
    const double* t = theta_dev + (long long)blockIdx.y * theta_stride;
#pragma unroll
    for (int d = 0; d < D; d++) th.l[d] = t[d];
    th.sf2 = t[D];
    th.sn2 = t[D + 1];
  }
  // lower-triangular tile enumeration
  const int b = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= b) ti++;
  while (ti * (ti + 1) / 2 > b) ti--;
  const int tj = b - ti * (ti + 1) / 2;
  X += blockIdx.y * batch_x;
  K += blockIdx.y * batch_k;
  const long long i0 = (long long)ti * CT, j0 = (long long)tj * CT;
  const int tid = threadIdx.x;
  for (int e = tid; e < 2 * CT * D; e += 256) {
    const int which = e / (CT * D), r = (e % (CT * D)) / D, d = e % D;
    const long long gr = (which ? j0 : i0) + r;
    const double v = (gr < N) ? X[gr * D + d] / th.l[d] : 0.0;
    if (which) xj[r][d] = v; else xi[r][d] = v;
  }
  __syncthreads();
  const int c2 = (tid & 31) * 2;
  const bool diag = (ti == tj);
  // lower_only: 1 = whole 128 x 128 tiles on / below the diagonal (the off-diagonal 64 x 64 tile of a diagonal 128-block is
  // mirrored), 2 = nothing above the diagonal 64-tiles (batched fits: no consumer reads the upper quadrant of a diagonal block)
  const bool mirror = !diag && (!lower_only || (lower_only == 1 && (ti >> 1) == (tj >> 1)));
  double bj0[3], bj1[3];
#pragma unroll
  for (int d = 0; d < D; d++) { bj0[d] = xj[c2][d]; bj1[d] = xj[c2 + 1][d]; }
  if (!diag && i0 + CT <= N && j0 + CT <= N) {
    // interior tile (all but O(N / 64) of them): no bounds or diagonal tests, one pointer walking down the rows.  The
    // kernel is bound by the instruction issue port, not by the FP64 pipe or HBM, so every integer instruction
    // around the 19 FP64 ones of a kernel value counts.
    double* dst = K + (i0 + (tid >> 5)) * ldk + j0 + c2;
#pragma unroll
    for (int p = 0; p < 8; p++) {
      const int r = (tid >> 5) + 8 * p;
      double a[3];
#pragma unroll
      for (int d = 0; d < D; d++) a[d] = xi[r][d];
      const double v0 = rbf_t<D>(a, bj0, th.sf2, etab), v1 = rbf_t<D>(a, bj1, th.sf2, etab);
      if (mirror) { tile[r][c2] = v0; tile[r][c2 + 1] = v1; }
      *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
      dst += 8 * ldk;
    }
  } else
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int r = (tid >> 5) + 8 * p;
    double a[3];
#pragma unroll
    for (int d = 0; d < D; d++) a[d] = xi[r][d];
    double v0 = rbf_t<D>(a, bj0, th.sf2, etab), v1 = rbf_t<D>(a, bj1, th.sf2, etab);
    const long long gi = i0 + r, gj = j0 + c2;
    if (gi == gj) v0 += th.sn2;
    if (gi == gj + 1) v1 += th.sn2;
    if (mirror) { tile[r][c2] = v0; tile[r][c2 + 1] = v1; }
    if (gi < N) {
      double* dst = K + gi * ldk + gj;
      if (gj + 1 < N) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
      else if (gj < N) *dst = v0;
    }
  }
  if (mirror) {
    __syncthreads();
    const int c = tid & 63;
#pragma unroll
    for (int p = 0; p < 16; p++) {
      const int r = (tid >> 6) + 4 * p;               // row of the transposed tile = column of the direct one
      const long long gi = j0 + r, gj = i0 + c;
      if (gi < N && gj < N) K[gi * ldk + gj] = tile[c][r];
    }
  }
}

// KsT[m, i] = k(xs_m, x_i); 32 query rows x 256 training columns per CTA; columns [N, ncols_pad) zeroed.
template <int D>
__global__ void __launch_bounds__(256)
cross_cov_t_kernel(const double* __restrict__ X, long long N, Theta th, const double* __restrict__ Xs,
                   gpm_grid_t grid, int use_grid, long long m0, long long M, double* __restrict__ KsT,
                   long long ldks, long long ncols_pad) {
  __shared__ double q[32][3];
  __shared__ gpm_exp_pair etab[GPM_EXP_J];
  const int tid = threadIdx.x;
  const long long mb = (long long)blockIdx.x * 32;          // row blocks in x: up to 2^31-1 of them
  gpm_exp_stage_table(etab, tid, 256);
  if (tid < 32) {
    const long long m = mb + tid;
    double c[3] = {0.0, 0.0, 0.0};
    if (m < M) {
      if (use_grid) { grid_point(grid, m0 + m, c[0], c[1]); c[2] = grid.t; }
      else { for (int d = 0; d < D; d++) c[d] = Xs[(m0 + m) * D + d]; }
    }
#pragma unroll
    for (int d = 0; d < D; d++) q[tid][d] = c[d] / th.l[d];
  }
  __syncthreads();
  const long long i = (long long)blockIdx.y * 256 + tid;
  if (i >= ncols_pad) return;
  const int nr = (int)((M - mb) < 32 ? (M - mb) : 32);
  double* dst = KsT + mb * ldks + i;
  if (i >= N) {                                              // padding columns
    for (int r = 0; r < nr; r++) dst[r * ldks] = 0.0;
    return;
  }
  double xi[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < D; d++) xi[d] = X[i * D + d] / th.l[d];
  if (nr == 32) {
    // straight-line groups of eight independent evaluations (no exits inside: the chains interleave and the
    // constants stay in registers)
    double* dp = dst;                    // one walking pointer: no 64-bit multiply per store
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 8) {   // unrolled too: the polynomial's constants are materialised once
      double v[8];
#pragma unroll
      for (int r = 0; r < 8; r++) v[r] = rbf_t<D>(q[r0 + r], xi, th.sf2, etab);
#pragma unroll
      for (int r = 0; r < 8; r++) { *dp = v[r]; dp += ldks; }
    }
  } else {
    for (int r = 0; r < nr; r++) dst[r * ldks] = rbf_t<D>(q[r], xi, th.sf2, etab);
  }
}

// Fused cross-covariance + posterior mean (used when the variance is wanted too, so K*^T has to be
// materialised anyway): a CTA owns QR query rows (8 for R <= 2, 4 for more right-hand sides: the QR x R partial
// sums live in registers) and walks all training columns, 256 at a time; every
// k(xs_m, x_i) is evaluated once, stored to KsT (coalesced 2 KB row segments) and accumulated into the
// mean; the 8 x R partial sums are reduced across the CTA at the end (fixed order: deterministic).
template <int D, int RR, int QR>
__global__ void __launch_bounds__(256)
cross_cov_mean_kernel(const double* __restrict__ X, long long N, Theta th, const double* __restrict__ alpha, int R,
                      const double* __restrict__ Xs, gpm_grid_t grid, int use_grid, long long m0, long long M,
                      double* __restrict__ KsT, long long ldks, long long ncols_pad, double* __restrict__ mu) {
  __shared__ double q[QR][3];
  __shared__ double red[8][QR * RR];
  __shared__ gpm_exp_pair etab[GPM_EXP_J];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long mb = (long long)blockIdx.x * QR;
  gpm_exp_stage_table(etab, tid, 256);
  if (tid < QR) {
    const long long m = mb + tid;
    double c[3] = {0.0, 0.0, 0.0};
    if (m < M) {
      if (use_grid) { grid_point(grid, m0 + m, c[0], c[1]); c[2] = grid.t; }
      else { for (int d = 0; d < D; d++) c[d] = Xs[(m0 + m) * D + d]; }
    }
#pragma unroll
    for (int d = 0; d < D; d++) q[tid][d] = c[d] / th.l[d];
  }
  __syncthreads();
  double qr[QR][3];
#pragma unroll
  for (int r = 0; r < QR; r++)
#pragma unroll
    for (int d = 0; d < D; d++) qr[r][d] = q[r][d];
  double acc[QR][RR];
#pragma unroll
  for (int r = 0; r < QR; r++)
#pragma unroll
    for (int k = 0; k < RR; k++) acc[r][k] = 0.0;
  for (long long i = tid; i < ncols_pad; i += 256) {
    double xi[3] = {0.0, 0.0, 0.0}, a[RR];
#pragma unroll
    for (int k = 0; k < RR; k++) a[k] = 0.0;
    const bool in = i < N;
    if (in) {
#pragma unroll
      for (int d = 0; d < D; d++) xi[d] = X[i * D + d] / th.l[d];
#pragma unroll
      for (int k = 0; k < RR; k++) if (k < R) a[k] = alpha[i * R + k];
    }
    // all QR evaluations first (independent chains, no branches between them), then the stores; rows beyond M read
    // the zero coordinates staged above and are simply not stored
    double v[QR];
#pragma unroll
    for (int r = 0; r < QR; r++) v[r] = in ? rbf_t<D>(qr[r], xi, th.sf2, etab) : 0.0;
#pragma unroll
    for (int r = 0; r < QR; r++) {
      if (mb + r < M) KsT[(mb + r) * ldks + i] = v[r];
#pragma unroll
      for (int k = 0; k < RR; k++) acc[r][k] = fma(v[r], a[k], acc[r][k]);
    }
  }
#pragma unroll
  for (int r = 0; r < QR; r++)
#pragma unroll
    for (int k = 0; k < RR; k++) {
      const double sv = warp_sum(acc[r][k]);
      if (lane == 0) red[warp][r * RR + k] = sv;
    }
  __syncthreads();
  if (tid < QR * RR) {
    const int r = tid / RR, k = tid % RR;
    double sv = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) sv += red[w][tid];
    if (mb + r < M && k < R) mu[(mb + r) * R + k] = sv;
  }
}

// mu[m, r] = sum_i k(xs_m, x_i) alpha[i, r]: one thread per query, training points and alpha staged
// through shared memory in chunks; K* is never stored.
template <int D, int RR>
__global__ void __launch_bounds__(256)
predict_mean_kernel(const double* __restrict__ X, long long N, Theta th, const double* __restrict__ alpha,
                    int R, const double* __restrict__ Xs, gpm_grid_t grid, int use_grid, long long m0,
                    long long M, double* __restrict__ mu) {
  constexpr int CH = 512;
  __shared__ double sx[CH][3];
  __shared__ double sa[CH][RR];
  __shared__ gpm_exp_pair etab[GPM_EXP_J];
  const int tid = threadIdx.x;
  gpm_exp_stage_table(etab, tid, 256);
  const long long m = (long long)blockIdx.x * 256 + tid;
  double c[3] = {0.0, 0.0, 0.0};
  if (m < M) {
    if (use_grid) { grid_point(grid, m0 + m, c[0], c[1]); c[2] = grid.t; }
    else { for (int d = 0; d < D; d++) c[d] = Xs[(m0 + m) * D + d]; }
  }
  double qs[3];
#pragma unroll
  for (int d = 0; d < 3; d++) qs[d] = d < D ? c[d] / th.l[d] : 0.0;
  double acc[RR];
#pragma unroll
  for (int r = 0; r < RR; r++) acc[r] = 0.0;
  for (long long i0 = 0; i0 < N; i0 += CH) {
    const int n = (int)((N - i0) < CH ? (N - i0) : CH);
    __syncthreads();
    for (int e = tid; e < n * D; e += 256) { const int r = e / D, d = e % D; sx[r][d] = X[(i0 + r) * D + d] / th.l[d]; }
    for (int e = tid; e < n * RR; e += 256) { const int r = e / RR, k = e % RR; sa[r][k] = k < R ? alpha[(i0 + r) * R + k] : 0.0; }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < n; i++) {
      const double kv = rbf_t<D>(qs, sx[i], th.sf2, etab);
#pragma unroll
      for (int r = 0; r < RR; r++) acc[r] = fma(kv, sa[i][r], acc[r]);
    }
  }
  if (m < M) {
#pragma unroll
    for (int r = 0; r < RR; r++) if (r < R) mu[m * R + r] = acc[r];
  }
}


// ----------------------------------------------------------------------------------------------
// Regular-grid queries: the RBF kernel is separable,
//     k((x_a, y_b), p_i) = [exp(-(x_a - p_ix)^2 / 2)] * [sf2 exp(-((y_b - p_iy)^2 [+ (t - p_it)^2]) / 2)]
// (coordinates pre-scaled by the lengthscales), so a PA x PB patch of grid points needs PA + PB exponentials
// per training point instead of PA * PB.  The two kernels below evaluate the factors on the fly per patch (no
// tables in HBM): the materialised cross-covariance becomes a pure HBM-write stream and the mean an FMA loop.
// A query's value does not depend on which patch or launch it falls into (bitwise), so sharded ranges
// concatenate exactly.  Both forms carry the rounding of their exponents (|arg| eps relative), so the product agrees
// with the single exponential of the oracle to (|arg| + 4) eps relative and a few eps * sf2 absolute (a CPU test
// pins that bound).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ double grid_x(const gpm_grid_t& g, long long ix) {
  const double sx = g.gx > 1 ? (g.x1 - g.x0) / (double)(g.gx - 1) : 0.0;
  return (ix == g.gx - 1 && g.gx > 1) ? g.x1 : __dadd_rn(__dmul_rn((double)ix, sx), g.x0);
}
__device__ __forceinline__ double grid_y(const gpm_grid_t& g, long long iy) {
  const double sy = g.gy > 1 ? (g.y1 - g.y0) / (double)(g.gy - 1) : 0.0;
  return (iy == g.gy - 1 && g.gy > 1) ? g.y1 : __dadd_rn(__dmul_rn((double)iy, sy), g.y0);
}

// KsT[m - m0, i] for the grid points m in [m0, m0 + M): a CTA owns a patch of 8 grid columns x 8 grid rows and a
// run of `cols_per_cta` training columns, one column per thread and step; every store instruction of a warp
// writes 256 contiguous bytes of one row.  Columns [N, ncols_pad) are zeroed.
template <int D>
__global__ void __launch_bounds__(256, 3)
cross_cov_grid_kernel(const double* __restrict__ X, long long N, Theta th, gpm_grid_t grid, long long m0,
                      long long M, long long row_first, double* __restrict__ KsT, long long ldks,
                      long long ncols_pad, int cols_per_cta) {
  constexpr int PA = 8, PB = 8;
  __shared__ double sqy[PB];
  const int a0 = blockIdx.x * PA;
  const long long b0 = row_first + (long long)blockIdx.y * PB;
  if (threadIdx.x < PB) sqy[threadIdx.x] = grid_y(grid, b0 + threadIdx.x) / th.l[1];
  double qx[PA];
#pragma unroll
  for (int r = 0; r < PA; r++) qx[r] = grid_x(grid, a0 + r) / th.l[0];
  const double qt = D == 3 ? grid.t / th.l[2] : 0.0;
  // which of the 64 patch points lie inside [m0, m0 + M) and inside the grid row
  unsigned long long mask = 0;
#pragma unroll
  for (int rb = 0; rb < PB; rb++)
#pragma unroll
    for (int ra = 0; ra < PA; ra++) {
      const long long m = (b0 + rb) * grid.gx + a0 + ra;
      if (a0 + ra < grid.gx && m >= m0 && m < m0 + M) mask |= 1ull << (rb * PA + ra);
    }
  __syncthreads();
  if (mask == 0) return;
  const long long c_begin = (long long)blockIdx.z * cols_per_cta;
  const long long c_end = min(ncols_pad, c_begin + cols_per_cta);
  double* const row0 = KsT + (b0 * grid.gx + a0 - m0) * ldks;      // patch point (0, 0); only masked-in rows are touched
  const long long row_pitch = (long long)grid.gx * ldks;            // one grid row further
  for (long long i = c_begin + threadIdx.x; i < c_end; i += 256) {
    const bool in = i < N;
    double ex[PA], py = 0.0, dt2 = 0.0;
    if (in) {
      const double px = X[i * D] / th.l[0];
      py = X[i * D + 1] / th.l[1];
      if (D == 3) { const double dt = qt - X[i * D + 2] / th.l[2]; dt2 = __dmul_rn(dt, dt); }
#pragma unroll
      for (int r = 0; r < PA; r++) { const double dx = qx[r] - px; ex[r] = gpm_exp_neg_half(__dmul_rn(dx, dx)); }
    } else {
#pragma unroll
      for (int r = 0; r < PA; r++) ex[r] = 0.0;
    }
    double* prow = row0 + i;
#pragma unroll 1
    for (int rb = 0; rb < PB; rb++, prow += row_pitch) {
      const unsigned rm = (unsigned)(mask >> (rb * PA)) & 0xffu;
      if (rm == 0) continue;
      const double dy = sqy[rb] - py;
      const double d2 = D == 3 ? __dadd_rn(__dmul_rn(dy, dy), dt2) : __dmul_rn(dy, dy);
      const double ey = in ? th.sf2 * gpm_exp_neg_half(d2) : 0.0;
#pragma unroll
      for (int ra = 0; ra < PA; ra++)
        if (rm >> ra & 1) prow[ra * ldks] = ex[ra] * ey;
    }
  }
}

// mu[m - m0, r] = sum_i k(grid point m, x_i) alpha[i, r] for m in [m0, m0 + M): a CTA owns a patch of 16 x 16 grid
// points (one per thread); per chunk of 128 training points the 2 x 16 x 128 factors are evaluated once into shared
// memory (16 exponentials per thread) and every thread runs a load-load-multiply-FMA loop over the chunk.
template <int D, int RR>
__global__ void __launch_bounds__(256)
predict_mean_grid_kernel(const double* __restrict__ X, long long N, Theta th, const double* __restrict__ alpha,
                         int R, gpm_grid_t grid, long long m0, long long M, long long row_first,
                         double* __restrict__ mu) {
  constexpr int PA = 16, PB = 16, CH = 128;
  __shared__ double exs[CH][PA], eys[CH][PB], sa[CH][RR], sx[CH][3], sq[2][16];
  const int tid = threadIdx.x, ta = tid & 15, tb = tid >> 4;
  const int a0 = blockIdx.x * PA;
  const long long b0 = row_first + (long long)blockIdx.y * PB;
  if (tid < 16) sq[0][tid] = grid_x(grid, a0 + tid) / th.l[0];
  else if (tid < 32) sq[1][tid - 16] = grid_y(grid, b0 + tid - 16) / th.l[1];
  const double qt = D == 3 ? grid.t / th.l[2] : 0.0;
  double acc[RR];
#pragma unroll
  for (int r = 0; r < RR; r++) acc[r] = 0.0;
  for (long long i0 = 0; i0 < N; i0 += CH) {
    const int n = (int)((N - i0) < CH ? (N - i0) : CH);
    __syncthreads();                       // previous chunk consumed (and sq written, first time round)
    for (int e = tid; e < n * D; e += 256) {
      const int c = e / D, d = e % D;
      sx[c][d] = X[(i0 + c) * D + d] / (d == 0 ? th.l[0] : (d == 1 ? th.l[1] : th.l[2]));
    }
    for (int e = tid; e < n * RR; e += 256) { const int c = e / RR, k = e % RR; sa[c][k] = k < R ? alpha[(i0 + c) * R + k] : 0.0; }
    __syncthreads();
#pragma unroll 4
    for (int e = tid; e < 2 * CH * 16; e += 256) {
      const int which = e / (CH * 16), c = (e / 16) % CH, r = e % 16;     // consecutive threads: consecutive r
      if (c < n) {
        if (which == 0) {
          const double dx = sq[0][r] - sx[c][0];
          exs[c][r] = gpm_exp_neg_half(__dmul_rn(dx, dx));
        } else {
          const double dy = sq[1][r] - sx[c][1];
          double d2 = __dmul_rn(dy, dy);
          if (D == 3) { const double dt = qt - sx[c][2]; d2 = __dadd_rn(d2, __dmul_rn(dt, dt)); }
          eys[c][r] = th.sf2 * gpm_exp_neg_half(d2);
        }
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < n; i++) {
      const double kv = exs[i][ta] * eys[i][tb];
#pragma unroll
      for (int r = 0; r < RR; r++) acc[r] = fma(kv, sa[i][r], acc[r]);
    }
  }
  const long long m = (b0 + tb) * grid.gx + a0 + ta;
  if (a0 + ta < grid.gx && m >= m0 && m < m0 + M) {
#pragma unroll
    for (int r = 0; r < RR; r++) if (r < R) mu[(m - m0) * R + r] = acc[r];
  }
}

// separable grid kernels are used for grid queries unless GPM_NO_SEPARABLE is set (debugging / comparison)
bool grid_separable_enabled(const gpm_handle_impl* h, const gpm_grid_t* grid, long long m0, long long M) {
  if (!grid || M <= 0 || grid->gx <= 0 || h->opt.no_separable) return false;
  const long long rows = (m0 + M - 1) / grid->gx - m0 / grid->gx + 1;
  return (rows + 7) / 8 <= 65535;        // gridDim.y
}

int launch_cov(const double* X, long long N, int D, const Theta& th, double* K, long long ldk,
               int lower_only, int batch, long long batch_x, long long batch_k, cudaStream_t stream,
               const double* theta_dev, int theta_stride) {
  const int T = (int)((N + CT - 1) / CT);
  dim3 grid(T * (T + 1) / 2, batch);
  if (D == 2) cov_kernel<2><<<grid, 256, 0, stream>>>(X, N, th, K, ldk, lower_only, batch_x, batch_k, theta_dev, theta_stride);
  else cov_kernel<3><<<grid, 256, 0, stream>>>(X, N, th, K, ldk, lower_only, batch_x, batch_k, theta_dev, theta_stride);
  GPM_LAUNCH_CHECK();
  return 0;
}

int launch_cross_cov_t(const gpm_handle_impl* h, const double* X, long long N, int D, const Theta& th, const double* Xs,
                       const gpm_grid_t* grid, long long m0, long long M, double* KsT, long long ldks,
                       long long ncols_pad, cudaStream_t stream) {
  if (M <= 0) return 0;
  gpm_grid_t g = {};
  if (grid) g = *grid;
  const int use_grid = Xs == nullptr;
  if (use_grid && grid_separable_enabled(h, grid, m0, M)) {
    const long long row_first = m0 / g.gx, rows = (m0 + M - 1) / g.gx - row_first + 1;
    const int cols_per_cta = 1024;
    dim3 gd((unsigned)((g.gx + 7) / 8), (unsigned)((rows + 7) / 8), (unsigned)((ncols_pad + cols_per_cta - 1) / cols_per_cta));
    if (D == 2) cross_cov_grid_kernel<2><<<gd, 256, 0, stream>>>(X, N, th, g, m0, M, row_first, KsT, ldks, ncols_pad, cols_per_cta);
    else cross_cov_grid_kernel<3><<<gd, 256, 0, stream>>>(X, N, th, g, m0, M, row_first, KsT, ldks, ncols_pad, cols_per_cta);
    GPM_LAUNCH_CHECK();
    return 0;
  }
  dim3 gridDim((unsigned)((M + 31) / 32), (unsigned)((ncols_pad + 255) / 256));
  if (D == 2) cross_cov_t_kernel<2><<<gridDim, 256, 0, stream>>>(X, N, th, Xs, g, use_grid, m0, M, KsT, ldks, ncols_pad);
  else cross_cov_t_kernel<3><<<gridDim, 256, 0, stream>>>(X, N, th, Xs, g, use_grid, m0, M, KsT, ldks, ncols_pad);
  GPM_LAUNCH_CHECK();
  return 0;
}

template <int D>
static int launch_ccm_d(const double* X, long long N, const Theta& th, const double* alpha, int R, const double* Xs,
                        const gpm_grid_t& g, int use_grid, long long m0, long long M, double* KsT, long long ldks,
                        long long ncols_pad, double* mu, cudaStream_t stream) {
  const unsigned b8 = (unsigned)((M + 7) / 8), b4 = (unsigned)((M + 3) / 4);
  if (R <= 1) cross_cov_mean_kernel<D, 1, 8><<<b8, 256, 0, stream>>>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, KsT, ldks, ncols_pad, mu);
  else if (R <= 2) cross_cov_mean_kernel<D, 2, 8><<<b8, 256, 0, stream>>>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, KsT, ldks, ncols_pad, mu);
  else if (R <= 4) cross_cov_mean_kernel<D, 4, 4><<<b4, 256, 0, stream>>>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, KsT, ldks, ncols_pad, mu);
  else if (R <= 8) cross_cov_mean_kernel<D, 8, 4><<<b4, 256, 0, stream>>>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, KsT, ldks, ncols_pad, mu);
  else return -1;
  GPM_LAUNCH_CHECK();
  return 0;
}

// fused cross-covariance + mean; returns -1 when R is too large for the fused kernel (caller falls back to
// the separate mean kernel + plain cross-covariance)
int launch_cross_cov_mean(const double* X, long long N, int D, const Theta& th, const double* alpha, int R,
                          const double* Xs, const gpm_grid_t* grid, long long m0, long long M, double* KsT,
                          long long ldks, long long ncols_pad, double* mu, cudaStream_t stream) {
  if (M <= 0) return 0;
  gpm_grid_t g = {};
  if (grid) g = *grid;
  const int use_grid = Xs == nullptr;
  return D == 2 ? launch_ccm_d<2>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, KsT, ldks, ncols_pad, mu, stream)
                : launch_ccm_d<3>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, KsT, ldks, ncols_pad, mu, stream);
}

template <int D>
static int launch_mean_d(const gpm_handle_impl* h, const double* X, long long N, const Theta& th, const double* alpha, int R,
                         const double* Xs, const gpm_grid_t& g, int use_grid, long long m0, long long M,
                         double* mu, cudaStream_t stream) {
  if (use_grid && grid_separable_enabled(h, &g, m0, M)) {
    const long long row_first = m0 / g.gx, rows = (m0 + M - 1) / g.gx - row_first + 1;
    dim3 gd((unsigned)((g.gx + 15) / 16), (unsigned)((rows + 15) / 16));
    if (R <= 1) predict_mean_grid_kernel<D, 1><<<gd, 256, 0, stream>>>(X, N, th, alpha, R, g, m0, M, row_first, mu);
    else if (R <= 2) predict_mean_grid_kernel<D, 2><<<gd, 256, 0, stream>>>(X, N, th, alpha, R, g, m0, M, row_first, mu);
    else if (R <= 4) predict_mean_grid_kernel<D, 4><<<gd, 256, 0, stream>>>(X, N, th, alpha, R, g, m0, M, row_first, mu);
    else predict_mean_grid_kernel<D, 8><<<gd, 256, 0, stream>>>(X, N, th, alpha, R, g, m0, M, row_first, mu);
    GPM_LAUNCH_CHECK();
    return 0;
  }
  const unsigned blocks = (unsigned)((M + 255) / 256);
  if (R <= 1) predict_mean_kernel<D, 1><<<blocks, 256, 0, stream>>>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, mu);
  else if (R <= 2) predict_mean_kernel<D, 2><<<blocks, 256, 0, stream>>>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, mu);
  else if (R <= 4) predict_mean_kernel<D, 4><<<blocks, 256, 0, stream>>>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, mu);
  else predict_mean_kernel<D, 8><<<blocks, 256, 0, stream>>>(X, N, th, alpha, R, Xs, g, use_grid, m0, M, mu);
  GPM_LAUNCH_CHECK();
  return 0;
}

int launch_predict_mean(const gpm_handle_impl* h, const double* X, long long N, int D, const Theta& th, const double* alpha, int R,
                        const double* Xs, const gpm_grid_t* grid, long long m0, long long M, double* mu,
                        cudaStream_t stream) {
  if (M <= 0) return 0;
  gpm_grid_t g = {};
  if (grid) g = *grid;
  const int use_grid = Xs == nullptr;
  return D == 2 ? launch_mean_d<2>(h, X, N, th, alpha, R, Xs, g, use_grid, m0, M, mu, stream)
                : launch_mean_d<3>(h, X, N, th, alpha, R, Xs, g, use_grid, m0, M, mu, stream);
}

}  // namespace gpm

using namespace gpm;

extern "C" int gpm_cov(gpm_handle_t h, const double* X, int64_t N, int32_t D, const double* theta,
                       double* K, int64_t ldk, int32_t flags, gpm_stream_t stream) {
  GPM_ARG(h != nullptr, 1);
  GPM_ARG(X != nullptr, 2);
  GPM_ARG(N > 0 && N <= (1 << 20), 3);
  Theta th;
  GPM_ARG(make_theta(theta, D, &th) == 0, 5);
  GPM_ARG(K != nullptr && ((uintptr_t)K & 15) == 0, 6);
  GPM_ARG(ldk >= N && (ldk & 1) == 0, 7);
  DeviceGuard guard(reinterpret_cast<gpm_handle_impl*>(h)->device);
  return launch_cov(X, N, D, th, K, ldk, (flags & GPM_COV_LOWER) ? 1 : 0, 1, 0, 0, (cudaStream_t)stream, nullptr, 0);
}

extern "C" int gpm_cross_cov(gpm_handle_t h, const double* X, int64_t N, int32_t D, const double* theta,
                             const double* Xs, const gpm_grid_t* grid, int64_t m0, int64_t m1,
                             double* KsT, int64_t ldks, gpm_stream_t stream) {
  GPM_ARG(h != nullptr, 1);
  GPM_ARG(X != nullptr, 2);
  GPM_ARG(N > 0, 3);
  Theta th;
  GPM_ARG(make_theta(theta, D, &th) == 0, 5);
  GPM_ARG(Xs != nullptr || grid != nullptr, 6);
  GPM_ARG(m0 >= 0 && m1 >= m0, 8);
  GPM_ARG(KsT != nullptr, 10);
  GPM_ARG(ldks >= N, 11);
  if (!Xs) {
    GPM_ARG(grid->gx > 0 && grid->gy > 0, 7);
    GPM_ARG(m1 <= (int64_t)grid->gx * grid->gy, 9);
  }
  gpm_handle_impl* hi = reinterpret_cast<gpm_handle_impl*>(h);
  DeviceGuard guard(hi->device);
  return launch_cross_cov_t(hi, X, N, D, th, Xs, grid, m0, m1 - m0, KsT, ldks, N, (cudaStream_t)stream);
}
