// Step 3: alpha = L^{-T} (L^{-1} Y) by blocked substitution with the inverted diagonal blocks that
// gpm_potrf left in its workspace, then the log marginal likelihood.  HBM-bound: L is read twice.
//
// Forward  (k = -1 .. nblk-2): every row block i > k does  y_i -= L_ik z_k ; block i = k+1 then
//                              finishes  z_i = inv(L_ii) y_i.
// Backward (k = nblk .. 1):    every column block j < k does  z_j -= L_kj^T a_k ; block j = k-1
//                              then finishes  a_j = inv(L_jj)^T z_j.
#include "common.cuh"

namespace gpm {

constexpr int RMAX = 8;

__global__ void __launch_bounds__(256)
fwd_step_kernel(const double* __restrict__ L, long long ldl, long long N, const double* __restrict__ invD,
                double* __restrict__ z, int R, int k, long long batch_l, long long batch_inv,
                long long batch_z) {
  __shared__ double zk[NB][RMAX];
  __shared__ double yi[NB][RMAX];
  L += blockIdx.y * batch_l; invD += blockIdx.y * batch_inv; z += blockIdx.y * batch_z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i = k + 1 + blockIdx.x;
  const long long r0 = (long long)i * NB;
  if (k >= 0) {
    for (int e = tid; e < NB * R; e += 256) zk[e / R][e % R] = z[(long long)k * NB * R + e];
  }
  __syncthreads();
  for (int rr = 0; rr < 16; rr++) {
    const int row = warp * 16 + rr;
    const long long gr = r0 + row;
    double acc[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; r++) acc[r] = 0.0;
    if (k >= 0 && gr < N) {
      const double* lrow = L + gr * ldl + (long long)k * NB;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double l = lrow[lane + 32 * j];
#pragma unroll
        for (int r = 0; r < RMAX; r++) if (r < R) acc[r] = fma(l, zk[lane + 32 * j][r], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RMAX; r++) if (r < R) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
      for (int r = 0; r < R; r++) {
        const double y = (gr < N) ? z[gr * R + r] - acc[r] : 0.0;
        yi[row][r] = y;
        if (blockIdx.x != 0 && gr < N) z[gr * R + r] = y;
      }
    }
  }
  if (blockIdx.x != 0) return;
  __syncthreads();
  const double* Di = invD + (long long)i * NB * NB;
  for (int rr = 0; rr < 16; rr++) {
    const int row = warp * 16 + rr;
    const long long gr = r0 + row;
    double acc[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; r++) acc[r] = 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const double d = Di[row * NB + lane + 32 * j];
#pragma unroll
      for (int r = 0; r < RMAX; r++) if (r < R) acc[r] = fma(d, yi[lane + 32 * j][r], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < RMAX; r++) if (r < R) acc[r] = warp_sum(acc[r]);
    if (lane == 0 && gr < N) for (int r = 0; r < R; r++) z[gr * R + r] = acc[r];
  }
}

__global__ void __launch_bounds__(256)
bwd_step_kernel(const double* __restrict__ L, long long ldl, long long N, const double* __restrict__ invD,
                double* __restrict__ z, int R, int k, int nblk, long long batch_l, long long batch_inv,
                long long batch_z) {
  __shared__ double ak[NB][RMAX];
  __shared__ double part[2][NB][RMAX];
  __shared__ double zj[NB][RMAX];
  L += blockIdx.y * batch_l; invD += blockIdx.y * batch_inv; z += blockIdx.y * batch_z;
  const int tid = threadIdx.x;
  const int j = (k < nblk) ? (int)blockIdx.x : nblk - 1;
  const long long c0 = (long long)j * NB;
  const int c = tid & 127, half = tid >> 7;
  double acc[RMAX];
#pragma unroll
  for (int r = 0; r < RMAX; r++) acc[r] = 0.0;
  if (k < nblk) {
    const long long k0 = (long long)k * NB;
    const int nv = (int)((N - k0) < NB ? (N - k0) : NB);
    for (int e = tid; e < NB * R; e += 256) {
      const int row = e / R;
      ak[row][e % R] = (row < nv) ? z[k0 * R + e] : 0.0;
    }
    __syncthreads();
    const int rend = (half + 1) * 64 < nv ? (half + 1) * 64 : nv;
    for (int row = half * 64; row < rend; row++) {
      const double l = L[(k0 + row) * ldl + c0 + c];
#pragma unroll
      for (int r = 0; r < RMAX; r++) if (r < R) acc[r] = fma(l, ak[row][r], acc[r]);
    }
  }
  for (int r = 0; r < R; r++) part[half][c][r] = acc[r];
  __syncthreads();
  const bool last = (k == nblk) || (j == k - 1);
  if (tid < NB) {
    const long long gc = c0 + c;
    for (int r = 0; r < R; r++) {
      const double v = (gc < N) ? z[gc * R + r] - (part[0][c][r] + part[1][c][r]) : 0.0;
      zj[c][r] = v;
      if (!last && gc < N) z[gc * R + r] = v;
    }
  }
  if (!last) return;
  __syncthreads();
  // a_j[c] = sum_r invD_j[r][c] * zj[r]
  const double* Dj = invD + (long long)j * NB * NB;
#pragma unroll
  for (int r = 0; r < RMAX; r++) acc[r] = 0.0;
  for (int row = half * 64; row < (half + 1) * 64; row++) {
    const double d = Dj[row * NB + c];
#pragma unroll
    for (int r = 0; r < RMAX; r++) if (r < R) acc[r] = fma(d, zj[row][r], acc[r]);
  }
  __syncthreads();
  for (int r = 0; r < R; r++) part[half][c][r] = acc[r];
  __syncthreads();
  if (tid < NB) {
    const long long gc = c0 + c;
    if (gc < N) for (int r = 0; r < R; r++) z[gc * R + r] = part[0][c][r] + part[1][c][r];
  }
}

// lml[r] = -0.5 * sum_i Y[i,r] alpha[i,r] - sum_i log L_ii - N/2 log(2 pi); one CTA per matrix.
__global__ void __launch_bounds__(1024)
lml_kernel(const double* __restrict__ L, long long ldl, long long N, const double* __restrict__ Y,
           const double* __restrict__ alpha, int R, double* __restrict__ lml, long long batch_l,
           long long batch_y) {
  __shared__ double red[32][RMAX + 1];
  L += blockIdx.x * batch_l; Y += blockIdx.x * batch_y; alpha += blockIdx.x * batch_y;
  lml += blockIdx.x * R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double acc[RMAX + 1];
#pragma unroll
  for (int r = 0; r <= RMAX; r++) acc[r] = 0.0;
  for (long long i = tid; i < N; i += 1024) {
    acc[RMAX] += log(L[i * ldl + i]);
#pragma unroll
    for (int r = 0; r < RMAX; r++) if (r < R) acc[r] = fma(Y[i * R + r], alpha[i * R + r], acc[r]);
  }
#pragma unroll
  for (int r = 0; r <= RMAX; r++) acc[r] = warp_sum(acc[r]);
  if (lane == 0) for (int r = 0; r <= RMAX; r++) red[warp][r] = acc[r];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int r = 0; r <= RMAX; r++) acc[r] = warp_sum(red[lane][r]);
    if (lane == 0) {
      const double c = 0.5 * (double)N * 1.8378770664093454835606594728112;   // log(2 pi)
      for (int r = 0; r < R; r++) lml[r] = -0.5 * acc[r] - acc[RMAX] - c;
    }
  }
}

// alpha (N x R per matrix) must already hold a copy of Y; solved in place.
int solve_blocked(const double* L, long long N, long long ldl, const double* invD, double* alpha, int R,
                  int batch, long long batch_l, long long batch_inv, long long batch_z,
                  cudaStream_t stream) {
  const int nblk = (int)((N + NB - 1) / NB);
  for (int k = -1; k <= nblk - 2; k++) {
    dim3 grid(k < 0 ? 1 : nblk - 1 - k, batch);
    fwd_step_kernel<<<grid, 256, 0, stream>>>(L, ldl, N, invD, alpha, R, k, batch_l, batch_inv, batch_z);
  }
  for (int k = nblk; k >= 1; k--) {
    dim3 grid(k == nblk ? 1 : k, batch);
    bwd_step_kernel<<<grid, 256, 0, stream>>>(L, ldl, N, invD, alpha, R, k, nblk, batch_l, batch_inv, batch_z);
  }
  count_launch(2 * nblk - 1);
  GPM_LAUNCH_CHECK();
  return 0;
}

int launch_lml(const double* L, long long N, long long ldl, const double* Y, const double* alpha, int R,
               double* lml, int batch, long long batch_l, long long batch_y, cudaStream_t stream) {
  lml_kernel<<<batch, 1024, 0, stream>>>(L, ldl, N, Y, alpha, R, lml, batch_l, batch_y);
  GPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace gpm

using namespace gpm;

extern "C" int gpm_solve_lml(gpm_handle_t h, const double* L, int64_t N, int64_t ldl, const void* potrf_ws,
                             const double* Y, int32_t R, double* alpha, double* lml, gpm_stream_t stream) {
  GPM_ARG(h != nullptr, 1);
  GPM_ARG(L != nullptr, 2);
  GPM_ARG(N > 0, 3);
  GPM_ARG(ldl >= N, 4);
  GPM_ARG(potrf_ws != nullptr, 5);
  GPM_ARG(Y != nullptr, 6);
  GPM_ARG(R >= 1 && R <= RMAX, 7);
  GPM_ARG(alpha != nullptr && alpha != Y, 8);
  cudaStream_t st = (cudaStream_t)stream;
  GPM_CUDA(cudaMemcpyAsync(alpha, Y, (size_t)N * R * sizeof(double), cudaMemcpyDeviceToDevice, st));
  int rc = solve_blocked(L, N, ldl, reinterpret_cast<const double*>(potrf_ws), alpha, R, 1, 0, 0, 0, st);
  if (rc) return rc;
  if (lml) return launch_lml(L, N, ldl, Y, alpha, R, lml, 1, 0, 0, st);
  return 0;
}
