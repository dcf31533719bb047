// Step 3: alpha = L^{-T} (L^{-1} Y) by blocked substitution with the inverted diagonal blocks that
// gpm_potrf left in its workspace, then the log marginal likelihood.  HBM-bound: L is read twice.
//
// Forward  (k = -1 .. nblk-2): every row block i > k does  y_i -= L_ik z_k ; block i = k+1 then
//                              finishes  z_i = inv(L_ii) y_i.
// Backward (k = nblk .. 1):    every column block j < k does  z_j -= L_kj^T a_k ; block j = k-1
//                              then finishes  a_j = inv(L_jj)^T z_j.
#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------------
// Single-matrix path: one cooperative launch per direction.  CTA c owns block rows c, c+G, ... and
// sweeps the 128x128 tiles of its block row (forward) / block column (backward) as the solution
// blocks it depends on are published by their owners (release/acquire flags in global memory, cleared
// on the stream before each solve so that a captured CUDA graph can be replayed).  L is streamed from HBM once
// per direction with 128 KB in flight per CTA; the critical path is one flag hand-off per block.
// ------------------------------------------------------------------------------------------------
// acc[0..RR) += s * zrow[0..RR)  with 128-bit shared loads (zrow 16-byte aligned when RR is even)
template <int RR>
__device__ __forceinline__ void fma_row(double (&acc)[RR], double s, const double* zrow) {
  if constexpr (RR == 1) {
    acc[0] = fma(s, zrow[0], acc[0]);
  } else {
#pragma unroll
    for (int r2 = 0; r2 < RR / 2; r2++) {
      const double2 v = *reinterpret_cast<const double2*>(zrow + 2 * r2);
      acc[2 * r2] = fma(s, v.x, acc[2 * r2]);
      acc[2 * r2 + 1] = fma(s, v.y, acc[2 * r2 + 1]);
    }
  }
}

constexpr int CH_THREADS = 512;
constexpr int ILD = NB + 1;
constexpr int CHAIN_SMEM = (NB * ILD + NB * RMAX + 4 * NB * RMAX) * 8;

__device__ __forceinline__ void flag_wait(const int* f, int epoch) {
  // spin with relaxed loads (an acquire load invalidates L1 on every iteration: CCTL.IVALL), then
  // one acquire fence once the flag is seen
  int v;
  do {
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  } while (v != epoch);
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void flag_set(int* f, int epoch) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}

#ifdef GPM_SOLVE_TIMING
__device__ unsigned long long g_solve_ts[10 * 8192];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ST_MARK(slot, blk) if (!BWD && threadIdx.x == 0) g_solve_ts[(slot) * 8192 + (blk)] = gtime();
#else
#define ST_MARK(slot, blk)
#endif

template <bool BWD, int RR>
__global__ void __launch_bounds__(CH_THREADS, 1)
solve_chain_kernel(const double* __restrict__ L, long long ldl, long long N, const double* __restrict__ invD,
                   double* z, int R, int nblk, int* flags, int epoch,
                   const double* __restrict__ Yl, double* lml_part, int* lml_flags, double* lml_out) {
  extern __shared__ double csm[];
  double* sinv = csm;                       // [128][ILD]  inverse of this block's diagonal block
  double* zs = sinv + NB * ILD;             // [128][RR] the dependency block just received
  double* part = zs + NB * RR;              // [4][128][RR] partial sums of the four quarters
  const int tid = threadIdx.x;
  const int e = tid & 127, qd = tid >> 7;   // element (row fwd / column bwd) and quarter of the tile
  const int G = gridDim.x;

  for (int it = blockIdx.x; it < nblk; it += G) {
    const int i = BWD ? nblk - 1 - it : it;                 // block owned in this iteration
    const long long i0 = (long long)i * NB;
    // stage inv(L_ii) (off the critical path)
    const double* Di = invD + (long long)i * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += CH_THREADS) sinv[(idx >> 7) * ILD + (idx & 127)] = Di[idx];
    double acc[RR], rhs[RR];
#pragma unroll
    for (int r = 0; r < RR; r++) {
      acc[r] = 0.0;
      rhs[r] = (tid < NB && i0 + tid < N && r < R) ? z[(i0 + tid) * R + r] : 0.0;   // own right-hand side, off the chain
    }

    const int ndep = BWD ? nblk - 1 - i : i;
    for (int dd = 0; dd < ndep; dd++) {
      const int j = BWD ? nblk - 1 - dd : dd;               // dependency block, in publication order
      const long long j0 = (long long)j * NB;
      // issue this tile's loads before waiting for the dependency: they do not depend on it
      double seg[32];
      if (!BWD) {                                            // row i0+e, columns j0 + 32 qd .. +32
        const long long gr = i0 + e;
        const double2* src = reinterpret_cast<const double2*>(L + gr * ldl + j0 + 32 * qd);
#pragma unroll
        for (int c = 0; c < 16; c++) {
          double2 v = (gr < N) ? __ldcs(src + c) : make_double2(0.0, 0.0);
          seg[2 * c] = v.x; seg[2 * c + 1] = v.y;
        }
      } else {                                               // column i0+e, rows j0 + 32 qd .. +32
#pragma unroll
        for (int r = 0; r < 32; r++) {
          const long long gr = j0 + 32 * qd + r;
          seg[r] = (gr < N) ? __ldcs(L + gr * ldl + i0 + e) : 0.0;
        }
      }
      if (dd == ndep - 1) { ST_MARK(0, i) }       // ready for the last dependency
      if (tid == 0) flag_wait(flags + j, epoch);
      if (dd == ndep - 1) { ST_MARK(1, i) }       // last dependency seen
      __syncthreads();
      for (int idx = tid; idx < NB * R; idx += CH_THREADS) {
        const long long gr = j0 + idx / R;
        zs[(idx / R) * RR + idx % R] = (gr < N) ? __ldcg(z + j0 * R + idx) : 0.0;
      }
      __syncthreads();
      if (dd == ndep - 1) { ST_MARK(3, i) }       // z_j staged
#pragma unroll
      for (int c = 0; c < 32; c++) fma_row<RR>(acc, seg[c], zs + (32 * qd + c) * RR);
    }
    // combine the quarters: y = rhs - sum
    __syncthreads();
    ST_MARK(4, i)                                  // tile product done
#pragma unroll
    for (int r = 0; r < RR; r++) part[(qd * NB + e) * RR + r] = acc[r];
    __syncthreads();
    if (tid < NB) {
#pragma unroll
      for (int r = 0; r < RR; r++) {
        const double s = (part[(0 * NB + tid) * RR + r] + part[(1 * NB + tid) * RR + r]) +
                         (part[(2 * NB + tid) * RR + r] + part[(3 * NB + tid) * RR + r]);
        zs[tid * RR + r] = rhs[r] - s;
      }
    }
    __syncthreads();
    ST_MARK(5, i)                                  // y formed
    // multiply by inv(L_ii) (forward) or inv(L_ii)^T (backward) from shared memory
#pragma unroll
    for (int r = 0; r < RR; r++) acc[r] = 0.0;
#pragma unroll 8
    for (int c = 0; c < 32; c++) {
      const int k = 32 * qd + c;
      const double d = BWD ? sinv[k * ILD + e] : sinv[e * ILD + k];
      fma_row<RR>(acc, d, zs + k * RR);
    }
#pragma unroll
    for (int r = 0; r < RR; r++) part[(qd * NB + e) * RR + r] = acc[r];
    __syncthreads();
    ST_MARK(6, i)                                  // inverse-block product done
    double aval[RR];
#pragma unroll
    for (int r = 0; r < RR; r++) aval[r] = 0.0;
    if (tid < NB) {
      const long long gr = i0 + tid;
      if (gr < N) {
#pragma unroll
        for (int r = 0; r < RR; r++) {
          if (r < R) {
            aval[r] = (part[(0 * NB + tid) * RR + r] + part[(1 * NB + tid) * RR + r]) +
                      (part[(2 * NB + tid) * RR + r] + part[(3 * NB + tid) * RR + r]);
            z[gr * R + r] = aval[r];
          }
        }
      }
    }
    __syncthreads();            // orders the block's stores before thread 0's release (cumulativity)
    ST_MARK(7, i)                                  // z stored
    if (tid == 0) { __threadfence(); flag_set(flags + i, epoch); }
    ST_MARK(2, i)                                 // published
    if (BWD && lml_out != nullptr) {
      // Log marginal likelihood folded into the backward pass (off the chain: after the block is published).  Each
      // block leaves its share  sum_rows Y alpha  and  sum_rows log L_ii = -sum log inv(L_ii)_ii  in lml_part and
      // raises a second flag; the owner of block 0 -- the last block of the pass -- adds the shares in block order.
      double v[RR + 1];
#pragma unroll
      for (int r = 0; r <= RR; r++) v[r] = 0.0;
      if (tid < NB && i0 + tid < N) {
#pragma unroll
        for (int r = 0; r < RR; r++)
          if (r < R) v[r] = Yl[(i0 + tid) * R + r] * aval[r];
        v[RR] = -log(sinv[tid * ILD + tid]);
      }
      if (tid < NB) {
#pragma unroll
        for (int r = 0; r <= RR; r++) {
          const double sv = warp_sum(v[r]);
          if ((tid & 31) == 0) part[(tid >> 5) * (RR + 1) + r] = sv;
        }
      }
      __syncthreads();
      if (tid <= RR) {
        const double sv = (part[0 * (RR + 1) + tid] + part[1 * (RR + 1) + tid]) + (part[2 * (RR + 1) + tid] + part[3 * (RR + 1) + tid]);
        lml_part[(long long)i * (RMAX + 1) + (tid == RR ? RMAX : tid)] = sv;
      }
      __syncthreads();
      if (tid == 0) { __threadfence(); flag_set(lml_flags + i, epoch); }
      if (i == 0) {
        for (int j = 1 + tid; j < nblk; j += CH_THREADS) flag_wait(lml_flags + j, epoch);
        __syncthreads();
        // thread t adds the shares of blocks t, t + 128, ... in increasing order; then a fixed-order combine
        if (tid < NB) {
#pragma unroll
          for (int r = 0; r <= RR; r++) v[r] = 0.0;
          for (int j = tid; j < nblk; j += NB) {
#pragma unroll
            for (int r = 0; r <= RR; r++) v[r] += __ldcg(lml_part + (long long)j * (RMAX + 1) + (r == RR ? RMAX : r));
          }
#pragma unroll
          for (int r = 0; r <= RR; r++) {
            const double sv = warp_sum(v[r]);
            if ((tid & 31) == 0) part[(tid >> 5) * (RR + 1) + r] = sv;
          }
        }
        __syncthreads();
        if (tid < R) {
          const double qf = (part[0 * (RR + 1) + tid] + part[1 * (RR + 1) + tid]) + (part[2 * (RR + 1) + tid] + part[3 * (RR + 1) + tid]);
          const double ld = (part[0 * (RR + 1) + RR] + part[1 * (RR + 1) + RR]) + (part[2 * (RR + 1) + RR] + part[3 * (RR + 1) + RR]);
          lml_out[tid] = -0.5 * qf - ld - 0.5 * (double)N * 1.8378770664093454835606594728112;   // log(2 pi)
        }
      }
      __syncthreads();          // part / sinv are rewritten by the next block of this CTA
    }
  }
}

#ifdef GPM_SOLVE_TIMING
extern "C" void gpm_debug_solve_ts(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_solve_ts, sizeof(g_solve_ts)); }
#endif

// first_dir = 1: alpha already holds z = L^{-1} Y (forward substitution fused into the factorisation): backward pass only
// Y / lml (optional): the log marginal likelihood comes out of the backward pass itself (no pass of its own over L, Y, alpha)
int solve_chain(gpm_handle_impl* h, const double* L, long long N, long long ldl, const double* invD,
                double* alpha, int R, cudaStream_t stream, int first_dir, const double* Y, double* lml) {
  const void* fns[2];
  if (R <= 1) { fns[0] = (const void*)solve_chain_kernel<false, 1>; fns[1] = (const void*)solve_chain_kernel<true, 1>; }
  else if (R <= 2) { fns[0] = (const void*)solve_chain_kernel<false, 2>; fns[1] = (const void*)solve_chain_kernel<true, 2>; }
  else if (R <= 4) { fns[0] = (const void*)solve_chain_kernel<false, 4>; fns[1] = (const void*)solve_chain_kernel<true, 4>; }
  else { fns[0] = (const void*)solve_chain_kernel<false, 8>; fns[1] = (const void*)solve_chain_kernel<true, 8>; }
  for (int d = 0; d < 2; d++)
    GPM_CUDA(cudaFuncSetAttribute(fns[d], cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM));
  int nblk = (int)((N + NB - 1) / NB);
  if (nblk > h->n_flags) { set_error("solve: N too large for the handle's flag arrays"); return 997; }
  int grid = nblk < h->sm_count ? nblk : h->sm_count;
  // flags are cleared on the stream before each direction (constant epoch), so the call sequence can be
  // captured into a CUDA graph and replayed
  GPM_CUDA(cudaMemsetAsync(h->flags, 0, 3 * (size_t)h->n_flags * sizeof(int), stream));
  for (int dir = first_dir; dir < 2; dir++) {
    int epoch = 1;
    int* flags = h->flags + dir * h->n_flags;
    int* lml_flags = h->flags + 2 * h->n_flags;
    double* lml_part = h->lml_part;
    double* lml_out = dir == 1 ? lml : nullptr;
    void* args[] = {(void*)&L, (void*)&ldl, (void*)&N, (void*)&invD, (void*)&alpha, (void*)&R, (void*)&nblk,
                    (void*)&flags, (void*)&epoch, (void*)&Y, (void*)&lml_part, (void*)&lml_flags, (void*)&lml_out};
    GPM_CUDA(cudaLaunchCooperativeKernel(fns[dir], dim3(grid), dim3(CH_THREADS), args, CHAIN_SMEM, stream));
    count_launch(1);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Batched path: one CTA per path runs forward + backward substitution and the LML over its own
// factor; the whole right-hand side lives in shared memory, L and the inverted diagonal blocks are
// streamed once per direction (128 KB in flight per CTA).  No inter-CTA dependencies.
// ------------------------------------------------------------------------------------------------
constexpr int SB_MAXN = 2048;
bool solve_paths_supported(long long N) { return (N + NB - 1) / NB * NB <= SB_MAXN; }

template <int RR>
__global__ void __launch_bounds__(CH_THREADS, 1)
solve_path_kernel(const double* __restrict__ Lb, long long ldl, long long N, const double* __restrict__ invDb,
                  const double* __restrict__ Yb, double* __restrict__ alphab, double* __restrict__ lmlb,
                  int R, int nblk, long long batch_l, long long batch_inv, long long batch_y,
                  const double* __restrict__ zfwd) {
  extern __shared__ double psm[];
  double* zs = psm;                                   // [nblk*128][RR]
  double* part = zs + (long long)nblk * NB * RR;      // [4][128][RR]
  double* ys = part + 4 * NB * RR;                    // [128][RR]
  __shared__ double red[CH_THREADS / 32][RMAX + 1];
  const double* L = Lb + blockIdx.x * batch_l;
  const double* invD = invDb + blockIdx.x * batch_inv;
  const double* Y = Yb + blockIdx.x * batch_y;
  double* alpha = alphab + blockIdx.x * batch_y;
  const int tid = threadIdx.x, e = tid & 127, qd = tid >> 7, warp = tid >> 5, lane = tid & 31;

  // zfwd != nullptr: the forward substitution was fused into the factorisation; start from z = L^{-1} Y
  const double* Z0 = zfwd ? zfwd + blockIdx.x * batch_y : Y;
  for (int idx = tid; idx < nblk * NB * RR; idx += CH_THREADS) {
    const int row = idx / RR, r = idx % RR;
    zs[idx] = (row < N && r < R) ? Z0[(long long)row * R + r] : 0.0;
  }
  double logdet = 0.0;
  for (long long i = tid; i < N; i += CH_THREADS) logdet += log(L[i * ldl + i]);
  __syncthreads();

  for (int dir = zfwd ? 1 : 0; dir < 2; dir++) {
    for (int it = 0; it < nblk; it++) {
      const int i = dir ? nblk - 1 - it : it;
      const long long i0 = (long long)i * NB;
      double acc[RR];
#pragma unroll
      for (int r = 0; r < RR; r++) acc[r] = 0.0;
      const int ndep = dir ? nblk - 1 - i : i;
      for (int dd = 0; dd < ndep; dd++) {
        const int j = dir ? nblk - 1 - dd : dd;
        const long long j0 = (long long)j * NB;
        double seg[32];
        if (!dir) {
          const long long gr = i0 + e;
          const double2* src = reinterpret_cast<const double2*>(L + gr * ldl + j0 + 32 * qd);
#pragma unroll
          for (int c = 0; c < 16; c++) {
            double2 v = (gr < N) ? __ldcs(src + c) : make_double2(0.0, 0.0);
            seg[2 * c] = v.x; seg[2 * c + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int r = 0; r < 32; r++) {
            const long long gr = j0 + 32 * qd + r;
            seg[r] = (gr < N) ? __ldcs(L + gr * ldl + i0 + e) : 0.0;
          }
        }
        const double* zj = zs + (j0 + 32 * qd) * RR;
#pragma unroll
        for (int c = 0; c < 32; c++) fma_row<RR>(acc, seg[c], zj + c * RR);
      }
      // diagonal-block inverse segment (issued before the reduction barriers)
      double dseg[32];
      {
        const double* Di = invD + (long long)i * NB * NB;
        // inv(L_ii) is lower triangular with exact zeros above the diagonal: a warp (32 consecutive e) whose whole
        // 32 x 32 segment lies above it skips the loads (6 of the 16 segments: 48 KB of the block's 128 KB)
        if (!dir) {
          const double2* src = reinterpret_cast<const double2*>(Di + e * NB + 32 * qd);
          const bool live = qd <= (e >> 5);
#pragma unroll
          for (int c = 0; c < 16; c++) {
            double2 v = live ? __ldcs(src + c) : make_double2(0.0, 0.0);
            dseg[2 * c] = v.x; dseg[2 * c + 1] = v.y;
          }
        } else {
          const bool live = qd >= (e >> 5);
#pragma unroll
          for (int r = 0; r < 32; r++) dseg[r] = live ? __ldcs(Di + (32 * qd + r) * NB + e) : 0.0;
        }
      }
#pragma unroll
      for (int r = 0; r < RR; r++) part[(qd * NB + e) * RR + r] = acc[r];
      __syncthreads();
      if (tid < NB) {
#pragma unroll
        for (int r = 0; r < RR; r++) {
          const double sgm = (part[(0 * NB + tid) * RR + r] + part[(1 * NB + tid) * RR + r]) +
                             (part[(2 * NB + tid) * RR + r] + part[(3 * NB + tid) * RR + r]);
          ys[tid * RR + r] = zs[(i0 + tid) * RR + r] - sgm;
        }
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < RR; r++) acc[r] = 0.0;
#pragma unroll
      for (int c = 0; c < 32; c++) fma_row<RR>(acc, dseg[c], ys + (32 * qd + c) * RR);
#pragma unroll
      for (int r = 0; r < RR; r++) part[(qd * NB + e) * RR + r] = acc[r];
      __syncthreads();
      if (tid < NB) {
#pragma unroll
        for (int r = 0; r < RR; r++) {
          const double v = (part[(0 * NB + tid) * RR + r] + part[(1 * NB + tid) * RR + r]) +
                           (part[(2 * NB + tid) * RR + r] + part[(3 * NB + tid) * RR + r]);
          zs[(i0 + tid) * RR + r] = (i0 + tid < N) ? v : 0.0;
        }
      }
      __syncthreads();
    }
  }
  // alpha out + LML
  double acc[RMAX + 1];
#pragma unroll
  for (int r = 0; r < RMAX; r++) acc[r] = 0.0;
  acc[RMAX] = logdet;
  for (long long i = tid; i < N; i += CH_THREADS) {
#pragma unroll
    for (int r = 0; r < RR; r++) {
      if (r < R) {
        const double a = zs[i * RR + r];
        alpha[i * R + r] = a;
        acc[r] = fma(Y[i * R + r], a, acc[r]);
      }
    }
  }
  if (lmlb == nullptr) return;
#pragma unroll
  for (int r = 0; r <= RMAX; r++) acc[r] = warp_sum(acc[r]);
  if (lane == 0) for (int r = 0; r <= RMAX; r++) red[warp][r] = acc[r];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int r = 0; r <= RMAX; r++) acc[r] = warp_sum(lane < CH_THREADS / 32 ? red[lane][r] : 0.0);
    if (lane == 0) {
      const double c = 0.5 * (double)N * 1.8378770664093454835606594728112;
      for (int r = 0; r < R; r++) lmlb[blockIdx.x * R + r] = -0.5 * acc[r] - acc[RMAX] - c;
    }
  }
}

// returns -1 when the shape is not supported by the one-CTA-per-path kernel
int solve_paths(const double* L, long long N, long long ldl, const double* invD, const double* Y,
                double* alpha, double* lml, int R, int batch, long long batch_l, long long batch_inv,
                long long batch_y, cudaStream_t stream, const double* zfwd) {
  const int nblk = (int)((N + NB - 1) / NB);
  if ((long long)nblk * NB > SB_MAXN) return -1;
  const int RR = R <= 1 ? 1 : (R <= 2 ? 2 : (R <= 4 ? 4 : 8));
  const int smem = (nblk * NB * RR + 4 * NB * RR + NB * RR) * 8;
  auto kern = RR == 1 ? solve_path_kernel<1> : (RR == 2 ? solve_path_kernel<2> : (RR == 4 ? solve_path_kernel<4> : solve_path_kernel<8>));
  GPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<batch, CH_THREADS, smem, stream>>>(L, ldl, N, invD, Y, alpha, lml, R, nblk, batch_l, batch_inv, batch_y, zfwd);
  GPM_LAUNCH_CHECK();
  return 0;
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
  GPM_ARG(L != nullptr && ((uintptr_t)L & 15) == 0, 2);
  GPM_ARG(ldl >= N && (ldl & 1) == 0, 4);
  GPM_ARG(potrf_ws != nullptr, 5);
  GPM_ARG(Y != nullptr, 6);
  GPM_ARG(R >= 1 && R <= RMAX, 7);
  GPM_ARG(alpha != nullptr && alpha != Y, 8);
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard(reinterpret_cast<gpm_handle_impl*>(h)->device);
  GPM_CUDA(cudaMemcpyAsync(alpha, Y, (size_t)N * R * sizeof(double), cudaMemcpyDeviceToDevice, st));
  gpm_handle_impl* hi = reinterpret_cast<gpm_handle_impl*>(h);
  int rc = hi->opt.solve_steps
               ? solve_blocked(L, N, ldl, reinterpret_cast<const double*>(potrf_ws), alpha, R, 1, 0, 0, 0, st)
               : solve_chain(hi, L, N, ldl, reinterpret_cast<const double*>(potrf_ws), alpha, R, st, 0, Y, lml);
  if (rc) return rc;
  if (lml && hi->opt.solve_steps) return launch_lml(L, N, ldl, Y, alpha, R, lml, 1, 0, 0, st);
  return 0;
}
