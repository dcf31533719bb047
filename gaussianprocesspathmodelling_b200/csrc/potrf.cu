// Step 2: blocked right-looking Cholesky, K = L L^T in place (lower, row-major), NB = 128.
//
//   for each block column k:
//     potf2_inv : factor the 128x128 diagonal block in shared memory and invert it      (1 CTA)
//     panel     : L[i,k] = K[i,k] * inv(L_kk)^T  for i > k       (DMMA GEMM, in place)
//     trailing  : K[i,j] -= L[i,k] L[j,k]^T      for k < j <= i  (DMMA GEMM, lower tiles only)
//
// Look-ahead of depth 1: the trailing update of step k is split into block column k+1 (done
// first, on the handle's high-priority stream, followed immediately by panel k+1) and the rest
// (on the caller's stream), so the latency-bound diagonal factorisation of step k+1 overlaps the
// bulk of step k's tensor-core work.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "gemm.cuh"

#ifdef GPM_POTF2_TIMING
// phase stamps of CTA 0 (cycles): the branch on a value loaded from shared memory after the barrier keeps
// the clock read behind the barrier's completion (BAR.SYNC.DEFER_BLOCKING lets independent work issue early)
namespace gpm { __device__ long long g_p2_marks[128]; }
#define P2_MARK(slot)                                                                        \
  if (tid == 0 && blockIdx.x == 0) {                                                         \
    const double pv_ = *reinterpret_cast<volatile double*>(sm + gpm::PACKED + 63);           \
    if (__double_as_longlong(pv_) != 0x7ff8dead0000beefLL) gpm::g_p2_marks[slot] = clock64(); \
  }
// the look-ahead warp's lane 0 around the 8x8 factorisation of the next diagonal tile (panel p): slots 64+2p, 65+2p
#define P2_MARK_LA(p, e) if (blockIdx.x == 0) gpm::g_p2_marks[64 + 2 * (p) + (e)] = clock64();
#endif
#include "potf2.cuh"

namespace gpm {

// Factor diagonal block kblk of K in shared memory, write L_kk back and inv(L_kk) to invD.
// Rows/columns beyond N are padded with the identity.  blockIdx.x = batch index.
//
// Right-looking over sixteen 8-column panels: (1) one thread factors the 8x8 diagonal block in
// registers, (2) one thread per row forward-substitutes the panel against it, (3) all warps apply
// the rank-8 update to the trailing 8x8 tiles with DMMA.8x8x4 on a static balanced schedule.  The
// 128x128 inverse is then assembled from the 8x8 diagonal inverses by recursive doubling, also on
// DMMA tiles.  The kernel is latency-bound (serial pivot chain, barrier hand-offs), so it is sized
// for three CTAs per SM (71 KB shared memory, 256 threads): batched fits keep all of them busy.
template <int P2_THREADS>
__global__ void __launch_bounds__(P2_THREADS, P2_THREADS == 256 ? 3 : 1)
potf2_inv_kernel(double* __restrict__ K, long long ldk, long long N, int kblk, double* __restrict__ invD,
                 int* __restrict__ info, long long batch_k, long long batch_inv,
                 const double* rhs_r, double* rhs_z, int R, long long batch_rhs_rows,     // rhs_z may alias rhs_r (solve in place)
                 int diag_tiles_only) {                   // L_kk is scratch (batched fits): store only its diagonal 8 x 8 tiles
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  K += blockIdx.x * batch_k;
  invD += blockIdx.x * batch_inv + (long long)kblk * NB * NB;
  info += blockIdx.x;
  const long long r0 = (long long)kblk * NB;
  const int nv = (int)((N - r0) < NB ? (N - r0) : NB);

#ifdef GPM_POTF2_TIMING
  if (tid == 0 && blockIdx.x == 0) g_p2_marks[58] = clock64();
#endif
  // Block <-> global memory copies go tile by tile: a warp moves one 8 x 8 tile per instruction (lane = row g, column
  // pair 2q: 16 bytes), which is 512 contiguous bytes of the packed block -- conflict-free in shared memory (a warp
  // along a matrix row puts its eight tiles on the same 64 bytes of banks: twice the wavefronts) -- and eight 64-byte
  // row segments in global memory.  A pair never straddles a tile and keeps its order under the in-tile swizzle.
  constexpr int P2_WARPS = P2_THREADS / 32;
  const int warp = tid >> 5, lg = (tid & 31) >> 2, lq = (tid & 3) * 2;
  const int inoff = in_tile(lg, lq);
  // load the lower triangle (LU independent loads in flight per thread); identity padding beyond nv
  constexpr int LU = 9;                            // 136 tiles: 17 per warp with 8 warps (9 + 8), at most 9 with 16
  {
    int ti = 0, tj = warp;
    while (tj > ti) { tj -= ti + 1; ti++; }
    for (int t0 = warp; t0 < PACKED / 64; t0 += LU * P2_WARPS) {
      double2 v[LU];
      int ti_u = ti, tj_u = tj;
#pragma unroll
      for (int u = 0; u < LU; u++) {
        const int i = ti_u * 8 + lg, c = tj_u * 8 + lq;
        if (t0 + u * P2_WARPS < PACKED / 64) {
          if (i < nv && c <= i) v[u] = *reinterpret_cast<const double2*>(K + (r0 + i) * ldk + r0 + c);
          else v[u] = make_double2((i == c && i >= nv) ? 1.0 : 0.0, (i == c + 1 && i >= nv) ? 1.0 : 0.0);
        }
        tj_u += P2_WARPS;
        while (tj_u > ti_u) { tj_u -= ti_u + 1; ti_u++; }
      }
#pragma unroll
      for (int u = 0; u < LU; u++)
        if (t0 + u * P2_WARPS < PACKED / 64) *reinterpret_cast<double2*>(sm + (t0 + u * P2_WARPS) * 64 + inoff) = v[u];
      ti = ti_u; tj = tj_u;
    }
  }
  if (rhs_r != nullptr)     // r_k is final before this launch: stage it now, its latency hides behind the factorisation
    potf2_stage_rhs<P2_THREADS>(sm, tid, nv, rhs_r + (blockIdx.x * batch_rhs_rows + r0) * R, R);
  P2_MARK(0)
  __syncthreads();
  P2_MARK(1)

  potf2_factor<P2_THREADS, 0>(sm, tid, nv, r0, info);

  // ---- write L_kk (lower part, valid rows), tile by tile ----
  {
    int ti = 0, tj = warp;
    while (tj > ti) { tj -= ti + 1; ti++; }
    for (int t = warp; t < PACKED / 64; t += P2_WARPS) {
      const int i = ti * 8 + lg, c = tj * 8 + lq;
      if (i < nv && c <= i && (!diag_tiles_only || ti == tj)) {
        const double2 v = *reinterpret_cast<const double2*>(sm + t * 64 + inoff);
        double* dst = K + (r0 + i) * ldk + r0 + c;
        if (c < i) *reinterpret_cast<double2*>(dst) = v; else *dst = v.x;
      }
      tj += P2_WARPS;
      while (tj > ti) { tj -= ti + 1; ti++; }
    }
  }
  P2_MARK(50)
  potf2_invert<P2_THREADS, 0>(sm, tid);
  P2_MARK(55)
  if (rhs_r != nullptr)
    potf2_fwd_z<P2_THREADS, 0, true>(sm, tid, nv, rhs_r + (blockIdx.x * batch_rhs_rows + r0) * R,
                               rhs_z + (blockIdx.x * batch_rhs_rows + r0) * R, R);
  P2_MARK(57)
  // the inverse, all 16 x 16 tiles (zeros above the diagonal), tile by tile
#pragma unroll 4
#pragma unroll 4
  for (int t = warp; t < NT8 * NT8; t += P2_WARPS) {
    const int ti = t >> 4, tj = t & 15;
    const double2 v = tj <= ti ? *reinterpret_cast<const double2*>(sm + tile_base(ti, tj) + inoff) : make_double2(0.0, 0.0);
    *reinterpret_cast<double2*>(invD + (ti * 8 + lg) * NB + tj * 8 + lq) = v;
  }
  P2_MARK(56)
}

__global__ void zero_info_kernel(int* info, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) info[i] = 0;
}

int launch_potf2(gpm_handle_impl* h, double* K, long long ldk, long long N, int kblk, double* invD, int* info, int batch,
                 long long batch_k, long long batch_inv, cudaStream_t stream, const double* rhs_r = nullptr,
                 double* rhs_z = nullptr, int R = 0, long long batch_rhs_rows = 0) {
  if (!h->potf2_attr) {
    GPM_CUDA(cudaFuncSetAttribute(potf2_inv_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
    GPM_CUDA(cudaFuncSetAttribute(potf2_inv_kernel<256>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    GPM_CUDA(cudaFuncSetAttribute(potf2_inv_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
    h->potf2_attr = true;
  }
  // a lone block is latency-critical (16 warps); batches are throughput-bound (8 warps, two CTAs per SM)
  if (batch >= 64)
    potf2_inv_kernel<256><<<batch, 256, POTF2_SMEM, stream>>>(K, ldk, N, kblk, invD, info, batch_k, batch_inv, rhs_r, rhs_z, R, batch_rhs_rows, h->scratch_factor ? 1 : 0);
  else
    potf2_inv_kernel<512><<<batch, 512, POTF2_SMEM, stream>>>(K, ldk, N, kblk, invD, info, batch_k, batch_inv, rhs_r, rhs_z, R, batch_rhs_rows, h->scratch_factor ? 1 : 0);
  GPM_LAUNCH_CHECK();
  return 0;
}

int launch_zero_info(int* info, int n, cudaStream_t stream) {
  zero_info_kernel<<<(n + 255) / 256, 256, 0, stream>>>(info, n);
  GPM_LAUNCH_CHECK();
  return 0;
}

static int ensure_events(gpm_handle_impl* h, int n) {
  if (h->n_ev >= n) return 0;
  cudaEvent_t* ne = new cudaEvent_t[n];
  for (int i = 0; i < h->n_ev; i++) ne[i] = h->ev[i];
  for (int i = h->n_ev; i < n; i++) GPM_CUDA(cudaEventCreateWithFlags(&ne[i], cudaEventDisableTiming));
  delete[] h->ev;
  h->ev = ne;
  h->n_ev = n;
  return 0;
}

// Factor `batch` matrices stacked along the rows of K (batch_rows rows apart; batch = 1 for the
// single-matrix case).  invD: batch x nblk x NB x NB.
// Optional fused forward substitution: rhs_r (N x R per matrix, batch_rhs_rows rows apart) enters holding the
// right-hand side Y and is consumed as the running residual; rhs_z receives z = L^{-1} Y block by block.
int potrf_blocked(gpm_handle_impl* h, double* K, long long N, long long ldk, double* invD, int* info,
                  int batch, long long batch_rows, cudaStream_t s0, double* rhs_r, double* rhs_z, int R,
                  long long batch_rhs_rows) {
  const int nblk = (int)((N + NB - 1) / NB);
  const long long total_rows = (batch - 1) * batch_rows + N;
  CUtensorMap mapK, mapInv;
  int rc = make_tmap(h, &mapK, K, total_rows, N, ldk, NB);
  if (rc) return rc;
  rc = make_tmap(h, &mapInv, invD, (long long)batch * nblk * NB, NB, NB, NB);
  if (rc) return rc;
  const long long batch_k = batch_rows * ldk, batch_inv = (long long)nblk * NB * NB;
  const bool lookahead = !h->opt.no_lookahead && nblk > 2 && batch < 32;
  cudaStream_t s1 = lookahead ? h->aux : s0;
  rc = ensure_events(h, 3 * nblk + 3);
  if (rc) return rc;

  rc = launch_zero_info(info, batch, s0);
  if (rc) return rc;

  auto diag_block = [&](int k, cudaStream_t st) -> int {
    return launch_potf2(h, K, ldk, N, k, invD, info, batch, batch_k, batch_inv, st, rhs_r, rhs_z, R, batch_rhs_rows);
  };
  auto panel_solve = [&](int k, cudaStream_t st) -> int {
    const int t = nblk - k - 1;   // row blocks below the diagonal
    if (t <= 0) return 0;
    GemmArgs a = {};
    a.C = K; a.ldc = ldk; a.rowsq = nullptr;
    a.tiles_m = t; a.tiles_n = 1; a.tri = 0;
    a.a_row0 = (k + 1) * NB; a.a_col0 = k * NB;
    a.b_row0 = k * NB; a.b_col0 = 0; a.b_tile_rows = 0;
    a.klen = NB;
    a.c_row0 = (long long)(k + 1) * NB; a.c_col0 = (long long)k * NB;
    a.c_rows_end = N; a.c_cols_end = (long long)(k + 1) * NB;
    a.epi = EPI_STORE;
    a.tri_b = 1;                                   // inv(L_kk) is lower triangular
    a.small_A = K; a.small_lda = ldk; a.small_a_rows_end = total_rows;
    a.small_B = invD; a.small_ldb = NB; a.small_b_rows_end = (long long)batch * nblk * NB;
    if (rhs_r) {
      a.rhs_r = rhs_r; a.rhs_z = rhs_z; a.rhs_R = R;
      a.rhs_z_row0 = (long long)k * NB; a.rhs_r_row0 = (long long)(k + 1) * NB;
      a.rhs_rows_end = N; a.batch_rhs_rows = batch_rhs_rows;
    }
    a.batch_a_rows = batch_rows; a.batch_b_rows = (long long)nblk * NB; a.batch_c_rows = batch_rows;
    return launch_gemm(h, mapK, mapInv, mapK, a, batch, st);
  };
  auto panel = [&](int k, cudaStream_t st) -> int {
    const int r = diag_block(k, st);
    return r ? r : panel_solve(k, st);
  };
  // trailing update with block columns [k0, k0+kw) of L (contraction length kw*NB) restricted to tile
  // columns [jlo, jhi) (block indices), rows >= column
  // part (single tile column only): 0 = the whole column, 1 = its diagonal tile, 2 = the tiles below the diagonal
  auto update = [&](int k0, int kw, int jlo, int jhi, cudaStream_t st, int part = 0) -> int {
    if (jlo >= jhi || jlo >= nblk) return 0;
    if (part == 2 && jlo + 1 >= nblk) return 0;
    GemmArgs a = {};
    a.C = K; a.ldc = ldk; a.rowsq = nullptr;
    a.a_col0 = k0 * NB; a.b_col0 = k0 * NB; a.b_tile_rows = NB; a.klen = kw * NB;
    a.c_rows_end = N; a.c_cols_end = N;
    a.epi = EPI_SUB;
    a.diag_lower = 1;                           // tile (0,0) / tiles ti == tj are L L^T: lower triangle only
    a.small_A = K; a.small_B = K; a.small_lda = a.small_ldb = ldk; a.small_a_rows_end = a.small_b_rows_end = total_rows;
    a.batch_a_rows = batch_rows; a.batch_b_rows = batch_rows; a.batch_c_rows = batch_rows;
    if (jhi - jlo == 1) {                       // a single tile column: rows jlo .. nblk-1
      a.tri = 0; a.tiles_m = nblk - jlo; a.tiles_n = 1;
    } else {                                    // full lower triangle from block jlo on (jhi == nblk)
      a.tri = 1; a.tiles_m = nblk - jlo; a.tiles_n = nblk - jlo;
    }
    a.a_row0 = jlo * NB; a.b_row0 = jlo * NB;
    a.c_row0 = (long long)jlo * NB; a.c_col0 = (long long)jlo * NB;
    if (part == 1) {
      a.tiles_m = 1;
    } else if (part == 2) {
      a.tiles_m = nblk - jlo - 1; a.diag_lower = 0;
      a.a_row0 = (jlo + 1) * NB; a.c_row0 = (long long)(jlo + 1) * NB;
    }
    const int tpc_wide = h->opt.tpc_wide, tpc_narrow = h->opt.tpc_narrow;
    a.max_tiles_per_cta = lookahead ? (kw > 2 ? 2 : (kw > 1 ? tpc_wide : tpc_narrow)) : 16;   // keep CTAs short enough for the panel stream
    return launch_gemm(h, mapK, mapK, mapK, a, batch, st);
  };
  // factor the block columns of one outer panel [b0, b0+w): potf2 + panel solve per 128-column block;
  // inside the panel each block column is first updated with the panel's earlier blocks (left-looking)
  auto outer_panel = [&](int b0, int w, cudaStream_t st) -> int {
    int r;
    for (int j = 0; j < w; j++) {
      if (j > 0 && (r = update(b0, j, b0 + j, b0 + j + 1, st))) return r;
      if ((r = panel(b0 + j, st))) return r;
    }
    return 0;
  };

  // Outer panels: wide panels make the trailing updates deep (K = 256 or 512: less C traffic and the tile
  // prologue/epilogue amortised over more slabs) but lengthen the serial panel chain, which must stay
  // hidden behind the rest-update: width 8 while >= 96 block columns remain, 4 while >= 64, 2 while >= 36,
  // then 1 (thresholds swept on B200 at N = 8192 and 16384, tools/potrf_sweep.sh and potrf_sweep2.sh; the response
  // is flat within 1 % around these values; 36 rather than 32 keeps N = 4096 all-narrow, tools/potrf_sweep3.sh).
  // Large batches are throughput-bound in every launch, so they use width 4 (option batch_width) and no look-ahead.
  const int wide_env = h->opt.wide_min, wide4_env = h->opt.wide4_min, wide8_env = h->opt.wide8_min;
  std::vector<int> pb(nblk + 1), pw(nblk + 1);
  int npanel = 0;
  for (int b = 0; b < nblk;) {
    const int left = nblk - b;
    int w = 1;
    if (batch >= 32) w = std::max(1, std::min(left, h->opt.batch_width));
    else if (left >= wide8_env) w = 8;
    else if (left >= wide4_env) w = 4;
    else if (left >= wide_env) w = 2;
    w = std::min(w, left);
    pb[npanel] = b; pw[npanel] = w; npanel++;
    b += w;
  }

  if (!lookahead) {
    for (int P = 0; P < npanel; P++) {
      if ((rc = outer_panel(pb[P], pw[P], s0))) return rc;
      if ((rc = update(pb[P], pw[P], pb[P] + pw[P], nblk, s0))) return rc;
    }
    return 0;
  }

  cudaEvent_t* ev_panel = h->ev;            // ev_panel[P]: outer panel P finished
  cudaEvent_t* ev_rest = h->ev + nblk + 1;  // ev_rest[P]:  rest-update of outer step P finished
  cudaEvent_t* ev_col = h->ev + 2 * nblk + 2;  // ev_col[P]: the below-diagonal tiles of the next panel's column are updated
  if ((rc = outer_panel(pb[0], pw[0], s0))) return rc;
  GPM_CUDA(cudaEventRecord(ev_panel[0], s0));
  for (int P = 0; P + 1 < npanel; P++) {
    const int b0 = pb[P], w = pw[P], n0 = pb[P + 1], wn = pw[P + 1];
    // helper stream: the block columns of the next outer panel first, then that panel
    GPM_CUDA(cudaStreamWaitEvent(s1, ev_panel[P], 0));
    if (P > 0) GPM_CUDA(cudaStreamWaitEvent(s1, ev_rest[P - 1], 0));
    const bool split = wn == 1 && !h->opt.no_split_column && n0 + 1 < nblk && nblk >= 24;
    if (split) {
      // narrow panels of a mid-sized matrix (24 <= blocks < 36; measured -4.4 % at N = 4096, -5 % at 4480, +1 % at 2048): the next diagonal block needs only its own tile
      // of the column update, so that tile goes first and the diagonal-block kernel starts while the caller's stream
      // updates the rest of the column (ahead of the rest-update); the panel solve joins both
      if ((rc = update(b0, w, n0, n0 + 1, s1, 1))) return rc;
      if ((rc = diag_block(n0, s1))) return rc;
      GPM_CUDA(cudaStreamWaitEvent(s0, ev_panel[P], 0));
      if ((rc = update(b0, w, n0, n0 + 1, s0, 2))) return rc;
      GPM_CUDA(cudaEventRecord(ev_col[P], s0));
      GPM_CUDA(cudaStreamWaitEvent(s1, ev_col[P], 0));
      if ((rc = panel_solve(n0, s1))) return rc;
    } else {
      for (int c = 0; c < wn; c++)
        if ((rc = update(b0, w, n0 + c, n0 + c + 1, s1))) return rc;
      if ((rc = outer_panel(n0, wn, s1))) return rc;
      GPM_CUDA(cudaStreamWaitEvent(s0, ev_panel[P], 0));
    }
    GPM_CUDA(cudaEventRecord(ev_panel[P + 1], s1));
    // caller's stream: the rest of the step-P update (tile columns beyond the next panel)
    if ((rc = update(b0, w, n0 + wn, nblk, s0))) return rc;
    GPM_CUDA(cudaEventRecord(ev_rest[P], s0));
  }
  GPM_CUDA(cudaStreamWaitEvent(s0, ev_panel[npanel - 1], 0));
  return 0;
}

}  // namespace gpm

using namespace gpm;

#ifdef GPM_POTF2_TIMING
extern "C" int gpm_debug_potf2_marks(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_p2_marks, sizeof(long long) * 128);
}
#endif

extern "C" size_t gpm_potrf_workspace_bytes(int64_t N) {
  if (N <= 0) return 0;
  const int64_t nblk = (N + NB - 1) / NB;
  return (size_t)nblk * NB * NB * sizeof(double);
}

extern "C" int gpm_potrf(gpm_handle_t handle, double* K, int64_t N, int64_t ldk, void* ws, int32_t* info,
                         gpm_stream_t stream) {
  GPM_ARG(handle != nullptr, 1);
  GPM_ARG(K != nullptr && ((uintptr_t)K & 15) == 0, 2);
  GPM_ARG(N > 0 && N <= (1 << 20), 3);
  GPM_ARG(ldk >= N && (ldk & 1) == 0, 4);
  GPM_ARG(ws != nullptr && ((uintptr_t)ws & 15) == 0, 5);
  GPM_ARG(info != nullptr, 6);
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  DeviceGuard guard(h->device);
  return potrf_blocked(h, K, N, ldk, reinterpret_cast<double*>(ws), info, 1, 0, (cudaStream_t)stream, nullptr, nullptr, 0, 0);
}
