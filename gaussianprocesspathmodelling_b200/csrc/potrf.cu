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

#include "gemm.cuh"

namespace gpm {

constexpr int PLD = NB + 1;   // padded smem pitch (doubles) of the 128x128 block
constexpr int XLD = 12;       // pitch of the 8-column panel staging buffer (conflict-free DMMA fragment reads)
constexpr int POTF2_SMEM = (NB * PLD + NB * XLD + 16 * 64) * 8;

// One 8x8 lower-triangular factor + its inverse, entirely in one thread's registers.
// a: packed lower (row-major, a[i*(i+1)/2 + j]); on exit a holds L, x holds inv(L) (same packing).
// Returns the 1-based index of the first non-positive pivot (0 if none).
__device__ __forceinline__ int chol8_inv(double (&a)[36], double (&x)[36]) {
  // every loop has constant bounds 0..7 with compile-time-foldable guards, so that full unrolling
  // keeps both arrays in registers
  double r[8];
  int bad = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    double d = a[j * (j + 1) / 2 + j];
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k < j) d = fma(-a[j * (j + 1) / 2 + k], a[j * (j + 1) / 2 + k], d);
    if (!(d > 0.0) || !(d < 1.0e300)) { if (!bad) bad = j + 1; d = 1.0; }
    const double ljj = sqrt(d);
    r[j] = 1.0 / ljj;
    a[j * (j + 1) / 2 + j] = ljj;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i > j) {
        double v = a[i * (i + 1) / 2 + j];
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k < j) v = fma(-a[i * (i + 1) / 2 + k], a[j * (j + 1) / 2 + k], v);
        a[i * (i + 1) / 2 + j] = v * r[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    x[j * (j + 1) / 2 + j] = r[j];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i > j) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k >= j && k < i) v = fma(a[i * (i + 1) / 2 + k], x[k * (k + 1) / 2 + j], v);
        x[i * (i + 1) / 2 + j] = -v * r[i];
      }
    }
  }
  return bad;
}

// One level of the recursive-doubling inverse on DMMA tiles:  X21 = -X22 * (L21 * X11)  for all
// 64/S pairs of SxS diagonal blocks (X11, X22 already inverted in place, upper parts zero).
template <int S>
__device__ __forceinline__ void inv_level_dmma(double* sm, int warp, int lane) {
  constexpr int TB = S / 8;                 // 8x8 tiles per block edge
  constexpr int TILES = (NB / (2 * S)) * TB * TB;
  constexpr int PER_WARP = (TILES + 7) / 8;
  const int g = lane >> 2, q = lane & 3;
  double c0[PER_WARP], c1[PER_WARP];
  // phase 1: T = L21 * X11   (X11 lower: contraction blocks kb >= b)
#pragma unroll
  for (int e = 0; e < PER_WARP; e++) {
    const int t = warp + 8 * e;
    c0[e] = c1[e] = 0.0;
    if (t < TILES) {
      const int pair = t / (TB * TB), a = (t / TB) % TB, b = t % TB;
      const int o1 = pair * 2 * S, o2 = o1 + S;
      for (int k0 = 8 * b; k0 < S; k0 += 4) {
        const double af = sm[(o2 + 8 * a + g) * PLD + o1 + k0 + q];
        const double bf = sm[(o1 + k0 + q) * PLD + o1 + 8 * b + g];
        dmma(c0[e], c1[e], af, bf);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < PER_WARP; e++) {
    const int t = warp + 8 * e;
    if (t < TILES) {
      const int pair = t / (TB * TB), a = (t / TB) % TB, b = t % TB;
      const int o1 = pair * 2 * S, o2 = o1 + S;
      double* dst = sm + (o2 + 8 * a + g) * PLD + o1 + 8 * b + 2 * q;
      dst[0] = c0[e]; dst[1] = c1[e];
    }
  }
  __syncthreads();
  // phase 2: X21 = -X22 * T   (X22 lower: contraction blocks kb <= a)
#pragma unroll
  for (int e = 0; e < PER_WARP; e++) {
    const int t = warp + 8 * e;
    c0[e] = c1[e] = 0.0;
    if (t < TILES) {
      const int pair = t / (TB * TB), a = (t / TB) % TB, b = t % TB;
      const int o1 = pair * 2 * S, o2 = o1 + S;
      for (int k0 = 0; k0 < 8 * a + 8; k0 += 4) {
        const double af = -sm[(o2 + 8 * a + g) * PLD + o2 + k0 + q];
        const double bf = sm[(o2 + k0 + q) * PLD + o1 + 8 * b + g];
        dmma(c0[e], c1[e], af, bf);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < PER_WARP; e++) {
    const int t = warp + 8 * e;
    if (t < TILES) {
      const int pair = t / (TB * TB), a = (t / TB) % TB, b = t % TB;
      const int o1 = pair * 2 * S, o2 = o1 + S;
      double* dst = sm + (o2 + 8 * a + g) * PLD + o1 + 8 * b + 2 * q;
      dst[0] = c0[e]; dst[1] = c1[e];
    }
  }
  __syncthreads();
}

// Factor diagonal block kblk of K in shared memory, write L_kk back and inv(L_kk) to invD.
// Rows/columns beyond N are padded with the identity.  blockIdx.x = batch index.
//
// Right-looking over sixteen 8-column panels: (1) one thread factors and inverts the 8x8 diagonal
// block in registers, (2) one thread per row solves the panel against it, (3) all warps apply the
// rank-8 update to the trailing 8x8 tiles with two DMMA.8x8x4 each.  The 128x128 inverse is then
// assembled from the 8x8 inverses by recursive doubling, also on DMMA tiles.
__global__ void __launch_bounds__(256, 1)
potf2_inv_kernel(double* __restrict__ K, long long ldk, long long N, int kblk, double* __restrict__ invD,
                 int* __restrict__ info, long long batch_k, long long batch_inv) {
  extern __shared__ double sm[];
  double* xp = sm + NB * PLD;          // [128][XLD] current panel
  double* inv8 = xp + NB * XLD;        // [16][64]  inverses of the 8x8 diagonal blocks (full, zero upper)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  K += blockIdx.x * batch_k;
  invD += blockIdx.x * batch_inv + (long long)kblk * NB * NB;
  info += blockIdx.x;
  const long long r0 = (long long)kblk * NB;
  const int nv = (int)((N - r0) < NB ? (N - r0) : NB);

  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx >> 7, c = idx & 127;
    double v;
    if (i < nv && c < nv) v = (c <= i) ? K[(r0 + i) * ldk + r0 + c] : 0.0;
    else v = (i == c) ? 1.0 : 0.0;
    sm[i * PLD + c] = v;
  }
  __syncthreads();

  const int g = lane >> 2, q = lane & 3;
  for (int p = 0; p < 16; p++) {
    const int c0 = 8 * p;
    // (1) 8x8 diagonal block: factor + invert in one thread
    if (tid == 0) {
      double a[36], x[36];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) a[i * (i + 1) / 2 + j] = sm[(c0 + i) * PLD + c0 + j];
      const int bad = chol8_inv(a, x);
      if (bad && c0 + bad - 1 < nv) atomicCAS(info, 0, (int)(r0 + c0 + bad));
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (j <= i) sm[(c0 + i) * PLD + c0 + j] = a[i * (i + 1) / 2 + j];
          inv8[p * 64 + i * 8 + j] = (j <= i) ? x[i * (i + 1) / 2 + j] : 0.0;
        }
    }
    __syncthreads();
    // (2) panel solve: X[i, 0:8] = A[i, c0:c0+8] * inv(L8)^T, one thread per row below the block
    if (tid < NB && tid >= c0 + 8) {
      double a[8], x[8];
      double* row = sm + tid * PLD + c0;
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] = row[k];
#pragma unroll
      for (int c = 0; c < 8; c++) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k <= c; k++) v = fma(a[k], inv8[p * 64 + c * 8 + k], v);
        x[c] = v;
      }
#pragma unroll
      for (int c = 0; c < 8; c++) { row[c] = x[c]; xp[tid * XLD + c] = x[c]; }
    }
    __syncthreads();
    // (3) trailing update on 8x8 tiles: C[ti][tj] -= X_ti X_tj^T, two DMMAs per tile
    const int nt = 15 - p, rb = c0 + 8;
    const int ntiles = nt * (nt + 1) / 2;
    for (int e = warp; e < ntiles; e += 8) {
      int ti = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
      while ((ti + 1) * (ti + 2) / 2 <= e) ti++;
      while (ti * (ti + 1) / 2 > e) ti--;
      const int tj = e - ti * (ti + 1) / 2;
      const double* xa = xp + (rb + 8 * ti + g) * XLD + q;
      const double* xb = xp + (rb + 8 * tj + g) * XLD + q;
      double* cp = sm + (rb + 8 * ti + g) * PLD + rb + 8 * tj + 2 * q;
      double c0v = cp[0], c1v = cp[1];
      dmma(c0v, c1v, -xa[0], xb[0]);
      dmma(c0v, c1v, -xa[4], xb[4]);
      cp[0] = c0v; cp[1] = c1v;
    }
    __syncthreads();
  }

  // ---- write L_kk (lower part, valid rows); clear the strict upper triangle in shared memory ----
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx >> 7, c = idx & 127;
    if (c <= i) { if (i < nv) K[(r0 + i) * ldk + r0 + c] = sm[i * PLD + c]; }
    else sm[i * PLD + c] = 0.0;
  }
  __syncthreads();
  // ---- inverse: 8x8 diagonal inverses in place, then levels 8 -> 16 -> 32 -> 64 ----
  for (int idx = tid; idx < 16 * 64; idx += 256) {
    const int p = idx >> 6, i = (idx >> 3) & 7, j = idx & 7;
    sm[(8 * p + i) * PLD + 8 * p + j] = inv8[idx];
  }
  __syncthreads();
  inv_level_dmma<8>(sm, warp, lane);
  inv_level_dmma<16>(sm, warp, lane);
  inv_level_dmma<32>(sm, warp, lane);
  inv_level_dmma<64>(sm, warp, lane);

  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx >> 7, c = idx & 127;
    invD[idx] = sm[i * PLD + c];
  }
}

__global__ void zero_info_kernel(int* info, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) info[i] = 0;
}

int launch_potf2(double* K, long long ldk, long long N, int kblk, double* invD, int* info, int batch,
                 long long batch_k, long long batch_inv, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    GPM_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
    attr_set = true;
  }
  potf2_inv_kernel<<<batch, 256, POTF2_SMEM, stream>>>(K, ldk, N, kblk, invD, info, batch_k, batch_inv);
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
int potrf_blocked(gpm_handle_impl* h, double* K, long long N, long long ldk, double* invD, int* info,
                  int batch, long long batch_rows, cudaStream_t s0) {
  const int nblk = (int)((N + NB - 1) / NB);
  const long long total_rows = (batch - 1) * batch_rows + N;
  CUtensorMap mapK, mapInv;
  int rc = make_tmap(h, &mapK, K, total_rows, N, ldk, NB);
  if (rc) return rc;
  rc = make_tmap(h, &mapInv, invD, (long long)batch * nblk * NB, NB, NB, NB);
  if (rc) return rc;
  const long long batch_k = batch_rows * ldk, batch_inv = (long long)nblk * NB * NB;
  const bool lookahead = getenv("GPM_NO_LOOKAHEAD") == nullptr && nblk > 2;
  cudaStream_t s1 = lookahead ? h->aux : s0;
  rc = ensure_events(h, 2 * nblk + 2);
  if (rc) return rc;

  rc = launch_zero_info(info, batch, s0);
  if (rc) return rc;

  auto panel = [&](int k, cudaStream_t st) -> int {
    int r = launch_potf2(K, ldk, N, k, invD, info, batch, batch_k, batch_inv, st);
    if (r) return r;
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
    a.batch_a_rows = batch_rows; a.batch_b_rows = (long long)nblk * NB; a.batch_c_rows = batch_rows;
    return launch_gemm(h, mapK, mapInv, mapK, a, batch, st);
  };
  // trailing update of step k restricted to tile columns [jlo, jhi) (block indices), rows >= column
  auto update = [&](int k, int jlo, int jhi, cudaStream_t st) -> int {
    if (jlo >= jhi) return 0;
    GemmArgs a = {};
    a.C = K; a.ldc = ldk; a.rowsq = nullptr;
    a.a_col0 = k * NB; a.b_col0 = k * NB; a.b_tile_rows = NB; a.klen = NB;
    a.c_rows_end = N; a.c_cols_end = N;
    a.epi = EPI_SUB;
    a.batch_a_rows = batch_rows; a.batch_b_rows = batch_rows; a.batch_c_rows = batch_rows;
    if (jhi - jlo == 1) {                       // a single tile column: rows jlo .. nblk-1
      a.tri = 0; a.tiles_m = nblk - jlo; a.tiles_n = 1;
    } else {                                    // full lower triangle from block jlo on
      a.tri = 1; a.tiles_m = nblk - jlo; a.tiles_n = nblk - jlo;
    }
    a.a_row0 = jlo * NB; a.b_row0 = jlo * NB;
    a.c_row0 = (long long)jlo * NB; a.c_col0 = (long long)jlo * NB;
    a.max_tiles_per_cta = lookahead ? 8 : 16;
    return launch_gemm(h, mapK, mapK, mapK, a, batch, st);
  };

  if (!lookahead) {
    for (int k = 0; k < nblk; k++) {
      if ((rc = panel(k, s0))) return rc;
      if ((rc = update(k, k + 1, nblk, s0))) return rc;
    }
    return 0;
  }

  cudaEvent_t* ev_panel = h->ev;            // ev_panel[k]: panel k finished
  cudaEvent_t* ev_rest = h->ev + nblk + 1;  // ev_rest[k]:  rest-update of step k finished
  if ((rc = panel(0, s0))) return rc;
  GPM_CUDA(cudaEventRecord(ev_panel[0], s0));
  for (int k = 0; k + 1 < nblk; k++) {
    // helper stream: column k+1 of the step-k update, then panel k+1
    GPM_CUDA(cudaStreamWaitEvent(s1, ev_panel[k], 0));
    if (k > 0) GPM_CUDA(cudaStreamWaitEvent(s1, ev_rest[k - 1], 0));
    if ((rc = update(k, k + 1, k + 2, s1))) return rc;
    if ((rc = panel(k + 1, s1))) return rc;
    GPM_CUDA(cudaEventRecord(ev_panel[k + 1], s1));
    // caller's stream: the rest of the step-k update (tile columns >= k+2)
    GPM_CUDA(cudaStreamWaitEvent(s0, ev_panel[k], 0));
    if ((rc = update(k, k + 2, nblk, s0))) return rc;
    GPM_CUDA(cudaEventRecord(ev_rest[k], s0));
  }
  GPM_CUDA(cudaStreamWaitEvent(s0, ev_panel[nblk - 1], 0));
  return 0;
}

}  // namespace gpm

using namespace gpm;

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
  return potrf_blocked(h, K, N, ldk, reinterpret_cast<double*>(ws), info, 1, 0, (cudaStream_t)stream);
}
