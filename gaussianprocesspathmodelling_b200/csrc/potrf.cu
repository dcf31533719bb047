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

constexpr int PLD = NB + 1;   // padded smem pitch (doubles)
constexpr int POTF2_SMEM = (NB * PLD + NB) * 8;

// C(SxS block at [ro, co]) = sign * A(SxS at [ar, ac]) * B(SxS at [br, bc]), for 64/S independent
// "pairs" laid out along the diagonal with period 2S; results are returned in registers so the
// caller can overwrite one of the operands after a barrier.
template <int S, int TR, int TC>
__device__ __forceinline__ void smem_block_mm(const double* sm, int tid, int a_dr, int a_dc, int b_dr,
                                              int b_dc, double (&out)[TR][TC], int& row0, int& col0,
                                              int& pair_base) {
  constexpr int TX = S / TC, TY = S / TR, TPP = TX * TY;
  const int pair = tid / TPP, lt = tid % TPP;
  const int ty = lt / TX, tx = lt % TX;
  pair_base = pair * 2 * S;
  row0 = ty;   // rows  ty + r*TY
  col0 = tx;   // cols  tx + c*TX
#pragma unroll
  for (int r = 0; r < TR; r++)
#pragma unroll
    for (int c = 0; c < TC; c++) out[r][c] = 0.0;
  const double* A = sm + (pair_base + a_dr) * PLD + pair_base + a_dc;
  const double* B = sm + (pair_base + b_dr) * PLD + pair_base + b_dc;
#pragma unroll 4
  for (int k = 0; k < S; k++) {
    double a[TR], b[TC];
#pragma unroll
    for (int r = 0; r < TR; r++) a[r] = A[(ty + r * TY) * PLD + k];
#pragma unroll
    for (int c = 0; c < TC; c++) b[c] = B[k * PLD + tx + c * TX];
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
      for (int c = 0; c < TC; c++) out[r][c] = fma(a[r], b[c], out[r][c]);
  }
}

// one level of the recursive-doubling triangular inverse: X21 = -X22 * (L21 * X11)
template <int S, int TR, int TC>
__device__ __forceinline__ void inv_level(double* sm, int tid) {
  constexpr int TX = S / TC, TY = S / TR;
  double t[TR][TC];
  int r0, c0, pb;
  smem_block_mm<S, TR, TC>(sm, tid, /*A=L21*/ S, 0, /*B=X11*/ 0, 0, t, r0, c0, pb);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < TR; r++)
#pragma unroll
    for (int c = 0; c < TC; c++) sm[(pb + S + r0 + r * TY) * PLD + pb + c0 + c * TX] = t[r][c];
  __syncthreads();
  smem_block_mm<S, TR, TC>(sm, tid, /*A=X22*/ S, S, /*B=T*/ S, 0, t, r0, c0, pb);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < TR; r++)
#pragma unroll
    for (int c = 0; c < TC; c++) sm[(pb + S + r0 + r * TY) * PLD + pb + c0 + c * TX] = -t[r][c];
  __syncthreads();
}

// Factor diagonal block kblk of K in shared memory, write L_kk back and inv(L_kk) to invD.
// Rows/columns beyond N are padded with the identity.  blockIdx.x = batch index.
__global__ void __launch_bounds__(256, 1)
potf2_inv_kernel(double* __restrict__ K, long long ldk, long long N, int kblk, double* __restrict__ invD,
                 int* __restrict__ info, long long batch_k, long long batch_inv) {
  extern __shared__ double sm[];
  double* dg = sm + NB * PLD;
  const int tid = threadIdx.x;
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

  // ---- unblocked right-looking factorisation, 2 barriers per column; 16x16 cyclic thread grid ----
  const int ty = tid >> 4, tx = tid & 15;
  for (int j = 0; j < NB; j++) {
    const double ajj = sm[j * PLD + j];
    const bool bad = !(ajj > 0.0) || !(ajj < 1.0e300);
    const double d = bad ? 1.0 : sqrt(ajj);
    if (tid > j && tid < NB) sm[tid * PLD + j] /= d;
    if (tid == 0) {
      dg[j] = d;
      if (bad && j < nv) atomicCAS(info, 0, (int)(r0 + j + 1));
    }
    __syncthreads();
    const int jp = j + 1;
    const int i0 = ty >= jp ? ty : ty + (((jp - ty) + 15) & ~15);
    const int c0 = tx >= jp ? tx : tx + (((jp - tx) + 15) & ~15);
    for (int i = i0; i < NB; i += 16) {
      const double lij = sm[i * PLD + j];
      for (int c = c0; c <= i; c += 16) sm[i * PLD + c] = fma(-lij, sm[c * PLD + j], sm[i * PLD + c]);
    }
    __syncthreads();
  }
  if (tid < NB) sm[tid * PLD + tid] = dg[tid];
  __syncthreads();

  // ---- write L_kk (lower part, valid rows) ----
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx >> 7, c = idx & 127;
    if (i < nv && c <= i) K[(r0 + i) * ldk + r0 + c] = sm[i * PLD + c];
  }
  __syncthreads();

  // ---- inverse, level 0: the eight 16x16 diagonal sub-blocks, one thread per column ----
  {
    double x[16];
    const int bb = (tid >> 4) * 16, jj = tid & 15;
    if (tid < NB) {
      const double* Lb = sm + bb * PLD + bb;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        double s = (i == jj) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; k++) s = fma(-Lb[i * PLD + k], x[k], s);
        x[i] = (i < jj) ? 0.0 : s / Lb[i * PLD + i];
      }
    }
    __syncthreads();
    if (tid < NB) {
#pragma unroll
      for (int i = 0; i < 16; i++) sm[(bb + i) * PLD + bb + jj] = x[i];
    }
    __syncthreads();
  }
  // ---- levels 16 -> 32 -> 64: X21 = -X22 L21 X11 ----
  inv_level<16, 2, 2>(sm, tid);
  inv_level<32, 2, 4>(sm, tid);
  inv_level<64, 4, 4>(sm, tid);

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
    return launch_gemm(h, mapK, mapInv, a, batch, st);
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
    return launch_gemm(h, mapK, mapK, a, batch, st);
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
