// Latency variant of the tile GEMM for launches that are a fraction of a wave: the factorisation's critical path
// (panel solve and column update of the next block column, every launch of a fit with N <= 4096, the tail of a large
// one) is a chain of launches with at most a few dozen 128 x 128 tiles, where one tile per SM leaves most SMs idle
// for the 15-20 us a tile takes.  Here every 128 x 128 tile is split into four 64 x 64 quarters on four SMs.
//
// The arithmetic is that of gemm_nt_kernel to the bit: the same DMMA.8x8x4 instruction sequence per output element
// (16-deep slabs in increasing order, the four k4 steps of a slab with the permuted contraction order
// kidx(t, q) = 2t + 8(q >> 1) + (q & 1)), so it does not matter numerically which of the two kernels a launch takes.
// Operands are staged with 16-byte cp.async copies (no TMA: the boxes would be a quarter of the swizzle atom's
// rows and this kernel is about latency, not bandwidth), double-buffered over 32-deep chunks of the contraction.
#include <stdlib.h>

#include "gemm.cuh"

namespace gpm {

constexpr int ST = 64;                   // small tile edge
constexpr int KC = 32;                   // contraction chunk: two slabs of 16
constexpr int PITCH = KC + 2;            // doubles per staged row: 272 B keeps the fragment loads of a half-warp conflict-free
constexpr int SMALL_THREADS = 256;       // 8 warps as 2 (m) x 4 (n), warp tile 32 x 16
constexpr int SMALL_SMEM = 2 /*stages*/ * 2 /*operands*/ * ST * PITCH * 8;
constexpr int SMALL_MAX_TILES = 37;      // 4 quarters each: at most one wave of 148 SMs

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

__global__ void __launch_bounds__(SMALL_THREADS)
gemm_nt_small_kernel(const GemmArgs p) {
  extern __shared__ __align__(16) double smem_d[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = blockIdx.x >> 2, qi = (blockIdx.x >> 1) & 1, qj = blockIdx.x & 1;
  int ti, tj;
  if (p.tri) {
    int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) i++;
    while (i * (i + 1) / 2 > t) i--;
    ti = i; tj = t - i * (i + 1) / 2;
    if (ti == tj && qi == 0 && qj == 1) return;          // strictly above the diagonal: never read
  } else {
    ti = t % p.tiles_m; tj = t / p.tiles_m;
  }
  const long long a_row = (long long)p.a_row0 + ti * NB + qi * ST;
  const long long b_row = (long long)p.b_row0 + tj * p.b_tile_rows + qj * ST;
  const long long c_row = p.c_row0 + (long long)ti * NB + qi * ST;
  const long long c_col = p.c_col0 + (long long)tj * NB + qj * ST;
  if (c_row >= p.c_rows_end || c_col >= p.c_cols_end) return;
  // contraction range; a lower-triangular B block (inverted diagonal block) is zero right of its diagonal
  int klen = p.klen;
  if (p.tri_b) klen = min(klen, qj * ST + ST);
  const int nchunk = klen / KC;

  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, q = lane & 3;
  const bool sub = p.epi == EPI_SUB;
  // C fragment (EPI_SUB) fetched before the main loop so that its latency hides behind it
  double cf[4][2][2];
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
      cf[mt][nt][0] = cf[mt][nt][1] = 0.0;
      const long long row = c_row + wm * 32 + mt * 8 + g, col = c_col + wn * 16 + nt * 8 + 2 * q;
      if (sub && row < p.c_rows_end) {
        const double* src = p.C + row * p.ldc + col;
        if (col + 1 < p.c_cols_end) { const double2 v = *reinterpret_cast<const double2*>(src); cf[mt][nt][0] = v.x; cf[mt][nt][1] = v.y; }
        else if (col < p.c_cols_end) cf[mt][nt][0] = *src;
      }
    }

  // staging: per chunk 64 rows x 32 doubles per operand = 2 x 1024 16-byte pieces, 8 per thread.  Rows past the
  // end of the matrix are clamped (their results are never stored).
  const long long a_rows_end = p.small_a_rows_end, b_rows_end = p.small_b_rows_end;
  auto stage = [&](int chunk, int buf) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int piece = tid + SMALL_THREADS * u;          // 0..2047
      const int op = piece >> 10, r = (piece >> 4) & 63, c16 = piece & 15;
      const long long grow = op ? min(b_row + r, b_rows_end - 1) : min(a_row + r, a_rows_end - 1);
      const double* src = op ? p.small_B + grow * p.small_ldb + p.b_col0 + chunk * KC + c16 * 2
                             : p.small_A + grow * p.small_lda + p.a_col0 + chunk * KC + c16 * 2;
      const uint32_t dst = smem_u32(smem_d + ((buf * 2 + op) * ST + r) * PITCH + c16 * 2);
      cp_async16(dst, src);
    }
    cp_async_commit();
  };

  double acc[4][2][2];
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  if (nchunk > 0) stage(0, 0);
  for (int c = 0; c < nchunk; c++) {
    if (c + 1 < nchunk) { stage(c + 1, (c + 1) & 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const double* As = smem_d + ((c & 1) * 2 + 0) * ST * PITCH + (wm * 32 + g) * PITCH;
    const double* Bs = smem_d + ((c & 1) * 2 + 1) * ST * PITCH + (wn * 16 + g) * PITCH;
#pragma unroll
    for (int s2 = 0; s2 < 2; s2++)
#pragma unroll
      for (int k4 = 0; k4 < 4; k4++) {
        const int k = s2 * 16 + 2 * k4 + 8 * (q >> 1) + (q & 1);
        double a[4], b[2];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) a[mt] = As[mt * 8 * PITCH + k];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) b[nt] = Bs[nt * 8 * PITCH + k];
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
          for (int nt = 0; nt < 2; nt++) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
      }
    __syncthreads();                     // the buffer is overwritten by the stage after next
  }

#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const long long row = c_row + wm * 32 + mt * 8 + g;
    if (row >= p.c_rows_end) continue;
    double* crow = p.C + row * p.ldc;
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
      const long long col = c_col + wn * 16 + nt * 8 + 2 * q;
      double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
      if (sub) { v0 = cf[mt][nt][0] - v0; v1 = cf[mt][nt][1] - v1; }
      else if (p.epi == EPI_NEG) { v0 = -v0; v1 = -v1; }
      if (col + 1 < p.c_cols_end) *reinterpret_cast<double2*>(crow + col) = make_double2(v0, v1);
      else if (col < p.c_cols_end) crow[col] = v0;
    }
  }
}


// In-place variant (C region == A region: the panel solve L[i,k] = K[i,k] inv(L_kk)^T of the factorisation).  With
// 64 x 64 quarters the quarter (qi, 1) would read the columns of A that quarter (qi, 0) overwrites, and nothing orders
// two CTAs.  Here a tile is split into four STRIPS of 32 rows x 128 columns instead: a CTA reads only the rows of A it
// writes, and every cp.async of its own has completed (and been consumed) before its first store.  Eight warps side
// by side, warp tile 32 x 16 as in the quarter kernel, the same DMMA sequence per output element.
constexpr int SR = 32;                   // strip rows
constexpr int STRIP_SMEM = 2 /*stages*/ * (SR + NB) * PITCH * 8;

__global__ void __launch_bounds__(SMALL_THREADS)
gemm_nt_strip_kernel(const GemmArgs p) {
  extern __shared__ __align__(16) double smem_d[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = blockIdx.x >> 2, strip = blockIdx.x & 3;
  const int ti = t % p.tiles_m, tj = t / p.tiles_m;          // never triangular: a panel is one tile column
  const long long a_row = (long long)p.a_row0 + ti * NB + strip * SR;
  const long long b_row = (long long)p.b_row0 + tj * p.b_tile_rows;
  const long long c_row = p.c_row0 + (long long)ti * NB + strip * SR;
  const long long c_col = p.c_col0 + (long long)tj * NB;
  if (c_row >= p.c_rows_end || c_col >= p.c_cols_end) return;
  const int nchunk = p.klen / KC;
  const int wn = warp, g = lane >> 2, q = lane & 3;
  // a lower-triangular B block (inverted diagonal block, B[n][c] = 0 for c > n): slab s (16 deep) is all zero for
  // this warp's columns [16 wn, 16 wn + 16) when s > wn
  const int last_slab = p.tri_b ? wn : 1 << 30;

  const long long a_rows_end = p.small_a_rows_end, b_rows_end = p.small_b_rows_end;
  constexpr int STAGE = (SR + NB) * PITCH;
  auto stage = [&](int chunk, int buf) {
#pragma unroll
    for (int u = 0; u < 10; u++) {
      const int piece = tid + SMALL_THREADS * u;          // 0..2559: 512 16-byte pieces of A (32 rows), 2048 of B (128 rows)
      const bool isb = piece >= SR * 16;
      const int pp = isb ? piece - SR * 16 : piece;
      const int r = pp >> 4, c16 = pp & 15;
      const long long grow = isb ? min(b_row + r, b_rows_end - 1) : min(a_row + r, a_rows_end - 1);
      const double* src = isb ? p.small_B + grow * p.small_ldb + p.b_col0 + chunk * KC + c16 * 2
                              : p.small_A + grow * p.small_lda + p.a_col0 + chunk * KC + c16 * 2;
      const uint32_t dst = smem_u32(smem_d + buf * STAGE + ((isb ? SR : 0) + r) * PITCH + c16 * 2);
      cp_async16(dst, src);
    }
    cp_async_commit();
  };

  double acc[4][2][2];
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  // fused forward substitution: z_k (128 x R, written by potf2 before this launch) and this strip's residual rows are
  // fetched now, so that their latency hides behind the main loop instead of following it on the critical path
  double zpre[4] = {0.0, 0.0, 0.0, 0.0}, rpre[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const long long rhs_row = p.rhs_r_row0 + (long long)ti * NB + strip * SR + tid;
  if (p.rhs_r != nullptr) {
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (tid + SMALL_THREADS * u < NB * p.rhs_R) zpre[u] = p.rhs_z[p.rhs_z_row0 * p.rhs_R + tid + SMALL_THREADS * u];
    if (tid < SR && rhs_row < p.rhs_rows_end) {
#pragma unroll
      for (int r = 0; r < 8; r++)
        if (r < p.rhs_R) rpre[r] = p.rhs_r[rhs_row * p.rhs_R + r];
    }
  }

  if (nchunk > 0) stage(0, 0);
  for (int c = 0; c < nchunk; c++) {
    if (c + 1 < nchunk) { stage(c + 1, (c + 1) & 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const double* As = smem_d + (c & 1) * STAGE + g * PITCH;
    const double* Bs = smem_d + (c & 1) * STAGE + (SR + wn * 16 + g) * PITCH;
#pragma unroll
    for (int s2 = 0; s2 < 2; s2++) {
      if (2 * c + s2 > last_slab) continue;
#pragma unroll
      for (int k4 = 0; k4 < 4; k4++) {
        const int k = s2 * 16 + 2 * k4 + 8 * (q >> 1) + (q & 1);
        double a[4], b[2];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) a[mt] = As[mt * 8 * PITCH + k];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) b[nt] = Bs[nt * 8 * PITCH + k];
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
          for (int nt = 0; nt < 2; nt++) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
      }
    }
    __syncthreads();                     // the buffer is overwritten by the stage after next
  }
  // every read of A by this CTA is complete (the last chunk was consumed before the barrier above); the stores
  // below touch only the 32 rows this CTA read
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const long long row = c_row + mt * 8 + g;
    if (row >= p.c_rows_end) continue;
    double* crow = p.C + row * p.ldc;
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
      const long long col = c_col + wn * 16 + nt * 8 + 2 * q;
      double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
      if (p.epi == EPI_NEG) { v0 = -v0; v1 = -v1; }
      if (col + 1 < p.c_cols_end) *reinterpret_cast<double2*>(crow + col) = make_double2(v0, v1);
      else if (col < p.c_cols_end) crow[col] = v0;
    }
  }
  if (p.rhs_r != nullptr) {
    // fused forward substitution (single-matrix fit): r_i -= L_ik z_k with the strip still in the accumulators.  A
    // strip holds all 128 columns of its 32 rows, so the row sums are complete inside the CTA: partial sums per
    // (row, warp) through shared memory, added in a fixed order, one plain read-modify-write per residual entry
    // (no other CTA of the launch touches these rows).
    const int R = p.rhs_R;
    double* zsm = smem_d;                        // [128][R]   z_k
    double* psm = smem_d + NB * 8;               // [32][8][R] partial sums
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (tid + SMALL_THREADS * u < NB * R) zsm[tid + SMALL_THREADS * u] = zpre[u];
    __syncthreads();
    for (int r = 0; r < R; r++) {
      double zv[2][2], sum[4];
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        const int col = wn * 16 + nt * 8 + 2 * q;
        zv[nt][0] = zsm[col * R + r];
        zv[nt][1] = zsm[(col + 1) * R + r];
      }
#pragma unroll
      for (int mt = 0; mt < 4; mt++) {
        sum[mt] = 0.0;
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
          sum[mt] = fma(acc[mt][nt][0], zv[nt][0], sum[mt]);
          sum[mt] = fma(acc[mt][nt][1], zv[nt][1], sum[mt]);
        }
        sum[mt] += __shfl_xor_sync(0xffffffffu, sum[mt], 1);
        sum[mt] += __shfl_xor_sync(0xffffffffu, sum[mt], 2);
        if (q == 0) psm[((mt * 8 + g) * 8 + wn) * R + r] = sum[mt];
      }
    }
    __syncthreads();
    if (tid < SR && rhs_row < p.rhs_rows_end) {
      double* rr = p.rhs_r + rhs_row * R;
      const double* ps = psm + tid * 8 * R;
#pragma unroll
      for (int r = 0; r < 8; r++)
        if (r < R)
          rr[r] = rpre[r] - (((ps[0 * R + r] + ps[1 * R + r]) + (ps[2 * R + r] + ps[3 * R + r])) +
                             ((ps[4 * R + r] + ps[5 * R + r]) + (ps[6 * R + r] + ps[7 * R + r])));
    }
  }
}

// the output tile range overlaps the rows and columns the A operand is read from
static bool gemm_small_inplace(const GemmArgs& a) {
  if (a.C != a.small_A) return false;
  const long long a_c0 = a.a_col0, a_c1 = a.a_col0 + a.klen;
  const long long c_c0 = a.c_col0, c_c1 = a.c_col0 + (long long)(a.tri ? a.tiles_m : a.tiles_n) * NB;
  return a_c0 < c_c1 && c_c0 < a_c1;      // column ranges intersect (the row ranges always do in the factorisation)
}

// true when `args` (tile mode, one matrix) is small enough for the latency kernel and carries raw operand pointers
bool gemm_small_eligible(const gpm_handle_impl* h, const GemmArgs& a, int batch) {
  if (h->opt.no_small_tiles || batch != 1 || a.small_A == nullptr || a.small_B == nullptr) return false;
  if (a.sweep_nblk > 0 || a.rowsq || a.kstart_mode || a.kend_mode || a.batch_cols) return false;
  // the fused forward substitution rides only on the strip kernel (a panel solve: in place, plain store, one tile column)
  if (a.rhs_r && !(gemm_small_inplace(a) && !a.tri && a.epi == EPI_STORE && a.tiles_n == 1)) return false;
  if (a.klen % KC != 0 || (a.small_lda & 1) || (a.small_ldb & 1) || (a.a_col0 & 1) || (a.b_col0 & 1)) return false;
  // in place (the output overwrites the A operand): only the strip kernel is safe, and it handles plain
  // EPI_STORE / EPI_NEG tile columns
  if (gemm_small_inplace(a) && (a.tri || a.epi == EPI_SUB)) return false;
  return gemm_grid_x(a) <= SMALL_MAX_TILES;
}

int launch_gemm_small(gpm_handle_impl* h, const GemmArgs& a, cudaStream_t stream) {
  if (!h->gemm_small_attr) {
    GPM_CUDA(cudaFuncSetAttribute(gemm_nt_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM));
    h->gemm_small_attr = true;
  }
  if (gemm_small_inplace(a)) {
    if (!h->gemm_strip_attr) {
      GPM_CUDA(cudaFuncSetAttribute(gemm_nt_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STRIP_SMEM));
      h->gemm_strip_attr = true;
    }
    gemm_nt_strip_kernel<<<gemm_grid_x(a) * 4, SMALL_THREADS, STRIP_SMEM, stream>>>(a);
    GPM_LAUNCH_CHECK();
    return 0;
  }
  gemm_nt_small_kernel<<<gemm_grid_x(a) * 4, SMALL_THREADS, SMALL_SMEM, stream>>>(a);
  GPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace gpm
