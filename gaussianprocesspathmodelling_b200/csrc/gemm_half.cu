// Half-tile variant of the FP64 tensor-core tile GEMM for the whole-batch launches of the batched fits (batch >= 32).
//
// gemm_nt_kernel gives a 128 x 128 tile to one CTA that owns the SM: its eight consumer warps leave the DMMA pipe idle
// while they store a finished tile (and, in a panel solve, update the fused forward substitution), which at the
// K = 128 / 256 contractions of an N = 512 path is 25 - 40 % of a tile (ncu on the six launches of a 4096-path call:
// DMMA sub-pipe 58 - 62 % active in the panel solves, 75 - 78 % in the updates).  Here a CTA works on 64 x 128
// half-tiles with 112 KB of shared memory and TWO CTAs share an SM, so that one's epilogue can run under the other's
// main loop; the DMMA sequence per output element is the tile kernel's, so the factor is bitwise the same.  Eight warps
// per CTA (32 x 32 warp tiles, 64 accumulator registers; four warps per scheduler) by default: two warps per scheduler
// (the first version: four warps on 64 x 32 warp tiles) do not keep its DMMA pipe busy.  There is no
// producer warp: thread 0 refills a ring stage when all warps have released it, as in pathfit.cu.
//
// A CTA takes one 128 x 128 tile of the launch as two half-tiles in turn; the [128 x 16] B slabs are loaded once per
// half (the second time from L2).  C -= A B^T prefetches the C half-tile by TMA under the main loop (two-stage ring);
// plain-store launches use the free C space for two more stages.
#include "gemm.cuh"
#include "halftile.cuh"

namespace gpm {

constexpr int GH_THREADS = 128;                          // four-warp variant; the eight-warp one runs 256
constexpr int GH_STAGES = 2;                             // ring stages of their own (plain-store launches add up to two in the C space)
constexpr int GH_A_BYTES = 64 * SLAB_K * 8;              // 8 KB   [64 rows x 16] slab
constexpr int GH_B_BYTES = NB * SLAB_K * 8;              // 16 KB  [128 rows x 16] slab
constexpr int GH_STAGE_BYTES = GH_A_BYTES + GH_B_BYTES;  // 24 KB
constexpr int GH_C_BYTES = 64 * NB * 8;                  // 64 KB: the C half-tile (C -= A B^T) / scratch of the fused forward substitution
constexpr int GH_RING_OFF = GH_C_BYTES;
constexpr int GH_BAR_OFF = GH_RING_OFF + GH_STAGES * GH_STAGE_BYTES;
constexpr int GH_SMEM = GH_BAR_OFF + 128;

template <int NT>
__device__ __forceinline__ void gh_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
__device__ __forceinline__ void gh_cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// NW = 4: 64 x 32 warp tiles (128 accumulator registers).  NW = 8: 32 x 32 warp tiles (64 accumulator registers, at most
// 128 registers per thread), warp (wr, wq) = sub-tile rows [4 wr, 4 wr + 4) x column set wq: sixteen warps per SM, four
// per scheduler -- a scheduler's DMMA pipe is not kept busy by two warps' LDS -> DMMA streams.
// NST: ring depth.  C -= A B^T needs the C space for its prefetched C half-tile (2 stages); a plain-store launch (the panel
// solves: 56 % of the DMMAs per slab, so a slab is consumed faster than a TMA round trip once four warps per scheduler
// keep the pipe busy: mbarrier waits were 15 % of the stall samples) puts stages 2 and 3 at the end of the C space
// (40 KB, 16 KB): three stages leave 40 KB of scratch for the fused forward substitution (3.5 KB per right-hand side),
// four stages 16 KB.
template <int NW, int NST>
__global__ void __launch_bounds__(NW * 32, 2)
gemm_nt_half_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ CUtensorMap mapC, const GemmArgs p) {
  extern __shared__ __align__(1024) uint8_t gh_smem[];
  const uint32_t base = smem_u32(gh_smem);
  const uint32_t cbuf = base, ring = base + GH_RING_OFF;
  const uint32_t bar_full = base + GH_BAR_OFF, bar_empty = bar_full + NST * 8, bar_c = bar_empty + NST * 8;
  double* gen = reinterpret_cast<double*>(gh_smem);
  constexpr int NT = NW * 32, NMT = 32 / NW;               // threads; sub-tile rows per warp
  const int tid = threadIdx.x, wq = (tid >> 5) & 3, wr = tid >> 7, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int w = wq;                                         // column set of this warp
  const int mt0 = wr * NMT;                                 // first sub-tile row of this warp (0 with four warps)
  const long long bz = blockIdx.y;
  if (tid == 0) {
    for (int s = 0; s < NST; s++) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, NW); }
    mbar_init(bar_c, 1);
    fence_mbar_init();
    prefetch_tmap(&mapA); prefetch_tmap(&mapB); prefetch_tmap(&mapC);
  }
  __syncthreads();
  if ((base & 1023u) != 0) return;                         // the 128-byte swizzle needs a 1 KB-aligned window (never taken)

  // the tile of this CTA (same enumeration as gemm_nt_kernel)
  int ti, tj;
  {
    const int t = blockIdx.x;
    if (p.tri) {
      int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
      while ((i + 1) * (i + 2) / 2 <= t) i++;
      while (i * (i + 1) / 2 > t) i--;
      ti = i; tj = t - i * (i + 1) / 2;
    } else {
      ti = t % p.tiles_m; tj = t / p.tiles_m;
    }
  }
  const int nslab = p.klen / SLAB_K;
  const int a_row = p.a_row0 + ti * NB + (int)(bz * p.batch_a_rows);
  const int b_row = p.b_row0 + tj * p.b_tile_rows + (int)(bz * p.batch_b_rows);
  const long long c_row = p.c_row0 + (long long)ti * NB + bz * p.batch_c_rows;
  const long long c_col = p.c_col0 + (long long)tj * NB;
  const long long rows_end = p.c_rows_end + bz * p.batch_c_rows;
  const bool sub = p.epi == EPI_SUB;
  const bool diag = p.diag_lower && ti == tj;

  uint32_t off[4];
#pragma unroll
  for (int t = 0; t < 4; t++) off[t] = frag_off(g, q, t);
  const int cset[4] = {w, 7 - w, 8 + w, 15 - w};           // this warp's 8-column sub-tile columns (ascending)
  uint32_t boff[4];
#pragma unroll
  for (int nt = 0; nt < 4; nt++) boff[nt] = (uint32_t)cset[nt] * 1024u;

  // fused forward substitution: z_k of this matrix and the residual rows of this tile come in by cp.async while the
  // main loop runs (no registers, no exposed global round trip in the epilogue: with staged loads there the panel
  // launches ran their DMMA pipe 63 % active); scratch in the C space (unused by a plain-store launch):
  // partial sums [64][4][R] | z_k [128][R] | residual rows [128][R]
  const int R = p.rhs_R;
  double* psm = gen;
  double* zsm = gen + 256 * R;
  double* rsm = zsm + NB * R;
  if (p.rhs_r != nullptr) {
    const double* zk = p.rhs_z + (p.rhs_z_row0 + bz * p.batch_rhs_rows) * R;
    const long long rrow0 = p.rhs_r_row0 + (long long)ti * NB;
    const double* rk = p.rhs_r + (rrow0 + bz * p.batch_rhs_rows) * R;
    const long long nr = p.rhs_rows_end - rrow0;           // valid residual rows of this tile (>= 1)
    for (int idx = tid; idx < NB * R; idx += NT) {
      gh_cp_async8(smem_u32(zsm + idx), zk + idx);
      if (idx < nr * R) gh_cp_async8(smem_u32(rsm + idx), rk + idx);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  constexpr int nst = NST;
  auto stage_addr = [&](int st) {
    return st < GH_STAGES ? ring + st * GH_STAGE_BYTES : cbuf + GH_C_BYTES - (st - GH_STAGES + 1) * GH_STAGE_BYTES;
  };
  int sg = 0;                                               // slabs consumed so far (all threads agree)
  int issued = 0;                                           // thread 0: slabs issued so far
  uint32_t cpar = 0;
  const int total_slabs = 2 * nslab;
  auto issue = [&](int n) {                                  // thread 0: slab n of the CTA's stream (half = n / nslab)
    const int h = n / nslab, s = n - h * nslab, st = n % nst;
    const uint32_t dst = stage_addr(st), bar = bar_full + st * 8;
    mbar_arrive_expect_tx(bar, GH_STAGE_BYTES);
    tma_load_2d(dst, &mapA, p.a_col0 + s * SLAB_K, a_row + 64 * h, bar);
    tma_load_2d(dst + GH_A_BYTES, &mapB, p.b_col0 + s * SLAB_K, b_row, bar);
  };
  if (tid == 0) { for (; issued < nst; issued++) issue(issued); }   // klen >= 32 (two slabs per half) is checked by the launcher

  for (int h = 0; h < 2; h++) {
    const long long row0 = c_row + 64 * h;
    if (sub && tid == 0) {                                  // the C half-tile lands under the main loop
      mbar_arrive_expect_tx(bar_c, GH_C_BYTES);
#pragma unroll
      for (int b = 0; b < NB / SLAB_K; b++)
        tma_load_2d(cbuf + b * GH_A_BYTES, &mapC, (int)c_col + b * SLAB_K, (int)row0, bar_c);
    }
    double acc[NMT][4][2];
#pragma unroll
    for (int mt = 0; mt < NMT; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    // one slab step of this warp's rows (the row range is a template parameter: warp-uniform dispatch)
#define GH_SLAB(SH)                                                                            \
  do {                                                                                         \
    if constexpr (NW == 4) pf_slab<SH, 0, 8>(acc, sa, sb, off, boff);                          \
    else if (wr == 0) pf_slab<SH, 0, 4>(acc, sa, sb, off, boff);                               \
    else pf_slab<SH, 4, 4>(acc, sa, sb, off, boff);                                            \
  } while (0)
    for (int s = 0; s < nslab; s++, sg++) {
      const int st = sg % nst;
      mbar_wait(bar_full + st * 8, (sg / nst) & 1);
      const uint32_t sa = stage_addr(st), sb = sa + GH_A_BYTES;
      if (diag) {
        switch (2 * w + h) {                                 // warp-uniform: lower-triangle shape of this warp's columns
          case 0: GH_SLAB(PF_DIAG + 0); break;
          case 1: GH_SLAB(PF_DIAG + 1); break;
          case 2: GH_SLAB(PF_DIAG + 2); break;
          case 3: GH_SLAB(PF_DIAG + 3); break;
          case 4: GH_SLAB(PF_DIAG + 4); break;
          case 5: GH_SLAB(PF_DIAG + 5); break;
          case 6: GH_SLAB(PF_DIAG + 6); break;
          default: GH_SLAB(PF_DIAG + 7); break;
        }
      } else if (p.tri_b) {
        int dead = 0;                                       // B lower triangular: sub-tile column c needs slab s iff 2 s <= c
#pragma unroll
        for (int nt = 0; nt < 4; nt++) dead += (cset[nt] < 2 * s) ? 1 : 0;
        switch (dead) {
          case 0: GH_SLAB(PF_COLS + 0); break;
          case 1: GH_SLAB(PF_COLS + 1); break;
          case 2: GH_SLAB(PF_COLS + 2); break;
          case 3: GH_SLAB(PF_COLS + 3); break;
          default: break;
        }
      } else {
        GH_SLAB(PF_FULL);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + st * 8);
      if (tid == 0 && issued < total_slabs) {               // refill the stage just released (by all warps)
        mbar_wait(bar_empty + st * 8, (sg / nst) & 1);
        issue(issued);
        issued++;
      }
    }
    // ---------------- epilogue of the half-tile ----------------
    const bool interior = row0 + 64 <= rows_end && c_col + NB <= p.c_cols_end;
    if (sub) {
      mbar_wait(bar_c, cpar);
      cpar ^= 1u;
#pragma unroll
      for (int ml = 0; ml < NMT; ml++) {
        const int mt = mt0 + ml;
        const long long row = row0 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          const int c = cset[nt];
          if (diag && c > mt + 8 * h) continue;             // above the diagonal of a symmetric tile: never read
          const uint32_t addr = cbuf + (c >> 1) * GH_A_BYTES + (mt * 8 + g) * 128 + ((((c & 1) * 4 + q) ^ g) << 4);
          double c0, c1;
          asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(c0), "=d"(c1) : "r"(addr));
          const long long col = c_col + c * 8 + 2 * q;
          const double v0 = c0 - acc[ml][nt][0], v1 = c1 - acc[ml][nt][1];
          if (interior) {
            *reinterpret_cast<double2*>(p.C + row * p.ldc + col) = make_double2(v0, v1);
          } else if (row < rows_end) {
            double* dst = p.C + row * p.ldc + col;
            if (col + 1 < p.c_cols_end) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            else if (col < p.c_cols_end) *dst = v0;
          }
        }
      }
      gh_sync<NT>();                                        // the C space may be refilled (next half's prefetch)
    } else {
      if (p.epi == EPI_STORE && row0 + 64 <= rows_end && c_col + NB <= p.c_cols_end) {
        // interior half-tile (all of them when N is a multiple of 128): no bounds tests, one walking pointer
        double* crow = p.C + (row0 + mt0 * 8 + g) * p.ldc + c_col + 2 * q;
#pragma unroll
        for (int mt = 0; mt < NMT; mt++) {
#pragma unroll
          for (int nt = 0; nt < 4; nt++)
            *reinterpret_cast<double2*>(crow + cset[nt] * 8) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
          crow += 8 * p.ldc;
        }
      } else
#pragma unroll
      for (int ml = 0; ml < NMT; ml++) {
        const long long row = row0 + (mt0 + ml) * 8 + g;
        if (row < rows_end) {
          double* crow = p.C + row * p.ldc;
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const long long col = c_col + cset[nt] * 8 + 2 * q;
            double v0 = acc[ml][nt][0], v1 = acc[ml][nt][1];
            if (p.epi == EPI_NEG) { v0 = -v0; v1 = -v1; }
            if (col + 1 < p.c_cols_end) *reinterpret_cast<double2*>(crow + col) = make_double2(v0, v1);
            else if (col < p.c_cols_end) crow[col] = v0;
          }
        }
      }
      if (p.rhs_r != nullptr) {
        // fused forward substitution: r_i -= L_ik z_k with the half-tile still in the accumulators; partial sums per
        // (row, warp) through shared memory, added in a fixed order; one plain read-modify-write per residual entry
        // (no other CTA of the launch touches these rows)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        gh_sync<NT>();                                      // z_k and the residual rows have landed; the previous half's partial sums were consumed
        for (int r = 0; r < R; r++) {
          double zv[4][2], sum[NMT];
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const int cl = cset[nt] * 8 + 2 * q;
            zv[nt][0] = zsm[cl * R + r];
            zv[nt][1] = zsm[(cl + 1) * R + r];
          }
#pragma unroll
          for (int mt = 0; mt < NMT; mt++) sum[mt] = 0.0;
#pragma unroll
          for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int mt = 0; mt < NMT; mt++) {
              sum[mt] = fma(acc[mt][nt][0], zv[nt][0], sum[mt]);
              sum[mt] = fma(acc[mt][nt][1], zv[nt][1], sum[mt]);
            }
#pragma unroll
          for (int mt = 0; mt < NMT; mt++) sum[mt] += __shfl_xor_sync(0xffffffffu, sum[mt], 1);
#pragma unroll
          for (int mt = 0; mt < NMT; mt++) sum[mt] += __shfl_xor_sync(0xffffffffu, sum[mt], 2);
          if (q == 0) {
#pragma unroll
            for (int mt = 0; mt < NMT; mt++) psm[(((mt0 + mt) * 8 + g) * 4 + w) * R + r] = sum[mt];
          }
        }
        gh_sync<NT>();
        if (tid < 64) {
          const long long row = p.rhs_r_row0 + (long long)ti * NB + 64 * h + tid;
          if (row < p.rhs_rows_end) {
            double* rr = p.rhs_r + (row + bz * p.batch_rhs_rows) * R;
            for (int r = 0; r < R; r++)
              rr[r] = rsm[(64 * h + tid) * R + r] - ((psm[(tid * 4 + 0) * R + r] + psm[(tid * 4 + 1) * R + r]) +
                                                     (psm[(tid * 4 + 2) * R + r] + psm[(tid * 4 + 3) * R + r]));
          }
        }
      }
    }
  }
}

// raw operand pointers present, a plain tile launch of a large batch: the half-tile kernel can take it
bool gemm_half_eligible(const gpm_handle_impl* h, const GemmArgs& a, int batch) {
  if (h->opt.no_half_tiles || batch < 32 || a.small_A == nullptr || a.small_B == nullptr) return false;
  if (a.sweep_nblk > 0 || a.rowsq || a.kstart_mode || a.kend_mode || a.batch_cols) return false;
  if (a.klen < 2 * SLAB_K || a.klen % SLAB_K != 0) return false;
  if (a.epi != EPI_STORE && a.epi != EPI_SUB) return false;
  if (a.rhs_r && (a.epi != EPI_STORE || a.rhs_R < 1 || a.rhs_R > 8)) return false;
  if ((a.small_lda & 1) || (a.small_ldb & 1) || (a.ldc & 1)) return false;
  return true;
}

int launch_gemm_half(gpm_handle_impl* h, const GemmArgs& a, int batch, cudaStream_t stream) {
  if (!h->gemm_half_attr) {
    auto opt_in = [](const void* f) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, GH_SMEM);
      return e != cudaSuccess ? e : cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    };
    GPM_CUDA(opt_in((const void*)gemm_nt_half_kernel<4, 2>));
    GPM_CUDA(opt_in((const void*)gemm_nt_half_kernel<8, 2>));
    GPM_CUDA(opt_in((const void*)gemm_nt_half_kernel<8, 3>));
    GPM_CUDA(opt_in((const void*)gemm_nt_half_kernel<8, 4>));
    h->gemm_half_attr = true;
  }
  CUtensorMap mapA, mapB, mapC;
  int rc;
  // [64 x 16] boxes over the A operand and over C, [128 x 16] boxes over B; the tensors span whole leading dimensions
  if ((rc = make_tmap(h, &mapA, a.small_A, a.small_a_rows_end, a.small_lda, a.small_lda, 64))) return rc;
  if ((rc = make_tmap(h, &mapB, a.small_B, a.small_b_rows_end, a.small_ldb, a.small_ldb, NB))) return rc;
  if ((rc = make_tmap(h, &mapC, a.C, a.small_a_rows_end, a.ldc, a.ldc, 64))) return rc;
  dim3 grid(gemm_grid_x(a), batch);
  // ring depth: C -= A B^T keeps the C space for its C half-tile (2); plain stores go as deep as the scratch of the
  // fused forward substitution allows (4 stages up to R = 4, 3 up to R = 8)
  int nst = 2;
  if (a.epi == EPI_STORE && a.klen >= 4 * SLAB_K) nst = (a.rhs_r == nullptr || a.rhs_R <= 4) ? 4 : 3;
  if (h->opt.half_stages >= 2 && h->opt.half_stages < nst) nst = h->opt.half_stages;
  if (h->opt.half_warps == 4) gemm_nt_half_kernel<4, 2><<<grid, 128, GH_SMEM, stream>>>(mapA, mapB, mapC, a);
  else if (nst == 4) gemm_nt_half_kernel<8, 4><<<grid, 256, GH_SMEM, stream>>>(mapA, mapB, mapC, a);
  else if (nst == 3) gemm_nt_half_kernel<8, 3><<<grid, 256, GH_SMEM, stream>>>(mapA, mapB, mapC, a);
  else gemm_nt_half_kernel<8, 2><<<grid, 256, GH_SMEM, stream>>>(mapA, mapB, mapC, a);
  GPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace gpm
