// FP64 tensor-core tile GEMM for sm_100a: 128x128 CTA tile, operands staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle) through a 3-stage mbarrier ring, one producer warp and
// eight consumer warps issuing mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) on 64x32 warp tiles.
//
// A CTA works through `tiles_per_cta` consecutive tiles; the operand ring keeps running across
// tiles, and for C -= A B^T the C tile is prefetched into shared memory by TMA while the main loop
// runs, so neither the TMA prologue nor the C read of the epilogue is exposed.
//
// FP64 has no tcgen05 kind, so the accumulators live in registers (128 per thread), not TMEM.
#include "gemm.cuh"

namespace gpm {

constexpr int STAGES = 3;
constexpr int CONSUMER_WARPS = 8;
constexpr int GEMM_THREADS = (CONSUMER_WARPS + 4) * 32;   // two consumer warpgroups + one producer warpgroup (setmaxnreg)
constexpr int C_BYTES = NB * NB * 8;
constexpr int GEMM_SMEM = STAGES * 2 * SLAB_BYTES + C_BYTES + 1024 /*align slack*/ + (2 * STAGES + 3) * 8;

struct TileCoord { int ti, tj; };

__device__ __forceinline__ TileCoord tile_coord(const GemmArgs& p, int t) {
  TileCoord c;
  if (p.tri) {
    int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= t) i++;
    while (i * (i + 1) / 2 > t) i--;
    c.ti = i;
    c.tj = t - i * (i + 1) / 2;
  } else {
    c.ti = t % p.tiles_m;
    c.tj = t / p.tiles_m;
  }
  return c;
}

// One unit of work on a tile: a contraction of nslab slabs and an epilogue.
struct OpDesc {
  int nslab, a_col0, b_col0, b_row, epi;
  long long c_col;
  bool second_b, rowsq;
  bool a_smem;     // sweep: the A operand is the previous op's result, resident in the C-tile buffer
  bool to_smem;    // sweep: the result stays in the C-tile buffer (swizzled slab layout) instead of going to global
  bool signal;     // sweep: publish the global stores of this op (the next update reads them by TMA)
  bool tri_b;      // B is a lower-triangular 128x128 block: slab s is all zero for columns below 16 s
};

__device__ __forceinline__ OpDesc make_op(const GemmArgs& p, const TileCoord& tc, int o, long long bz) {
  OpDesc d;
  if (p.sweep_nblk > 0) {
    const int k0 = p.sweep_tri ? tc.ti : 0;        // first block column this row block touches
    const int k = k0 + ((o + 1) >> 1);
    const bool isU = (o & 1) != 0;                 // o odd: update, o even: diagonal multiply
    d.nslab = isU ? (k - k0) * (NB / SLAB_K) : NB / SLAB_K;
    d.a_col0 = isU ? k0 * NB : k * NB;
    d.b_col0 = isU ? k0 * NB : 0;
    d.b_row = k * NB;
    d.epi = isU ? EPI_SUB : EPI_STORE;
    d.c_col = (long long)k * NB;
    d.second_b = !isU;
    d.rowsq = !isU;
    d.a_smem = !isU && o > 0;     // D(k), k > k0: A = R left in shared memory by U(k); the first D reads K*^T from global
    d.to_smem = isU;
    d.signal = !isU;
    d.tri_b = !isU;               // inv(L_kk)
  } else {
    const int lo = p.kstart_mode == 1 ? tc.ti * NB : 0;
    int hi = p.klen;
    if (p.kend_mode == 1) hi = min(hi, (tc.ti + 1) * NB);
    if (p.kend_mode == 2) hi = min(hi, (tc.tj + 1) * NB);
    const int bcol = (int)(bz * p.batch_cols);
    d.nslab = (hi - lo) / SLAB_K;
    d.a_col0 = p.a_col0 + lo + bcol;
    d.b_col0 = p.b_col0 + lo + bcol;
    d.b_row = p.b_row0 + tc.tj * p.b_tile_rows + (int)(bz * p.batch_b_rows);
    d.epi = p.epi;
    d.c_col = p.c_col0 + (long long)tc.tj * NB + bz * p.batch_cols;
    d.second_b = false;
    d.rowsq = p.rowsq != nullptr;
    d.a_smem = d.to_smem = d.signal = false;
    d.tri_b = p.tri_b != 0;
  }
  return d;
}

#ifdef GPM_GEMM_TIMING
__device__ long long g_gemm_marks[8 * 2 * 4 * 8];   // [launch % 8][warp 0 / 7][tile of the CTA][slot]
#define GM_MARK(slot) if (DG && blockIdx.y == 7 && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 7) && t - t_begin < 4) g_gemm_marks[((p.dbg_launch * 2 + (warp ? 1 : 0)) * 4 + (t - t_begin)) * 8 + (slot)] = clock64();
#else
#define GM_MARK(slot)
#endif

// One slab (16 contraction steps) of a warp's 64 x 32 tile: 8 x 4 sub-tiles of 8 x 8, DMMA.8x8x4.  Sub-tile
// (mt, nt) is computed iff nt <= mt + D; D = 8 keeps all of them, D = 0 / -4 are the two staircase shapes a warp
// sees on a symmetric diagonal tile (GemmArgs::diag_lower).
template <int D>
__device__ __forceinline__ void slab_mma(double (&acc)[8][4][2], uint32_t sa, uint32_t sb, const uint32_t (&off)[4]) {
#pragma unroll
  for (int k4 = 0; k4 < 4; k4++) {
    double a[8], b[4];
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
      if (mt + D >= 0) a[mt] = lds_f64(sa + mt * 1024 + off[k4]);
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
      if (nt <= 7 + D) b[nt] = lds_f64(sb + nt * 1024 + off[k4]);
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++)
        if (nt <= mt + D) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}

// DG: instantiation for the factorisation (diag_lower staircase paths, fused forward substitution); the plain
// one keeps the variance sweep's loops free of the extra branches
template <bool DG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapB2,
               const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // swizzle needs 1024 B
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t cbuf = base + STAGES * 2 * SLAB_BYTES;
  const uint32_t bar_full = cbuf + C_BYTES;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_cfull = bar_empty + STAGES * 8;
  const uint32_t bar_cempty = bar_cfull + 8;
  const uint32_t bar_dep = bar_cempty + 8;          // sweep mode: previous op's global writes are visible
  double* red = reinterpret_cast<double*>(gen + STAGES * 2 * SLAB_BYTES);  // aliases cbuf (EPI_STORE only)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef GPM_GEMM_TIMING
  const long long t_entry = clock64();
#endif
  const int total = p.tri ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
  // tiles of this CTA: a run of consecutive tiles, or (triangular sweep) the pair (b, tiles_m-1-b), which
  // balances the quadratically decreasing work of the row blocks
  const int t_begin = p.sweep_tri ? 0 : blockIdx.x * p.tiles_per_cta;
  const int t_end = p.sweep_tri ? ((2 * (int)blockIdx.x == p.tiles_m - 1) ? 1 : 2) : min(total, t_begin + p.tiles_per_cta);
  auto tile_id = [&](int t) { return p.sweep_tri ? (t == 0 ? (int)blockIdx.x : p.tiles_m - 1 - (int)blockIdx.x) : t; };
  const long long bz = blockIdx.y;
  const bool chained = p.sweep_nblk > 0;
  auto ops_of = [&](const TileCoord& tc) {
    return chained ? 2 * (p.sweep_nblk - (p.sweep_tri ? tc.ti : 0)) - 1 : 1;
  };

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, CONSUMER_WARPS);
    }
    mbar_init(bar_cfull, 1);
    mbar_init(bar_cempty, CONSUMER_WARPS);
    mbar_init(bar_dep, CONSUMER_WARPS);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp >= CONSUMER_WARPS) {
    // ===== producer warpgroup: hand its registers to the consumers; one elected lane of its first warp drives TMA =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp != CONSUMER_WARPS) return;
    if (lane == 0) {
      prefetch_tmap(&mapA);
      prefetch_tmap(&mapB);
      prefetch_tmap(&mapC);
      if (chained) prefetch_tmap(&mapB2);
      int sg = 0, ct = 0, nsig = 0;        // ring slot counter, C prefetches issued, signalling ops issued
      for (int t = t_begin; t < t_end; t++) {
        const TileCoord tc = tile_coord(p, tile_id(t));
        const int a_row = p.a_row0 + tc.ti * NB + (int)(bz * p.batch_a_rows);
        const int nops = ops_of(tc);
        for (int o = 0; o < nops; o++) {
          const OpDesc d = make_op(p, tc, o, bz);
          const bool sub = d.epi == EPI_SUB;
          const int c_at = min(STAGES - 1, d.nslab - 1);      // issue the C prefetch after this slab
          // sweep: an update reads the block column the previous diagonal multiply stored (its last 8 slabs)
          const int dep_at = (chained && sub && nsig > 0) ? max(0, d.nslab - NB / SLAB_K) : -1;
          for (int s = 0; s < d.nslab; s++, sg++) {
            if (s == dep_at) mbar_wait(bar_dep, (nsig - 1) & 1);
            const int st = sg % STAGES;
            if (sg >= STAGES) mbar_wait(bar_empty + st * 8, ((sg / STAGES) - 1) & 1);
            const uint32_t dst = base + st * 2 * SLAB_BYTES;
            if (d.a_smem) {
              mbar_arrive_expect_tx(bar_full + st * 8, SLAB_BYTES);
            } else {
              mbar_arrive_expect_tx(bar_full + st * 8, 2 * SLAB_BYTES);
              tma_load_2d(dst, &mapA, d.a_col0 + s * SLAB_K, a_row, bar_full + st * 8);
            }
            tma_load_2d(dst + SLAB_BYTES, d.second_b ? &mapB2 : &mapB, d.b_col0 + s * SLAB_K, d.b_row,
                        bar_full + st * 8);
            if (sub && s == c_at) {
              // the C-tile buffer must be free: in the sweep it is busy until the previous diagonal multiply has
              // finished (operand + row-sum scratch); otherwise until the previous epilogue has read it
              if (chained) { if (nsig > 0) mbar_wait(bar_dep, (nsig - 1) & 1); }
              else if (ct > 0) mbar_wait(bar_cempty, (ct - 1) & 1);
              mbar_arrive_expect_tx(bar_cfull, C_BYTES);
              const int c_row = (int)(p.c_row0 + (long long)tc.ti * NB + bz * p.batch_c_rows);
#pragma unroll
              for (int b = 0; b < NB / SLAB_K; b++)
                tma_load_2d(cbuf + b * SLAB_BYTES, &mapC, (int)d.c_col + b * SLAB_K, c_row, bar_cfull);
              ct++;
            }
          }
          if (d.signal) nsig++;
        }
      }
    }
    return;
  }

  // ===== consumers: 2 (m) x 4 (n) warps, warp tile 64 x 32 =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  // The second row of warps takes the column groups in reverse order, so that each scheduler (warp % 4) holds
  // column groups {w, 3-w}: when slabs are skipped for a triangular B (group w needs 2w+2 of 8 slabs), every
  // scheduler keeps 10 of 16 slab-units and the saving is not lost to imbalance.
  const int wm = warp >> 2, wn = (warp & 3) ^ (wm ? 3 : 0);
  const int g = lane >> 2, q = lane & 3;
  uint32_t off[4];
#pragma unroll
  for (int t = 0; t < 4; t++) off[t] = frag_off(g, q, t);
  const uint32_t a_warp = wm * 64 * 128, b_warp = SLAB_BYTES + wn * 32 * 128;
  const long long rows_end = p.c_rows_end + bz * p.batch_c_rows;
  int sg = 0, ct = 0;
  // Symmetric diagonal tiles (diag_lower): sub-tile (mt, nt) of this warp lies on or below the diagonal iff
  // nt <= mt + dd with dd = 8 wm - 4 wn.  dd >= 3: all kept; dd = 0 / -4: staircases of 26 / 10 sub-tiles;
  // dd <= -8: none.  With the column-group reversal above every scheduler keeps 32 or 36 of its 64 sub-tiles.
  const int dd = 8 * wm - 4 * wn;
  const int dcode = dd >= 3 ? 0 : (dd == 0 ? 1 : (dd == -4 ? 2 : 3));

  int rhs_tiles = 0;
  if (DG && p.rhs_r != nullptr) {
    // z_k is the same for every tile of this launch and matrix: stage it once; the C-tile buffer is free in
    // an EPI_STORE launch.  The loads overlap the first tile's main loop.
    const double* zk = p.rhs_z + (p.rhs_z_row0 + bz * p.batch_rhs_rows) * p.rhs_R;
    for (int idx = tid; idx < NB * p.rhs_R; idx += CONSUMER_WARPS * 32) red[idx] = zk[idx];
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
  }

  for (int t = t_begin; t < t_end; t++) {
    const TileCoord tc = tile_coord(p, tile_id(t));
    const int nops = ops_of(tc);
    for (int o = 0; o < nops; o++) {
      const OpDesc d = make_op(p, tc, o, bz);
      const bool sub = d.epi == EPI_SUB;
      const int dsel = (DG && p.diag_lower && tc.ti == tc.tj) ? dcode : 0;
      double acc[8][4][2];
#pragma unroll
      for (int mt = 0; mt < 8; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
      GM_MARK(0)
#ifdef GPM_GEMM_TIMING
      if (DG && blockIdx.y == 7 && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 7) && t == t_begin)
        g_gemm_marks[((p.dbg_launch * 2 + (warp ? 1 : 0)) * 4) * 8 + 7] = t_entry;
#endif

      for (int s = 0; s < d.nslab; s++, sg++) {
        const int st = sg % STAGES;
        mbar_wait(bar_full + st * 8, (sg / STAGES) & 1);
        const uint32_t sa = (d.a_smem ? cbuf + s * SLAB_BYTES : base + st * 2 * SLAB_BYTES) + a_warp;
        const uint32_t sb = base + st * 2 * SLAB_BYTES + b_warp;
        if (dsel == 0) {
          if (!d.tri_b || s <= 2 * wn + 1) slab_mma<8>(acc, sa, sb, off);
        } else if (dsel == 1) {
          slab_mma<0>(acc, sa, sb, off);
        } else if (dsel == 2) {
          slab_mma<-4>(acc, sa, sb, off);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + st * 8);
      }

      GM_MARK(1)
      // ===== epilogue: registers (-> C from smem) -> global, 16-byte stores =====
      const long long crow_base = p.c_row0 + (long long)tc.ti * NB + wm * 64 + bz * p.batch_c_rows;
      const long long ccol_base = d.c_col + wn * 32 + 2 * q;
      if (sub) {
        mbar_wait(bar_cfull, ct & 1);
        GM_MARK(6)
#pragma unroll
        for (int mt = 0; mt < 8; mt++) {
          const int r = wm * 64 + mt * 8 + g;
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            // C tile in smem: 8 boxes of [128 rows x 16 cols], 128B-swizzled; this thread's column pair is
            // one 16-byte chunk:  box = col/16, chunk = (col%16)/2 ^ (row&7)
            const uint32_t addr = cbuf + (wn * 2 + (nt >> 1)) * SLAB_BYTES + r * 128 + ((((nt & 1) * 4 + q) ^ g) << 4);
            double c0, c1;
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(c0), "=d"(c1) : "r"(addr));
            acc[mt][nt][0] = c0 - acc[mt][nt][0];
            acc[mt][nt][1] = c1 - acc[mt][nt][1];
            if (d.to_smem)        // same layout as an A-operand slab: the next op contracts over these columns
              asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(acc[mt][nt][0]), "d"(acc[mt][nt][1]) : "memory");
          }
        }
        ct++;
        if (d.to_smem) {
          // the result stays on chip; every consumer warp must have written its part before anyone reads it
          asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
          continue;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_cempty);
      }
      if (d.epi == EPI_NEG) {
#pragma unroll
        for (int mt = 0; mt < 8; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++) { acc[mt][nt][0] = -acc[mt][nt][0]; acc[mt][nt][1] = -acc[mt][nt][1]; }
      }
#pragma unroll
      for (int mt = 0; mt < 8; mt++) {
        const long long row = crow_base + mt * 8 + g;
        double sq = 0.0;
        if (row < rows_end) {
          double* crow = p.C + row * p.ldc;
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const long long col = ccol_base + nt * 8;
            const double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
            if (col + 1 < p.c_cols_end) {
              *reinterpret_cast<double2*>(crow + col) = make_double2(v0, v1);
              sq += v0 * v0 + v1 * v1;
            } else if (col < p.c_cols_end) {
              crow[col] = v0;
              sq += v0 * v0;
            }
          }
        }
        if (d.rowsq) {
          sq += __shfl_xor_sync(0xffffffffu, sq, 1);
          sq += __shfl_xor_sync(0xffffffffu, sq, 2);
          if (q == 0) red[(wm * 64 + mt * 8 + g) * 4 + wn] = sq;
        }
      }
      if (d.rowsq) {
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
        if (tid < NB) {
          const long long row = p.c_row0 + (long long)tc.ti * NB + tid + bz * p.batch_c_rows;
          if (row < rows_end) {
            const double s4 = (red[tid * 4 + 0] + red[tid * 4 + 1]) + (red[tid * 4 + 2] + red[tid * 4 + 3]);
            p.rowsq[row + bz * p.batch_rowsq] += s4;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
      }
      GM_MARK(2)
      if (DG && p.rhs_r != nullptr) {
        // fused forward substitution: r_i -= L_ik z_k with the tile still in the accumulators (z_k was staged
        // in shared memory before the main loop).  Partial sums per (row, column group) go through shared
        // memory and are added in a fixed order; the two partial-sum buffers alternate between tiles so that
        // one barrier per tile suffices, and the residual is updated with a fire-and-forget reduction (one
        // addition per element and launch, so the result does not depend on timing).
        const int R = p.rhs_R;
        const double* zsm = red;                                   // [128][R]
        double* psm = red + NB * 8 + (rhs_tiles & 1) * (NB * 4 * 8);   // [128][4][R], double-buffered
        rhs_tiles++;
        for (int r = 0; r < R; r++) {
          // this thread's 8 entries of z_k (columns 2q, 2q+1 of each of its 4 sub-tile columns), loaded once;
          // then eight independent accumulation chains, one per sub-tile row
          double zv[4][2], sum[8];
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const int col = wn * 32 + nt * 8 + 2 * q;
            zv[nt][0] = zsm[col * R + r];
            zv[nt][1] = zsm[(col + 1) * R + r];
          }
#pragma unroll
          for (int mt = 0; mt < 8; mt++) sum[mt] = 0.0;
#pragma unroll
          for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int mt = 0; mt < 8; mt++) {
              sum[mt] = fma(acc[mt][nt][0], zv[nt][0], sum[mt]);
              sum[mt] = fma(acc[mt][nt][1], zv[nt][1], sum[mt]);
            }
#pragma unroll
          for (int mt = 0; mt < 8; mt++) sum[mt] += __shfl_xor_sync(0xffffffffu, sum[mt], 1);
#pragma unroll
          for (int mt = 0; mt < 8; mt++) sum[mt] += __shfl_xor_sync(0xffffffffu, sum[mt], 2);
          if (q == 0) {
#pragma unroll
            for (int mt = 0; mt < 8; mt++) psm[((wm * 64 + mt * 8 + g) * 4 + wn) * R + r] = sum[mt];
          }
        }
        GM_MARK(3)
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
        GM_MARK(4)
        if (tid < NB) {
          const long long row = p.rhs_r_row0 + (long long)tc.ti * NB + tid;
          if (row < p.rhs_rows_end) {
            double* rr = p.rhs_r + (row + bz * p.batch_rhs_rows) * R;
            for (int r = 0; r < R; r++)
              atomicAdd(rr + r, -((psm[(tid * 4 + 0) * R + r] + psm[(tid * 4 + 1) * R + r]) +
                                  (psm[(tid * 4 + 2) * R + r] + psm[(tid * 4 + 3) * R + r])));
          }
        }
        GM_MARK(5)
      }
      if (d.signal) {
        // publish this op's global stores to the async proxy (TMA) before the producer loads them; this also
        // tells the producer that the C-tile buffer (operand + row-sum scratch of this op) is free again
        fence_proxy_async();
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dep);
      }
    }
  }
}

#ifdef GPM_GEMM_TIMING
int g_dbg_launch = 0;
extern "C" int gpm_debug_gemm_marks(long long* out) {
  g_dbg_launch = 0;
  return (int)cudaMemcpyFromSymbol(out, g_gemm_marks, sizeof(long long) * 512);
}
#endif
}  // namespace gpm
namespace gpm {
#ifdef GPM_GEMM_TIMING
extern int g_dbg_launch;
#endif

int launch_gemm(gpm_handle_impl* h, const CUtensorMap& mapA, const CUtensorMap& mapB,
                const CUtensorMap& mapC, const GemmArgs& args_in, int batch, cudaStream_t stream,
                const CUtensorMap* mapB2) {
  if (!h->gemm_attr) {
    GPM_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    GPM_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    h->gemm_attr = true;
  }
  GemmArgs args = args_in;
  if (args.sweep_nblk > 0) {
    if (!mapB2 || !args.rowsq || args.tri || args.tiles_n != 1) { set_error("gemm: bad sweep arguments"); return 998; }
    args.klen = NB;
  }
  if (args.klen <= 0 || args.klen % SLAB_K != 0) {
    set_error("gemm: contraction length %d is not a positive multiple of %d", args.klen, SLAB_K);
    return 998;
  }
  if (args.rhs_r && (args.epi != EPI_STORE || args.rowsq || args.sweep_nblk > 0 || args.rhs_R < 1 || args.rhs_R > 8)) {
    set_error("gemm: the fused forward substitution needs a plain EPI_STORE tile launch with 1 <= R <= 8");
    return 998;
  }
  if (args.sweep_nblk == 0 && args.epi == EPI_SUB && args.rowsq) {
    set_error("gemm: rowsq is only supported with EPI_STORE");
    return 998;
  }
  const int total = gemm_grid_x(args);
  if (total <= 0 || batch <= 0) return 0;
  if (gemm_half_eligible(h, args, batch)) return launch_gemm_half(h, args, batch, stream);   // whole-batch launches: half-tiles, two CTAs per SM
  if (gemm_small_eligible(h, args, batch)) return launch_gemm_small(h, args, stream);   // a fraction of a wave: latency kernel
  // tiles per CTA: minimise the makespan ceil(ceil(T/c)/slots)*c over c <= cmax, prefer the larger c
  const int cmax = args.max_tiles_per_cta > 0 ? args.max_tiles_per_cta : 16;
  const long long slots = h->sm_count;
  int best_c = 1;
  long long best_span = -1;
  for (int c = 1; c <= cmax; c++) {
    const long long ctas = ((long long)(total + c - 1) / c) * batch;
    const long long span = ((ctas + slots - 1) / slots) * c;
    if (best_span < 0 || span <= best_span) { best_span = span; best_c = c; }
  }
  if (args.sweep_tri || args.kstart_mode || args.kend_mode) best_c = 1;   // tiles differ in work: let the hardware balance them
  args.tiles_per_cta = best_c;
#ifdef GPM_GEMM_TIMING
  if (args.diag_lower || args.rhs_r) args.dbg_launch = (g_dbg_launch++) % 8;
#endif
  dim3 grid(args.sweep_tri ? (total + 1) / 2 : (total + best_c - 1) / best_c, batch);
  if (args.diag_lower || args.rhs_r)
    gemm_nt_kernel<true><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(mapA, mapB, mapC, mapB2 ? *mapB2 : mapB, args);
  else
    gemm_nt_kernel<false><<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(mapA, mapB, mapC, mapB2 ? *mapB2 : mapB, args);
  GPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace gpm
