// Batched per-path fits for 112 < N <= 1024 (BASELINE config 3: 4096 paths x N = 512): ONE CTA PER PATH, two CTAs
// (two paths) in flight per SM, the whole fit inside one persistent kernel.
//
//   left-looking over 128-wide block columns k:
//     diagonal block   C = K_kk - sum_{j<k} L_kj L_kj^T      DMMA tiles, K generated in registers (never stored)
//                      potf2: L_kk, inv(L_kk), z_k = inv(L_kk) r_k      in shared memory (potf2.cuh)
//     blocks i > k     R = K_ik - sum_{j<k} L_ij L_kj^T      DMMA tiles, result handed over in shared memory
//                      L_ik = R inv(L_kk)^T                  DMMA tiles; L_ik -> global scratch (read back by TMA)
//                      r_i -= L_ik z_k                       forward substitution in the epilogue
//   backward substitution, alpha, log marginal likelihood.
//
// Why this shape.  The tiled batched pipeline (batched.cu) advances all paths together through whole-batch launches
// over an 8.6 GB workspace: every launch streams its operands from HBM (9.2 MB per path) and every 128 x 128 x 128 tile
// pays a store epilogue and a C re-read.  Here a path's covariance is never materialised, the update result goes to
// the solve through shared memory, and what a path writes (its sub-diagonal L tiles and the inverted diagonal
// blocks, 1.3 MB in a per-CTA scratch that is reused for every path of the CTA) is read back while it is still in L2.
// The latency-bound diagonal factorisations of one path overlap the tensor-core phases of the other path on the
// same SM -- in part: DMMA and DFMA share one pipe per scheduler, so the serial FP64 chains of one CTA queue behind the
// other CTA's tile loops (measured: two co-resident paths take 898 us each against 725 us alone).  ncu
// counts 5.0 MB of HBM traffic per path (the 296 scratches do not all stay in L2) against the tiled pipeline's 9.2 MB.
// gpm_fit_batched takes this kernel for batches of at most one wave of CTAs (and when the tiled workspace would pass
// 32 GB); larger batches measure 2 - 6 % faster through the tiled pipeline (batched.cu: use_path_fit).
//
// CTA = 4 warps (tile = 64 rows x 128 columns, warp tile 64 x 32 as 8 x 4 DMMA sub-tiles, 128 accumulator registers
// per thread).  There is no producer warp: ten warps per SM would put three on one scheduler and cap the kernel at
// 168 registers (spills); with eight, every thread may use 255.  Thread 0 drives TMA instead: after its warp has
// finished a slab it waits for the other three warps to release the stage and refills it with the slab two ahead,
// so one slab of compute (>= 2048 cycles) covers each load.  A warp owns the 8-column sub-tile columns
// {w, 7-w, 8+w, 15-w}: the triangular shapes of the path (lower triangle of the diagonal block, the zero slabs of
// the lower-triangular inv(L_kk)) then cost every warp the same number of DMMAs.
// Shared memory: R buffer 64 KB (A-operand slabs of the solve; the packed triangle of potf2 aliases it) + a 2-stage
// TMA ring of (8 KB A + 16 KB B) slabs = 112 KB -> two CTAs per SM.
#include <algorithm>
#include <type_traits>

#include "gemm.cuh"
#include "halftile.cuh"
#include "potf2.cuh"

namespace gpm {

constexpr int PF_CONS_WARPS = 4;
constexpr int PF_CONS = PF_CONS_WARPS * 32;              // 128 threads, all of them consumers
constexpr int PF_THREADS = PF_CONS;
constexpr int PF_STAGES = 2;
constexpr int PF_HALF = 64;                              // rows of a tile
constexpr int PF_A_BYTES = PF_HALF * SLAB_K * 8;         // 8 KB
constexpr int PF_B_BYTES = NB * SLAB_K * 8;              // 16 KB
constexpr int PF_STAGE_BYTES = PF_A_BYTES + PF_B_BYTES;  // 24 KB
constexpr int PF_R_BYTES = PF_HALF * NB * 8;             // 64 KB
constexpr int PF_RING_OFF = PF_R_BYTES;
constexpr int PF_BAR_OFF = PF_RING_OFF + PF_STAGES * PF_STAGE_BYTES;
constexpr int PF_TAB_OFF = PF_BAR_OFF + 64;                // copy of the exp table
constexpr int PF_SMEM = PF_TAB_OFF + GPM_EXP_J * 16;
constexpr int PF_MAX_N = 1024;
static_assert(POTF2_DOUBLES * 8 <= PF_BAR_OFF, "potf2 workspace must fit the R buffer + ring");

struct PathFitArgs {
  const double* Xb; const double* Yb;
  double* alpha; double* lml; int* info;
  long long B;
  int N, D, R, nblk;
  Theta th;
  const double* theta_dev; int theta_stride;
  double* Ls;        // per CTA: Np x Np scratch (sub-diagonal blocks of L), ld = Np
  double* invs;      // per CTA: nblk x 128 x 128 inverted diagonal blocks (upper triangles stay zero)
  double* xs;        // per CTA: Np x 3 coordinates / lengthscale
  double* rs;        // per CTA: Np x R running residual of the forward substitution
  double* zs;        // per CTA: Np x R  z = L^{-1} Y
};

#ifdef GPM_PATHFIT_TIMING
// phase cycles of thread 0 of CTA 0, accumulated over its paths (instrumentation builds only)
__device__ long long g_pf_cycles[16];
#define PF_T(slot) if (tid == 0 && blockIdx.x == 0) { const long long now_ = clock64(); g_pf_cycles[slot] += now_ - t_last; t_last = now_; }
#else
#define PF_T(slot)
#endif

__device__ __forceinline__ void pf_cons_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(PF_CONS) : "memory");
}

// pre-scaled coordinates of one sample / of a thread's eight fragment columns of block column col0 (clamped to N - 1)
template <int D>
__device__ __forceinline__ void pf_load_row(const double* xs, int row, double (&x)[3]) {
  x[0] = xs[row * 3]; x[1] = xs[row * 3 + 1]; x[2] = D == 3 ? xs[row * 3 + 2] : 0.0;
}
template <int D>
__device__ __forceinline__ void pf_load_cols(const double* xs, int col0, const int (&cset)[4], int q, int N, double (&xc)[4][2][3]) {
#pragma unroll
  for (int nt = 0; nt < 4; nt++)
#pragma unroll
    for (int e = 0; e < 2; e++) pf_load_row<D>(xs, min(col0 + cset[nt] * 8 + 2 * q + e, N - 1), xc[nt][e]);
}

// One slab of the diagonal block's update C -= L_k* L_k*^T in ONE pass over the block: both operands are the same
// [128 rows x 16] slab and only the lower triangle is needed, so warp W owns the sub-tiles (mt, nt), mt = 0..15, with
// cset(W, nt) <= mt -- 34 of the 136 for every W -- and the B fragment of sub-tile column c is the A fragment of sub-tile
// row c.  (Two passes over 64-row halves loaded every slab twice and left each with half the DMMAs of a full slab,
// less than the TMA round trip the two-stage ring has to cover.)
template <int W>
__device__ __forceinline__ void pf_slab_diag(double (&acc)[16][4][2], uint32_t s, const uint32_t (&off)[4]) {
#pragma unroll
  for (int k4 = 0; k4 < 4; k4++) {
    double a[16];
#pragma unroll
    for (int mt = W; mt < 16; mt++) a[mt] = lds_f64(s + mt * 1024 + off[k4]);
#pragma unroll
    for (int mt = W; mt < 16; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++)
        if (pf_cset(W, nt) <= mt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], a[pf_cset(W, nt)]);
  }
}

template <int D, int RR>
__global__ void __launch_bounds__(PF_THREADS, 2)
path_fit_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapI, const PathFitArgs p) {
  extern __shared__ __align__(1024) uint8_t pf_smem[];
  const uint32_t base = smem_u32(pf_smem);
  const uint32_t rbuf = base;
  const uint32_t ring = base + PF_RING_OFF;
  const uint32_t bar_full = base + PF_BAR_OFF;            // [2] slab landed (TMA transaction bytes)
  const uint32_t bar_empty = bar_full + PF_STAGES * 8;    // [2] every warp is done with the slab
  double* smd = reinterpret_cast<double*>(pf_smem);       // potf2 workspace / epilogue scratch / solve vectors

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, R = p.R, nblk = p.nblk;
  const int Np = nblk * NB;
  const long long cta = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < PF_STAGES; s++) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, PF_CONS_WARPS);
    }
    fence_mbar_init();
    prefetch_tmap(&mapA); prefetch_tmap(&mapB); prefetch_tmap(&mapI);
  }
  const gpm_exp_pair* etab = reinterpret_cast<const gpm_exp_pair*>(pf_smem + PF_TAB_OFF);
  gpm_exp_stage_table(reinterpret_cast<gpm_exp_pair*>(pf_smem + PF_TAB_OFF), tid, PF_THREADS);
  __syncthreads();
  if ((base & 1023u) != 0) {                               // the 128-byte swizzle pattern needs a 1 KB-aligned window
    if (tid == 0 && blockIdx.x == 0) p.info[0] = -1;
    return;
  }
  const int row_base = (int)(cta * Np);                    // this CTA's first row in the stacked L scratch
  const int inv_base = (int)(cta * nblk * NB);

  // ---- the slab stream.  A SEGMENT is a run of slabs with no dependency on anything the CTA still has to compute:
  //   kind 0, column k: the update of the diagonal block, 8k [128 x 16] slabs of L[k, :] (both operands)
  //   kind 1, column k: the blocks below, 64-row halves t = 0 .. T-1 from row (k+1) 128: 8k update slabs (A and B
  //                     from the L scratch) then 8 solve slabs (B = inv(L_kk) only; A is the R buffer)
  // Slab n of a segment goes to ring stage (global slab count) % 2.  Only thread 0 calls issue().
  auto issue = [&](int kind, int k, int n, int gsl) {
    const int st = gsl % PF_STAGES;
    const uint32_t dst = ring + st * PF_STAGE_BYTES, bar = bar_full + st * 8;
    if (kind == 0) {
      mbar_arrive_expect_tx(bar, PF_B_BYTES);
      tma_load_2d(dst + PF_A_BYTES, &mapB, n * SLAB_K, row_base + k * NB, bar);
    } else {
      const int per = 8 * k + 8, t = n / per, sl = n - t * per;
      if (sl < 8 * k) {
        mbar_arrive_expect_tx(bar, PF_STAGE_BYTES);
        tma_load_2d(dst, &mapA, sl * SLAB_K, row_base + (k + 1) * NB + t * PF_HALF, bar);
        tma_load_2d(dst + PF_A_BYTES, &mapB, sl * SLAB_K, row_base + k * NB, bar);
      } else {
        mbar_arrive_expect_tx(bar, PF_B_BYTES);
        tma_load_2d(dst + PF_A_BYTES, &mapI, (sl - 8 * k) * SLAB_K, inv_base + k * NB, bar);
      }
    }
  };
  int sg = 0;                                              // slabs consumed so far (all threads agree)
  int seg_kind = 0, seg_k = 0, seg_len = 0, seg_pos = 0;
  // start a segment: every warp has passed a CTA barrier since the previous one, so both stages are free
  auto seg_begin = [&](int kind, int k, int len) {
    seg_kind = kind; seg_k = k; seg_len = len; seg_pos = 0;
    if (tid == 0)
      for (int n = 0; n < min(PF_STAGES, len); n++) issue(kind, k, n, sg + n);
  };
  // wait for the next slab; returns the ring stage's shared address
  auto slab_wait = [&]() -> uint32_t {
    const int st = sg % PF_STAGES;
    mbar_wait(bar_full + st * 8, (sg / PF_STAGES) & 1);
    return ring + st * PF_STAGE_BYTES;
  };
  // release the slab; thread 0 refills the stage with the slab two ahead once all four warps have released it
  auto slab_done = [&]() {
    const int st = sg % PF_STAGES;
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + st * 8);
    if (tid == 0 && seg_pos + PF_STAGES < seg_len) {
      mbar_wait(bar_empty + st * 8, (sg / PF_STAGES) & 1);
      issue(seg_kind, seg_k, seg_pos + PF_STAGES, sg + PF_STAGES);
    }
    sg++; seg_pos++;
  };

  // ============================================ the path loop ============================================
  const int w = warp, g = lane >> 2, q = lane & 3;
  uint32_t off[4];
#pragma unroll
  for (int t = 0; t < 4; t++) off[t] = frag_off(g, q, t);
  const int cset[4] = {w, 7 - w, 8 + w, 15 - w};           // this warp's 8-column sub-tile columns (ascending)
  uint32_t boff[4];
#pragma unroll
  for (int nt = 0; nt < 4; nt++) boff[nt] = (uint32_t)cset[nt] * 1024u;

  double* Ls = p.Ls + cta * (long long)Np * Np;
  double* invs = p.invs + cta * (long long)nblk * NB * NB;
  double* xs = p.xs + cta * (long long)Np * 3;
  double* rs = p.rs + cta * (long long)Np * R;
  double* zs = p.zs + cta * (long long)Np * R;

#ifdef GPM_PATHFIT_TIMING
  long long t_last = clock64();
#endif
  for (long long path = cta; path < p.B; path += gridDim.x) {
    Theta th = p.th;
    if (p.theta_dev) {
      const double* t = p.theta_dev + path * p.theta_stride;
#pragma unroll
      for (int d = 0; d < D; d++) th.l[d] = t[d];
      th.sf2 = t[D];
      th.sn2 = t[D + 1];
    }
    const double* X = p.Xb + path * (long long)N * D;
    const double* Y = p.Yb + path * (long long)N * R;
    // eight independent loads in flight per thread (xs / rs may alias X / Y as far as the compiler knows, so a plain
    // load-store loop would pay one HBM round trip per element)
    for (int e0 = 0; e0 < N * D; e0 += 8 * PF_CONS) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { const int e = e0 + u * PF_CONS + tid; v[u] = e < N * D ? X[e] : 0.0; }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int e = e0 + u * PF_CONS + tid;
        const int i = e / D, d = e - i * D;
        if (e < N * D) xs[i * 3 + d] = v[u] / (d == 0 ? th.l[0] : (d == 1 ? th.l[1] : th.l[2]));
      }
    }
    for (int e0 = 0; e0 < N * R; e0 += 8 * PF_CONS) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { const int e = e0 + u * PF_CONS + tid; v[u] = e < N * R ? Y[e] : 0.0; }
#pragma unroll
      for (int u = 0; u < 8; u++) { const int e = e0 + u * PF_CONS + tid; if (e < N * R) rs[e] = v[u]; }
    }
    double logdet = 0.0;
    pf_cons_sync();
    PF_T(0)

    for (int k = 0; k < nblk; k++) {
      const int nv = min(NB, N - k * NB);
      // ------------------------------ diagonal block: update, then potf2 ------------------------------
      seg_begin(0, k, 8 * k);
      auto diag_block = [&](auto wtag) {
        constexpr int W = decltype(wtag)::value;
        double acc[16][4][2];
#pragma unroll
        for (int mt = 0; mt < 16; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        for (int s = 0; s < 8 * k; s++) {
          const uint32_t sb = slab_wait() + PF_A_BYTES;
          pf_slab_diag<W>(acc, sb, off);
          slab_done();
        }
        // the packed triangle's last tiles alias the first ring stage: every warp must be past its last slab
        pf_cons_sync();
        PF_T(1)
        // pass 1: the update (zero for k = 0) at this thread's fragment positions of the lower triangle
#pragma unroll
        for (int mt = W; mt < 16; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++)
            if (pf_cset(W, nt) <= mt)
              *reinterpret_cast<double2*>(smd + toff(mt * 8 + g, pf_cset(W, nt) * 8 + 2 * q)) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        // pass 2: K_kk (+ noise on the diagonal, identity beyond N) minus the update, in place (each thread re-reads
        // only what it wrote itself).  Rolled over the sub-tile rows (unrolled kernel evaluations would be > 100 KB of
        // code); the column coordinates are loaded once, the row coordinates one iteration ahead, the exp table sits
        // in shared memory, so the (up to) eight evaluations of an iteration overlap each other's latencies.
        double xc[4][2][3];
        pf_load_cols<D>(xs, k * NB, cset, q, N, xc);
        double xn[3];
        pf_load_row<D>(xs, min(k * NB + W * 8 + g, N - 1), xn);
#pragma unroll 1
        for (int mt = W; mt < 16; mt++) {
          const int il = mt * 8 + g;                                     // row inside the block
          const int row = k * NB + il;
          const double xr[3] = {xn[0], xn[1], xn[2]};
          if (mt < 15) pf_load_row<D>(xs, min(row + 8, N - 1), xn);
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            if (pf_cset(W, nt) > mt) continue;                           // above the diagonal: not stored
            const int cl = pf_cset(W, nt) * 8 + 2 * q;
            double2* ptr = reinterpret_cast<double2*>(smd + toff(il, cl));
            const double2 u = *ptr;
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int cc = k * NB + cl + e;
              const double kv = rbf_t<D>(xr, xc[nt][e], th.sf2, etab);
              v[e] = (row >= N || cc >= N) ? ((il == cl + e) ? 1.0 : 0.0) : ((row == cc ? kv + th.sn2 : kv) - (e ? u.y : u.x));
            }
            *ptr = make_double2(v[0], v[1]);
          }
        }
        PF_T(2)
      };
      switch (w) {                                           // warp-uniform: the lower-triangle shape of this warp's columns
        case 0: diag_block(std::integral_constant<int, 0>{}); break;
        case 1: diag_block(std::integral_constant<int, 1>{}); break;
        case 2: diag_block(std::integral_constant<int, 2>{}); break;
        default: diag_block(std::integral_constant<int, 3>{}); break;
      }
      pf_cons_sync();
      potf2_factor<PF_CONS, 1>(smd, tid, nv, (long long)k * NB, p.info + path);
      if (tid < nv) logdet += log(smd[toff(tid, tid)]);
      pf_cons_sync();
      PF_T(3)
      potf2_invert<PF_CONS, 1>(smd, tid);
      PF_T(4)
      potf2_fwd_z<PF_CONS, 1>(smd, tid, nv, rs + (long long)k * NB * R, zs + (long long)k * NB * R, R);
      PF_T(5)
      {
        // inv(L_kk): the 8 x 8 tiles on or below the diagonal go to the scratch (the upper triangle was zeroed once)
        double* dst = invs + (long long)k * NB * NB;
        for (int idx = tid; idx < NB * NB / 2; idx += PF_CONS) {
          const int i = idx >> 6, c = (idx & 63) * 2;
          if ((c >> 3) <= (i >> 3)) *reinterpret_cast<double2*>(dst + i * NB + c) = *reinterpret_cast<const double2*>(smd + toff(i, c));
        }
      }
      fence_proxy_async();                                 // inv(L_kk) was written through the generic proxy, TMA reads it
      __threadfence_block();
      pf_cons_sync();                                      // potf2's shared memory is free; z_k is visible to the CTA
      {
        const int rows_below = N - (k + 1) * NB;
        seg_begin(1, k, rows_below > 0 ? ((rows_below + PF_HALF - 1) / PF_HALF) * (8 * k + 8) : 0);
      }
      PF_T(6)

      // ------------------------------ blocks below: update, solve, forward substitution ------------------------------
      for (int i = k + 1; i < nblk; i++) {
        for (int h = 0; h < 2; h++) {
          const int row0 = i * NB + h * PF_HALF;
          if (row0 >= N) continue;
          // K_ik at this thread's fragment positions -> R buffer, in the swizzled layout of A-operand slabs ([64 rows
          // x 16] boxes); rolled over the sub-tile rows.  It runs while the first slabs of the tile are in flight.
          {
            double xc[4][2][3];
            pf_load_cols<D>(xs, k * NB, cset, q, N, xc);
            double xn[3];
            pf_load_row<D>(xs, min(row0 + g, N - 1), xn);
#pragma unroll 1
            for (int mt = 0; mt < 8; mt++) {
              const int il = mt * 8 + g;
              const int row = row0 + il;
              const double xr[3] = {xn[0], xn[1], xn[2]};
              if (mt < 7) pf_load_row<D>(xs, min(row + 8, N - 1), xn);
#pragma unroll
              for (int nt = 0; nt < 4; nt++) {
                const int c = cset[nt];                                   // columns < (k + 1) 128 <= row0: always valid samples
                double v0 = rbf_t<D>(xr, xc[nt][0], th.sf2, etab), v1 = rbf_t<D>(xr, xc[nt][1], th.sf2, etab);
                if (row >= N) v0 = v1 = 0.0;
                *reinterpret_cast<double2*>(pf_smem + (c >> 1) * PF_A_BYTES + il * 128 + ((((c & 1) * 4 + q) ^ g) << 4)) = make_double2(v0, v1);
              }
            }
          }
          PF_T(7)
          double acc[8][4][2];
#pragma unroll
          for (int mt = 0; mt < 8; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
          for (int s = 0; s < 8 * k; s++) {
            const uint32_t sa = slab_wait(), sb = sa + PF_A_BYTES;
            pf_slab<PF_FULL>(acc, sa, sb, off, boff);
            slab_done();
          }
          PF_T(8)
          // R = K_ik - acc, in place (rows beyond N stay zero)
          if (k > 0) {
#pragma unroll
            for (int mt = 0; mt < 8; mt++) {
              const int il = mt * 8 + g;
              if (row0 + il < N) {
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                  const int c = cset[nt];
                  const uint32_t addr = rbuf + (c >> 1) * PF_A_BYTES + il * 128 + ((((c & 1) * 4 + q) ^ g) << 4);
                  double v0, v1;
                  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr));
                  v0 -= acc[mt][nt][0]; v1 -= acc[mt][nt][1];
                  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(v0), "d"(v1) : "memory");
                }
              }
            }
          }
          pf_cons_sync();
          // L_ik = R inv(L_kk)^T: A from the R buffer, B = inv(L_kk) slabs; sub-tile column c needs slab s iff 2 s <= c
#pragma unroll
          for (int mt = 0; mt < 8; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
          for (int s = 0; s < NB / SLAB_K; s++) {
            const uint32_t sa = rbuf + s * PF_A_BYTES, sb = slab_wait() + PF_A_BYTES;
            int dead = 0;                                   // cset is ascending: the live columns are a suffix
#pragma unroll
            for (int nt = 0; nt < 4; nt++) dead += (cset[nt] < 2 * s) ? 1 : 0;
            switch (dead) {
              case 0: pf_slab<PF_COLS + 0>(acc, sa, sb, off, boff); break;
              case 1: pf_slab<PF_COLS + 1>(acc, sa, sb, off, boff); break;
              case 2: pf_slab<PF_COLS + 2>(acc, sa, sb, off, boff); break;
              case 3: pf_slab<PF_COLS + 3>(acc, sa, sb, off, boff); break;
              default: break;
            }
            slab_done();
          }
          PF_T(9)
          // epilogue: L_ik -> scratch (16-byte stores), partial sums of L_ik z_k per (row, warp)
#pragma unroll
          for (int mt = 0; mt < 8; mt++) {
            const int row = row0 + mt * 8 + g;
            if (row < N) {
              double* lrow = Ls + (long long)row * Np + k * NB;
#pragma unroll
              for (int nt = 0; nt < 4; nt++)
                *reinterpret_cast<double2*>(lrow + cset[nt] * 8 + 2 * q) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            }
          }
          pf_cons_sync();                                   // every warp is past its last read of the R buffer
          double* psm = smd;                                // [64][4][R] partial sums (aliases the R buffer)
          const double* zk = zs + (long long)k * NB * R;
#pragma unroll
          for (int r = 0; r < RR; r++) {
            if (r >= R) break;
            double zv[4][2], sum[8];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
              const int cl = cset[nt] * 8 + 2 * q;
              zv[nt][0] = zk[cl * R + r];
              zv[nt][1] = zk[(cl + 1) * R + r];
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
              for (int mt = 0; mt < 8; mt++) psm[((mt * 8 + g) * 4 + w) * R + r] = sum[mt];
            }
          }
          pf_cons_sync();
          if (tid < PF_HALF && row0 + tid < N) {
            double* rr = rs + (long long)(row0 + tid) * R;
            for (int r = 0; r < R; r++)
              rr[r] -= (psm[(tid * 4 + 0) * R + r] + psm[(tid * 4 + 1) * R + r]) +
                       (psm[(tid * 4 + 2) * R + r] + psm[(tid * 4 + 3) * R + r]);
          }
          pf_cons_sync();                                   // psm is free again (the next tile writes the R buffer)
          PF_T(10)
        }
      }
      fence_proxy_async();                                 // this column's L tiles: generic-proxy stores, read back by TMA
      __threadfence_block();
      pf_cons_sync();
    }

    // ------------------------------ backward substitution L^T alpha = z, alpha out, LML ------------------------------
    PF_T(11)
    double* zsm = smd;                                      // [Np][R]
    double* ys = zsm + (long long)Np * R;                   // [128][R]
    double* red = ys + NB * R;                              // [4][RR + 1]
    for (int e = tid; e < Np * R; e += PF_CONS) zsm[e] = (e / R < N) ? __ldcg(zs + e) : 0.0;
    pf_cons_sync();
    for (int i = nblk - 1; i >= 0; i--) {
      const int i0 = i * NB;
      double accr[RR];
#pragma unroll
      for (int r = 0; r < RR; r++) accr[r] = 0.0;
      for (int j = nblk - 1; j > i; j--) {                   // column tid of block (j, i): rows j0 .. j0+127
        const int j0 = j * NB;
        for (int rb = 0; rb < NB; rb += 32) {
          double seg[32];
#pragma unroll
          for (int r2 = 0; r2 < 32; r2++) {
            const int gr = j0 + rb + r2;
            seg[r2] = (gr < N) ? __ldcg(Ls + (long long)gr * Np + i0 + tid) : 0.0;
          }
#pragma unroll
          for (int r2 = 0; r2 < 32; r2++) {
            const double* zr = zsm + (long long)(j0 + rb + r2) * R;
#pragma unroll
            for (int r = 0; r < RR; r++)
              if (r < R) accr[r] = fma(seg[r2], zr[r], accr[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RR; r++)
        if (r < R) ys[tid * R + r] = zsm[(long long)(i0 + tid) * R + r] - accr[r];
      pf_cons_sync();
      // a_i[tid] = sum_{row >= tid} inv(L_ii)[row][tid] y[row]   (inv lower triangular: rows below the column)
#pragma unroll
      for (int r = 0; r < RR; r++) accr[r] = 0.0;
      const double* Di = invs + (long long)i * NB * NB;
      for (int rb = (tid & ~31); rb < NB; rb += 32) {
        double seg[32];
#pragma unroll
        for (int r2 = 0; r2 < 32; r2++) seg[r2] = __ldcg(Di + (rb + r2) * NB + tid);
#pragma unroll
        for (int r2 = 0; r2 < 32; r2++) {
          const double* yr = ys + (rb + r2) * R;
#pragma unroll
          for (int r = 0; r < RR; r++)
            if (r < R) accr[r] = fma(seg[r2], yr[r], accr[r]);
        }
      }
      pf_cons_sync();                                        // every thread has read ys and its z_i entries
#pragma unroll
      for (int r = 0; r < RR; r++)
        if (r < R) zsm[(long long)(i0 + tid) * R + r] = (i0 + tid < N) ? accr[r] : 0.0;
      pf_cons_sync();
    }
    // alpha out; lml[r] = -1/2 y^T alpha - sum log L_ii - N/2 log(2 pi)
    double part[RR + 1];
#pragma unroll
    for (int r = 0; r < RR; r++) part[r] = 0.0;
    part[RR] = logdet;
    double* al = p.alpha + path * (long long)N * R;
    for (int i = tid; i < N; i += PF_CONS) {
#pragma unroll
      for (int r = 0; r < RR; r++)
        if (r < R) {
          const double a = zsm[(long long)i * R + r];
          al[i * R + r] = a;
          part[r] = fma(Y[i * R + r], a, part[r]);
        }
    }
    if (p.lml) {
#pragma unroll
      for (int r = 0; r <= RR; r++) {
        const double sv = warp_sum(part[r]);
        if (lane == 0) red[w * (RR + 1) + r] = sv;
      }
      pf_cons_sync();
      if (tid < R) {
        double qf = 0.0, ld = 0.0;
        for (int ww = 0; ww < PF_CONS_WARPS; ww++) { qf += red[ww * (RR + 1) + tid]; ld += red[ww * (RR + 1) + RR]; }
        p.lml[path * R + tid] = -0.5 * qf - ld - 0.5 * (double)N * 1.8378770664093454835606594728112;
      }
    }
    pf_cons_sync();                                          // the solve's shared memory is free for the next path
    PF_T(12)
  }
}

#ifdef GPM_PATHFIT_TIMING
extern "C" int gpm_debug_pathfit_cycles(long long* out, int reset) {
  if (reset) { long long z[16] = {0}; return (int)cudaMemcpyToSymbol(g_pf_cycles, z, sizeof(z)); }
  return (int)cudaMemcpyFromSymbol(out, g_pf_cycles, sizeof(long long) * 16);
}
#endif

bool path_fit_supported(long long N, int R) { return N > 112 && N <= PF_MAX_N && R >= 1 && R <= 8; }

size_t path_fit_workspace_bytes(const gpm_handle_impl* h, long long B, long long N) {
  const long long np = (N + NB - 1) / NB * NB, nblk = np / NB;
  const long long ctas = std::min<long long>(B, 2ll * h->sm_count);
  return (size_t)(ctas * (np * np + nblk * NB * NB + np * 3 + 2 * np * 8) + B * 8) * sizeof(double);
}

int launch_path_fit(gpm_handle_impl* h, const double* Xb, const double* Yb, long long B, long long N, int D, int R,
                    const Theta& th, const double* theta_host, int theta_stride, double* alpha, double* lml,
                    int* info, void* ws, cudaStream_t st) {
  const long long np = (N + NB - 1) / NB * NB;
  const int nblk = (int)(np / NB);
  const long long ctas = std::min<long long>(B, 2ll * h->sm_count);
  double* Ls = reinterpret_cast<double*>(ws);
  double* invs = Ls + ctas * np * np;
  double* xs = invs + ctas * nblk * NB * NB;
  double* rs = xs + ctas * np * 3;
  double* zs = rs + ctas * np * 8;
  double* tdev = zs + ctas * np * 8;
  PathFitArgs a = {};
  a.Xb = Xb; a.Yb = Yb; a.alpha = alpha; a.lml = lml; a.info = info;
  a.B = B; a.N = (int)N; a.D = D; a.R = R; a.nblk = nblk; a.th = th;
  a.Ls = Ls; a.invs = invs; a.xs = xs; a.rs = rs; a.zs = zs;
  if (theta_stride) {
    GPM_CUDA(cudaMemcpyAsync(tdev, theta_host, (size_t)B * (D + 2) * sizeof(double), cudaMemcpyHostToDevice, st));
    a.theta_dev = tdev; a.theta_stride = theta_stride;
  }
  // the strict upper triangles of the inverse blocks are never written: zero them once per call
  GPM_CUDA(cudaMemsetAsync(invs, 0, (size_t)ctas * nblk * NB * NB * sizeof(double), st));
  GPM_CUDA(cudaMemsetAsync(info, 0, (size_t)B * sizeof(int), st));
  CUtensorMap mapA, mapB, mapI;
  int rc;
  if ((rc = make_tmap(h, &mapA, Ls, ctas * np, np, np, PF_HALF))) return rc;
  if ((rc = make_tmap(h, &mapB, Ls, ctas * np, np, np, NB))) return rc;
  if ((rc = make_tmap(h, &mapI, invs, ctas * nblk * NB, NB, NB, NB))) return rc;
  auto launch = [&](auto kern) -> int {
    GPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM));
    GPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    kern<<<(unsigned)ctas, PF_THREADS, PF_SMEM, st>>>(mapA, mapB, mapI, a);
    return 0;
  };
  const int RRsel = R <= 1 ? 1 : (R <= 2 ? 2 : (R <= 4 ? 4 : 8));
  if (D == 2) rc = RRsel == 1 ? launch(path_fit_kernel<2, 1>) : RRsel == 2 ? launch(path_fit_kernel<2, 2>)
                 : RRsel == 4 ? launch(path_fit_kernel<2, 4>) : launch(path_fit_kernel<2, 8>);
  else rc = RRsel == 1 ? launch(path_fit_kernel<3, 1>) : RRsel == 2 ? launch(path_fit_kernel<3, 2>)
            : RRsel == 4 ? launch(path_fit_kernel<3, 4>) : launch(path_fit_kernel<3, 8>);
  if (rc) return rc;
  GPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace gpm
