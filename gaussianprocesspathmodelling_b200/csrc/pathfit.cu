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
// same SM: that is what the hardware's warp schedulers do with two resident CTAs, no software pipelining needed.
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

#include "gemm.cuh"
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
constexpr int PF_SMEM = PF_BAR_OFF + 64;
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

__device__ __forceinline__ void pf_cons_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(PF_CONS) : "memory");
}

// One slab (16 contraction steps) of a warp's 64 x 32 tile with a run-time mask of the 8 x 4 sub-tiles (bit mt*4+nt)
// and per-column-group B offsets (the warp's sub-tile columns are not contiguous).
__device__ __forceinline__ void pf_slab_mma(double (&acc)[8][4][2], uint32_t sa, uint32_t sb, const uint32_t (&off)[4],
                                            const uint32_t (&boff)[4], uint32_t mask) {
#pragma unroll
  for (int k4 = 0; k4 < 4; k4++) {
    double a[8], b[4];
#pragma unroll
    for (int mt = 0; mt < 8; mt++) a[mt] = lds_f64(sa + mt * 1024 + off[k4]);
#pragma unroll
    for (int nt = 0; nt < 4; nt++) b[nt] = lds_f64(sb + boff[nt] + off[k4]);
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++)
        if (mask >> (mt * 4 + nt) & 1u) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}
__device__ __forceinline__ void pf_slab_mma_full(double (&acc)[8][4][2], uint32_t sa, uint32_t sb, const uint32_t (&off)[4],
                                                 const uint32_t (&boff)[4]) {
#pragma unroll
  for (int k4 = 0; k4 < 4; k4++) {
    double a[8], b[4];
#pragma unroll
    for (int mt = 0; mt < 8; mt++) a[mt] = lds_f64(sa + mt * 1024 + off[k4]);
#pragma unroll
    for (int nt = 0; nt < 4; nt++) b[nt] = lds_f64(sb + boff[nt] + off[k4]);
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}

template <int D>
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
  __syncthreads();
  if ((base & 1023u) != 0) {                               // the 128-byte swizzle pattern needs a 1 KB-aligned window
    if (tid == 0 && blockIdx.x == 0) p.info[0] = -1;
    return;
  }
  const int row_base = (int)(cta * Np);                    // this CTA's first row in the stacked L scratch
  const int inv_base = (int)(cta * nblk * NB);

  // ---- the slab stream.  A SEGMENT is a run of slabs with no dependency on anything the CTA still has to compute:
  //   kind 0, column k: the update of the diagonal block, live 64-row halves t = 0, 1, 8k slabs each
  //   kind 1, column k: the blocks below, 64-row halves t = 0 .. T-1 from row (k+1) 128: 8k update slabs (A and B
  //                     from the L scratch) then 8 solve slabs (B = inv(L_kk) only; A is the R buffer)
  // Slab n of a segment goes to ring stage (global slab count) % 2.  Only thread 0 calls issue().
  auto issue = [&](int kind, int k, int n, int gsl) {
    const int st = gsl % PF_STAGES;
    const uint32_t dst = ring + st * PF_STAGE_BYTES, bar = bar_full + st * 8;
    if (kind == 0) {
      const int per = 8 * k, t = n / per, sl = n - t * per;
      mbar_arrive_expect_tx(bar, PF_STAGE_BYTES);
      tma_load_2d(dst, &mapA, sl * SLAB_K, row_base + k * NB + t * PF_HALF, bar);
      tma_load_2d(dst + PF_A_BYTES, &mapB, sl * SLAB_K, row_base + k * NB, bar);
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

  // ============================================ consumers ============================================
  const int w = warp, g = lane >> 2, q = lane & 3;
  uint32_t off[4];
#pragma unroll
  for (int t = 0; t < 4; t++) off[t] = frag_off(g, q, t);
  const int cset[4] = {w, 7 - w, 8 + w, 15 - w};           // this warp's 8-column sub-tile columns
  uint32_t boff[4];
#pragma unroll
  for (int nt = 0; nt < 4; nt++) boff[nt] = (uint32_t)cset[nt] * 1024u;
  // sub-tile masks of the two halves of a diagonal block: (mt, c) is on or below the diagonal iff c <= mt + 8 h
  uint32_t dmask[2] = {0u, 0u};
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++)
        if (cset[nt] <= mt + 8 * h) dmask[h] |= 1u << (mt * 4 + nt);

  double* Ls = p.Ls + cta * (long long)Np * Np;
  double* invs = p.invs + cta * (long long)nblk * NB * NB;
  double* xs = p.xs + cta * (long long)Np * 3;
  double* rs = p.rs + cta * (long long)Np * R;
  double* zs = p.zs + cta * (long long)Np * R;

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
    for (int e = tid; e < N * D; e += PF_CONS) {
      const int i = e / D, d = e - i * D;
      xs[i * 3 + d] = X[e] / (d == 0 ? th.l[0] : (d == 1 ? th.l[1] : th.l[2]));
    }
    for (int e = tid; e < N * R; e += PF_CONS) rs[e] = Y[e];
    double logdet = 0.0;
    pf_cons_sync();

    // K(row, col) + noise on the diagonal, identity padding beyond N; same arithmetic as cov_kernel
    auto kval = [&](int row, int col, const double* xr) -> double {
      if (row >= N || col >= N) return row == col ? 1.0 : 0.0;
      const double* xc = xs + col * 3;
      double v = rbf<D>(xr, xc, th.sf2);
      if (row == col) v += th.sn2;
      return v;
    };

    for (int k = 0; k < nblk; k++) {
      const int nv = min(NB, N - k * NB);
      // ------------------------------ diagonal block: update, then potf2 ------------------------------
      seg_begin(0, k, 8 * k * min(2, (nv + PF_HALF - 1) / PF_HALF));
      for (int h = 0; h < 2; h++) {
        const bool live = k * NB + h * PF_HALF < N;
        double acc[8][4][2];
#pragma unroll
        for (int mt = 0; mt < 8; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        if (live) {
          for (int s = 0; s < 8 * k; s++) {
            const uint32_t sa = slab_wait(), sb = sa + PF_A_BYTES;
            pf_slab_mma(acc, sa, sb, off, boff, dmask[h]);
            slab_done();
          }
        }
        // the packed triangle's last tiles alias the first ring stage: every warp must be past its last slab
        if (h == 1) pf_cons_sync();
#pragma unroll
        for (int mt = 0; mt < 8; mt++) {
          const int il = h * PF_HALF + mt * 8 + g;                     // row inside the block
          const int row = k * NB + il;
          const double xr[3] = {xs[row * 3], xs[row * 3 + 1], D == 3 ? xs[row * 3 + 2] : 0.0};
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            if (!(dmask[h] >> (mt * 4 + nt) & 1u)) continue;
            const int cl = cset[nt] * 8 + 2 * q;
            const int col = k * NB + cl;
            const double v0 = kval(row, col, xr) - acc[mt][nt][0];
            const double v1 = kval(row, col + 1, xr) - acc[mt][nt][1];
            const bool pad = row >= N;
            *reinterpret_cast<double2*>(smd + toff(il, cl)) =
                make_double2(pad ? (il == cl ? 1.0 : 0.0) : (col >= N ? 0.0 : v0),
                             pad ? (il == cl + 1 ? 1.0 : 0.0) : (col + 1 >= N ? 0.0 : v1));
          }
        }
      }
      pf_cons_sync();
      potf2_factor<PF_CONS, 1>(smd, tid, nv, (long long)k * NB, p.info + path);
      if (tid < nv) logdet += log(smd[toff(tid, tid)]);
      pf_cons_sync();
      potf2_invert<PF_CONS, 1>(smd, tid);
      potf2_fwd_z<PF_CONS, 1>(smd, tid, nv, rs + (long long)k * NB * R, zs + (long long)k * NB * R, R);
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

      // ------------------------------ blocks below: update, solve, forward substitution ------------------------------
      for (int i = k + 1; i < nblk; i++) {
        for (int h = 0; h < 2; h++) {
          const int row0 = i * NB + h * PF_HALF;
          if (row0 >= N) continue;
          double acc[8][4][2];
#pragma unroll
          for (int mt = 0; mt < 8; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
          for (int s = 0; s < 8 * k; s++) {
            const uint32_t sa = slab_wait(), sb = sa + PF_A_BYTES;
            pf_slab_mma_full(acc, sa, sb, off, boff);
            slab_done();
          }
          // R = K_ik - acc into the R buffer, in the swizzled layout of A-operand slabs ([64 rows x 16] boxes)
#pragma unroll
          for (int mt = 0; mt < 8; mt++) {
            const int il = mt * 8 + g;
            const int row = row0 + il;
            const double xr[3] = {xs[min(row, N - 1) * 3], xs[min(row, N - 1) * 3 + 1], D == 3 ? xs[min(row, N - 1) * 3 + 2] : 0.0};
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
              const int c = cset[nt];
              const int col = k * NB + c * 8 + 2 * q;
              double v0 = 0.0, v1 = 0.0;
              if (row < N) {
                v0 = kval(row, col, xr) - acc[mt][nt][0];
                v1 = kval(row, col + 1, xr) - acc[mt][nt][1];
              }
              const uint32_t addr = rbuf + (c >> 1) * PF_A_BYTES + il * 128 + ((((c & 1) * 4 + q) ^ g) << 4);
              asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(v0), "d"(v1) : "memory");
            }
          }
          pf_cons_sync();
          // L_ik = R inv(L_kk)^T: A from the R buffer, B = inv(L_kk) slabs; slab s is all zero for sub-tile column c > 2s+1
#pragma unroll
          for (int mt = 0; mt < 8; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
          for (int s = 0; s < NB / SLAB_K; s++) {
            const uint32_t sa = rbuf + s * PF_A_BYTES, sb = slab_wait() + PF_A_BYTES;
            uint32_t m4 = 0;                                // sub-tile column c needs slab s iff 16 s <= 8 c + 7
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
              if (2 * s <= cset[nt]) m4 |= 1u << nt;
            pf_slab_mma(acc, sa, sb, off, boff, m4 * 0x11111111u);
            slab_done();
          }
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
          for (int r = 0; r < R; r++) {
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
        }
      }
      fence_proxy_async();                                 // this column's L tiles: generic-proxy stores, read back by TMA
      __threadfence_block();
      pf_cons_sync();
    }

    // ------------------------------ backward substitution L^T alpha = z, alpha out, LML ------------------------------
    pf_cons_sync();
    double* zsm = smd;                                      // [Np][R]
    double* ys = zsm + (long long)Np * R;                   // [128][R]
    double* red = ys + NB * R;                              // [4][R + 1]
    for (int e = tid; e < Np * R; e += PF_CONS) zsm[e] = (e / R < N) ? __ldcg(zs + e) : 0.0;
    pf_cons_sync();
    for (int i = nblk - 1; i >= 0; i--) {
      const int i0 = i * NB;
      double accr[8];
#pragma unroll
      for (int r = 0; r < 8; r++) accr[r] = 0.0;
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
            for (int r = 0; r < 8; r++)
              if (r < R) accr[r] = fma(seg[r2], zr[r], accr[r]);
          }
        }
      }
      for (int r = 0; r < R; r++) ys[tid * R + r] = zsm[(long long)(i0 + tid) * R + r] - accr[r];
      pf_cons_sync();
      // a_i[tid] = sum_{row >= tid} inv(L_ii)[row][tid] y[row]   (inv lower triangular: rows below the column)
#pragma unroll
      for (int r = 0; r < 8; r++) accr[r] = 0.0;
      const double* Di = invs + (long long)i * NB * NB;
      for (int rb = (tid & ~31); rb < NB; rb += 32) {
        double seg[32];
#pragma unroll
        for (int r2 = 0; r2 < 32; r2++) seg[r2] = __ldcg(Di + (rb + r2) * NB + tid);
#pragma unroll
        for (int r2 = 0; r2 < 32; r2++) {
          const double* yr = ys + (rb + r2) * R;
#pragma unroll
          for (int r = 0; r < 8; r++)
            if (r < R) accr[r] = fma(seg[r2], yr[r], accr[r]);
        }
      }
      pf_cons_sync();                                        // every thread has read ys and its z_i entries
      for (int r = 0; r < R; r++) zsm[(long long)(i0 + tid) * R + r] = (i0 + tid < N) ? accr[r] : 0.0;
      pf_cons_sync();
    }
    // alpha out; lml[r] = -1/2 y^T alpha - sum log L_ii - N/2 log(2 pi)
    double part[9];
#pragma unroll
    for (int r = 0; r < 8; r++) part[r] = 0.0;
    part[8] = logdet;
    double* al = p.alpha + path * (long long)N * R;
    for (int i = tid; i < N; i += PF_CONS)
      for (int r = 0; r < R; r++) {
        const double a = zsm[(long long)i * R + r];
        al[i * R + r] = a;
        part[r] = fma(Y[i * R + r], a, part[r]);
      }
    if (p.lml) {
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const double sv = warp_sum(part[r]);
        if (lane == 0) red[w * 9 + r] = sv;
      }
      pf_cons_sync();
      if (tid < R) {
        double qf = 0.0, ld = 0.0;
        for (int ww = 0; ww < PF_CONS_WARPS; ww++) { qf += red[ww * 9 + tid]; ld += red[ww * 9 + 8]; }
        p.lml[path * R + tid] = -0.5 * qf - ld - 0.5 * (double)N * 1.8378770664093454835606594728112;
      }
    }
    pf_cons_sync();                                          // the solve's shared memory is free for the next path
  }
}

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
  if (!h->pathfit_attr) {
    GPM_CUDA(cudaFuncSetAttribute(path_fit_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM));
    GPM_CUDA(cudaFuncSetAttribute(path_fit_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM));
    GPM_CUDA(cudaFuncSetAttribute(path_fit_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    GPM_CUDA(cudaFuncSetAttribute(path_fit_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    h->pathfit_attr = true;
  }
  if (D == 2) path_fit_kernel<2><<<(unsigned)ctas, PF_THREADS, PF_SMEM, st>>>(mapA, mapB, mapI, a);
  else path_fit_kernel<3><<<(unsigned)ctas, PF_THREADS, PF_SMEM, st>>>(mapA, mapB, mapI, a);
  GPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace gpm
