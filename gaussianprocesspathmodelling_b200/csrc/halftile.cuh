// 64 x 128 half-tiles on four warps: the DMMA slab step shared by the one-CTA-per-path fit (pathfit.cu) and the
// batched half-tile GEMM (gemm_half.cu).
#pragma once
#include "common.cuh"

namespace gpm {

// A warp owns the 8-column sub-tile columns cset(w) = {w, 7 - w, 8 + w, 15 - w} (ascending).  One slab (16 contraction
// steps) of its 64 x 32 tile is 8 x 4 DMMA sub-tiles per k4 step; which of them are needed is a compile-time shape
// (mma.sync is warp-collective, so a run-time mask would put a branch around every DMMA):
//   PF_FULL            all 32
//   PF_COLS + n        sub-tile columns nt >= n   (solve against the lower-triangular inv(L_kk): column c needs slab s
//                      iff 2 s <= c, and cset is ascending, so the live columns are a suffix)
//   PF_DIAG0/1 + 2 w   half h of a diagonal block for warp w: (mt, nt) is on or below the diagonal iff cset[nt] <= mt + 8 h
enum { PF_FULL = 0, PF_COLS = 1, PF_DIAG = 8 };
__host__ __device__ constexpr int pf_cset(int w, int nt) { return nt == 0 ? w : (nt == 1 ? 7 - w : (nt == 2 ? 8 + w : 15 - w)); }
template <int SHAPE>
__host__ __device__ constexpr bool pf_live(int mt, int nt) {
  if (SHAPE == PF_FULL) return true;
  if (SHAPE >= PF_COLS && SHAPE < PF_DIAG) return nt >= SHAPE - PF_COLS;
  const int w = (SHAPE - PF_DIAG) >> 1, h = (SHAPE - PF_DIAG) & 1;
  return pf_cset(w, nt) <= mt + 8 * h;
}
template <int SHAPE>
__host__ __device__ constexpr bool pf_row_live(int mt) { return pf_live<SHAPE>(mt, 0) || pf_live<SHAPE>(mt, 1) || pf_live<SHAPE>(mt, 2) || pf_live<SHAPE>(mt, 3); }
template <int SHAPE>
__host__ __device__ constexpr bool pf_col_live(int nt) {
  for (int mt = 0; mt < 8; mt++) if (pf_live<SHAPE>(mt, nt)) return true;
  return false;
}

// MT0, NMT: the sub-tile rows [MT0, MT0 + NMT) of the 64-row half-tile this warp owns (all eight with four warps per
// half-tile; four with eight warps, gemm_half.cu); acc is indexed by the local row mt - MT0.
template <int SHAPE, int MT0 = 0, int NMT = 8>
__device__ __forceinline__ void pf_slab(double (&acc)[NMT][4][2], uint32_t sa, uint32_t sb, const uint32_t (&off)[4],
                                        const uint32_t (&boff)[4]) {
#pragma unroll
  for (int k4 = 0; k4 < 4; k4++) {
    double a[NMT], b[4];
#pragma unroll
    for (int mt = 0; mt < NMT; mt++)
      if (pf_row_live<SHAPE>(MT0 + mt)) a[mt] = lds_f64(sa + (MT0 + mt) * 1024 + off[k4]);
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      bool live = false;
#pragma unroll
      for (int mt = 0; mt < NMT; mt++) live = live || pf_live<SHAPE>(MT0 + mt, nt);
      if (live) b[nt] = lds_f64(sb + boff[nt] + off[k4]);
    }
#pragma unroll
    for (int mt = 0; mt < NMT; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++)
        if (pf_live<SHAPE>(MT0 + mt, nt)) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}

}  // namespace gpm
