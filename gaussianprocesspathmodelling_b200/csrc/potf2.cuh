// The 128 x 128 diagonal-block kernel of the factorisation as device functions, shared by potf2_inv_kernel (one CTA
// per block, potrf.cu) and the one-CTA-per-path fit (pathfit.cu), where the block arrives in shared memory straight
// from the tensor-core update and the threads are a sub-group of the CTA synchronised by a named barrier.
//
//   potf2_factor  -- Cholesky of the block held as its packed lower triangle of 8 x 8 tiles in shared memory
//   potf2_invert  -- in-place inverse of the factor (recursive doubling on DMMA tiles)
//   potf2_fwd_z   -- z_k = inv(L_kk) r_k on DMMA tiles (forward substitution fused into the factorisation)
//
// Template parameters: P2_THREADS = participating threads (128, 256 or 512; `tid` is the index inside the group);
// BAR = 0 synchronises with __syncthreads(), otherwise with the named barrier BAR over P2_THREADS threads.
#pragma once
#include "common.cuh"

#ifndef P2_MARK
#define P2_MARK(slot)
#endif
#ifndef P2_MARK_LA
#define P2_MARK_LA(p, e)
#endif

namespace gpm {

template <int NT, int BAR>
__device__ __forceinline__ void p2_sync() {
  if constexpr (BAR == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
}

constexpr int NT8 = NB / 8;                         // 16 tiles per block edge
constexpr int PACKED = NT8 * (NT8 + 1) / 2 * 64;    // doubles in the packed lower triangle (136 tiles)
constexpr int POTF2_DOUBLES = PACKED + 64 + NB + 4 * NB;      // packed triangle + 8x8 factor + 1/diag + right-hand-side stage
constexpr int POTF2_SMEM = POTF2_DOUBLES * 8;                 // 69.5 KB + 5.5 KB: three CTAs per SM (<= 75 KB each)

// The 128x128 diagonal block lives in shared memory as its lower triangle of 8x8 tiles (tile (ti,tj),
// tj <= ti, at index ti(ti+1)/2 + tj, 64 contiguous doubles, row-major).  Inside a tile the column is
// XOR-swizzled with bit 1 of the row, which keeps every access pattern of the kernel at the minimum
// number of shared-memory wavefronts: DMMA C fragments (128-bit, 32 lanes = 512 contiguous bytes),
// A fragments (row g, col q) and B fragments (row q, col g) two lanes per 8-byte bank pair.
__device__ __forceinline__ int tile_base(int ti, int tj) { return (ti * (ti + 1) / 2 + tj) * 64; }
__device__ __forceinline__ int in_tile(int i, int c) { return (i & 7) * 8 + ((c & 7) ^ (((i >> 1) & 1) << 2)); }
__device__ __forceinline__ int toff(int i, int c) { return tile_base(i >> 3, c >> 3) + in_tile(i, c); }

// 4x4 lower Cholesky in one thread's registers (right-looking, so the serial chain per column is
// rsqrt -> scale -> one FMA).  a: packed lower (a[i*(i+1)/2 + j]); on exit a holds L and r[j] = 1/L_jj.
// Every loop has constant bounds with compile-time-foldable guards so the arrays stay in registers.
// Returns the 1-based index of the first non-positive pivot (0 if none).
__device__ __forceinline__ int chol4(double (&a)[10], double (&r)[4]) {
  int bad = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double d = a[j * (j + 1) / 2 + j];
    if (!(d > 0.0) || !(d < 1.0e300)) { if (!bad) bad = j + 1; d = 1.0; }
    r[j] = rsqrt(d);                       // 1 ulp; sqrt + divide would cost ~5x the latency on this serial path
    a[j * (j + 1) / 2 + j] = d * r[j];
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (i > j) a[i * (i + 1) / 2 + j] *= r[j];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int c = 0; c < 4; c++)
        if (i > j && c > j && c <= i)
          a[i * (i + 1) / 2 + c] = fma(-a[i * (i + 1) / 2 + j], a[c * (c + 1) / 2 + j], a[i * (i + 1) / 2 + c]);
  }
  return bad;
}

// 8x8 lower Cholesky of the diagonal tile at sm[tb0..] by ONE thread, as a 2x2 blocking of 4x4 blocks:
// factor A11, solve L21, update and factor A22.  At most 26 matrix entries are live at a time (a flat 8x8
// keeps 36 + 8 doubles live, which no longer fits the 85-register budget of three CTAs per SM); every
// element still sees the same sequence of operations as the flat right-looking form.  Writes L into the
// tile and into l8 (packed), 1/L_jj into rd8.  Returns the 1-based index of the first bad pivot or 0.
__device__ __forceinline__ int chol8_tile(double* sm, int tb0, double* l8, double* rd8) {
  double a[10], r[4], x[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j <= i) a[i * (i + 1) / 2 + j] = sm[tb0 + in_tile(i, j)];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int c = 0; c < 4; c++) x[i][c] = sm[tb0 + in_tile(4 + i, c)];
  int bad = chol4(a, r);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    rd8[i] = r[i];
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j <= i) { sm[tb0 + in_tile(i, j)] = a[i * (i + 1) / 2 + j]; l8[i * (i + 1) / 2 + j] = a[i * (i + 1) / 2 + j]; }
  }
  // L21 = A21 * inv(L11)^T
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double v = x[i][c];
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k < c) v = fma(-x[i][k], a[c * (c + 1) / 2 + k], v);
      x[i][c] = v * r[c];
    }
  double b[10], r2[4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j <= i) b[i * (i + 1) / 2 + j] = sm[tb0 + in_tile(4 + i, 4 + j)];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      sm[tb0 + in_tile(4 + i, c)] = x[i][c];
      l8[(4 + i) * (5 + i) / 2 + c] = x[i][c];
    }
  // A22 -= L21 L21^T
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (j <= i) b[i * (i + 1) / 2 + j] = fma(-x[i][k], x[j][k], b[i * (i + 1) / 2 + j]);
  const int bad2 = chol4(b, r2);
  if (!bad && bad2) bad = 4 + bad2;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    rd8[4 + i] = r2[i];
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j <= i) {
        sm[tb0 + in_tile(4 + i, 4 + j)] = b[i * (i + 1) / 2 + j];
        l8[(4 + i) * (5 + i) / 2 + 4 + j] = b[i * (i + 1) / 2 + j];
      }
  }
  return bad;
}

// One level of the recursive-doubling inverse on DMMA tiles:  X21 = -X22 * (L21 * X11)  for all
// 64/S pairs of SxS diagonal blocks (X11, X22 already inverted in place, upper parts zero).
// Only tiles on or below the diagonal are touched (they are the only ones stored).  Each 8x8 result tile
// is a run of DMMA pairs over its non-zero contraction tiles; the operand addresses advance by constant
// (A) or linearly growing (B, packed rows) strides, so the inner loop is two shared loads, two DMMAs and
// two pointer updates.  Tile (ta, tb) costs 2(TB - tb) DMMAs in phase 1 and 2(ta + 1) in phase 2, so tiles
// are handed out in balanced pairs {(a, b), (TB-1-a, TB-1-b)} and {(a, TB-1-b), (TB-1-a, b)}.
template <int S, int P2_WARPS, int BAR>
__device__ __forceinline__ void inv_level_dmma(double* sm, int warp, int lane) {
  constexpr int TB = S / 8;                 // 8x8 tiles per block edge
  constexpr int TILES = (NB / (2 * S)) * TB * TB;
  constexpr int PER_WARP = (TILES + P2_WARPS - 1) / P2_WARPS;
  static_assert((TB * TB) % PER_WARP == 0 || PER_WARP % (TB * TB) == 0, "a warp's tiles must cover whole pairs of blocks or lie inside one");
  const int g = lane >> 2, q = lane & 3;
  const int sw = ((g >> 1) & 1) << 2;
  const int a_in = g * 8 + (q ^ sw);                              // A fragment: row g, col q
  const int a4 = ((q ^ sw) ^ 4) - (q ^ sw);                       // ... col q + 4
  const int b_in = q * 8 + (g ^ (((q >> 1) & 1) << 2));           // B fragment: row q, col g; row q + 4 is +32
  const int c_in = g * 8 + ((2 * q) ^ sw);                        // C fragment: row g, cols 2q, 2q+1
  double c0[PER_WARP], c1[PER_WARP];
  int ta[PER_WARP], tb[PER_WARP], t1s[PER_WARP];
  const int tw = warp * PER_WARP;
  const bool active = tw < TILES;
#pragma unroll
  for (int e = 0; e < PER_WARP; e++) {
    const int u = (tw + e) % (TB * TB);
    t1s[e] = (active ? (tw + e) / (TB * TB) : 0) * 2 * TB;        // first tile row/column of the pair of blocks tile e belongs to
    if (TB >= 2) {
      constexpr int H = TB >= 2 ? TB / 2 : 1;
      const int quad = u >> 2, m = u & 3, a = quad / H, b = quad % H;
      ta[e] = (m & 1) ? TB - 1 - a : a;
      tb[e] = (m == 1 || m == 2) ? TB - 1 - b : b;
    } else {
      ta[e] = 0; tb[e] = 0;
    }
  }
  // A warp that owns whole quads {(a, b), (TB-1-a, TB-1-b), (a, TB-1-b), (TB-1-a, b)} runs two tiles at a time: in
  // phase 1 the two of a quad that share their column (same B operand, same contraction range), in phase 2 the two
  // that share their row (same A operand) -- three operand loads per DMMA pair of a tile instead of four (the batched
  // kernel is bound by shared-memory wavefronts, and these loads were 28 % of them) and four independent DMMA chains.
  // Every tile still sees the same sequence of operations.
  constexpr bool PAIRED = TB >= 2 && PER_WARP % 4 == 0 && P2_WARPS >= 8;   // (the four-warp group of pathfit.cu measured 2 % slower paired)
  // phase 1: T = L21 * X11   (X11 lower: contraction tiles kt >= tb)
  if (PAIRED && active) {
#pragma unroll
    for (int e4 = 0; e4 < PER_WARP; e4 += 4)
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++) {
        const int ea = e4 + h2, eb = e4 + 3 - h2;             // (m = 0, 3) and (m = 1, 2): same tb
        const int t1 = t1s[ea];
        const double* pa = sm + tile_base(t1 + TB + ta[ea], t1 + tb[ea]) + a_in;
        const double* pa2 = sm + tile_base(t1 + TB + ta[eb], t1 + tb[ea]) + a_in;
        const double* pb = sm + tile_base(t1 + tb[ea], t1 + tb[ea]) + b_in;
        double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0, u0 = 0.0, u1 = 0.0, v0 = 0.0, v1 = 0.0;
        for (int kt = tb[ea]; kt < TB; kt++) {
          const double b0 = pb[0], b1 = pb[32];
          dmma(x0, x1, pa[0], b0);
          dmma(u0, u1, pa2[0], b0);
          dmma(y0, y1, pa[a4], b1);
          dmma(v0, v1, pa2[a4], b1);
          pa += 64; pa2 += 64;
          pb += (t1 + kt + 1) * 64;
        }
        c0[ea] = x0 + y0; c1[ea] = x1 + y1;
        c0[eb] = u0 + v0; c1[eb] = u1 + v1;
      }
  } else if (active) {
#pragma unroll
    for (int e = 0; e < PER_WARP; e++) {
      const int t1 = t1s[e];
      const double* pa = sm + tile_base(t1 + TB + ta[e], t1 + tb[e]) + a_in;
      const double* pb = sm + tile_base(t1 + tb[e], t1 + tb[e]) + b_in;
      double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
      for (int kt = tb[e]; kt < TB; kt++) {
        dmma(x0, x1, pa[0], pb[0]);
        dmma(y0, y1, pa[a4], pb[32]);
        pa += 64;
        pb += (t1 + kt + 1) * 64;
      }
      c0[e] = x0 + y0; c1[e] = x1 + y1;
    }
  }
  p2_sync<P2_WARPS * 32, BAR>();
  if (active) {
#pragma unroll
    for (int e = 0; e < PER_WARP; e++)
      *reinterpret_cast<double2*>(sm + tile_base(t1s[e] + TB + ta[e], t1s[e] + tb[e]) + c_in) = make_double2(c0[e], c1[e]);
  }
  p2_sync<P2_WARPS * 32, BAR>();
  // phase 2: X21 = -X22 * T   (X22 lower: contraction tiles kt <= ta)
  if (PAIRED && active) {
#pragma unroll
    for (int e4 = 0; e4 < PER_WARP; e4 += 4)
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++) {
        const int ea = e4 + h2, eb = e4 + 2 + h2;             // (m = 0, 2) and (m = 1, 3): same ta
        const int t1 = t1s[ea];
        const double* pa = sm + tile_base(t1 + TB + ta[ea], t1 + TB) + a_in;
        const double* pb = sm + tile_base(t1 + TB, t1 + tb[ea]) + b_in;
        const double* pb2 = sm + tile_base(t1 + TB, t1 + tb[eb]) + b_in;
        double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0, u0 = 0.0, u1 = 0.0, v0 = 0.0, v1 = 0.0;
        for (int kt = 0; kt <= ta[ea]; kt++) {
          const double a0 = pa[0], a1 = pa[a4];
          dmma(x0, x1, a0, pb[0]);
          dmma(u0, u1, a0, pb2[0]);
          dmma(y0, y1, a1, pb[32]);
          dmma(v0, v1, a1, pb2[32]);
          pa += 64;
          pb += (t1 + TB + kt + 1) * 64; pb2 += (t1 + TB + kt + 1) * 64;
        }
        c0[ea] = -(x0 + y0); c1[ea] = -(x1 + y1);
        c0[eb] = -(u0 + v0); c1[eb] = -(u1 + v1);
      }
  } else if (active) {
#pragma unroll
    for (int e = 0; e < PER_WARP; e++) {
      const int t1 = t1s[e];
      const double* pa = sm + tile_base(t1 + TB + ta[e], t1 + TB) + a_in;
      const double* pb = sm + tile_base(t1 + TB, t1 + tb[e]) + b_in;
      double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
      for (int kt = 0; kt <= ta[e]; kt++) {
        dmma(x0, x1, pa[0], pb[0]);
        dmma(y0, y1, pa[a4], pb[32]);
        pa += 64;
        pb += (t1 + TB + kt + 1) * 64;
      }
      c0[e] = -(x0 + y0); c1[e] = -(x1 + y1);
    }
  }
  p2_sync<P2_WARPS * 32, BAR>();
  if (active) {
#pragma unroll
    for (int e = 0; e < PER_WARP; e++)
      *reinterpret_cast<double2*>(sm + tile_base(t1s[e] + TB + ta[e], t1s[e] + tb[e]) + c_in) = make_double2(c0[e], c1[e]);
  }
  p2_sync<P2_WARPS * 32, BAR>();
}


// Shared-memory layout of a block: sm[0, PACKED) packed lower triangle; l8 = sm + PACKED [64]; rd = l8 + 64 [128]
// (reciprocals of the diagonal of L); rs = rd + 128 [4 * 128] (right-hand sides of potf2_fwd_z).
//
// Right-looking over sixteen 8-column panels: (1) one thread factors the 8x8 diagonal block in registers, (2) one
// thread per row forward-substitutes the panel against it, (3) all warps apply the rank-8 update to the trailing
// 8x8 tiles with DMMA.8x8x4 on a static balanced schedule, the last warp looking ahead to the next diagonal tile.
// Rows / columns >= nv must hold the identity.  A non-positive pivot j (0-based, < nv) sets *info = r0 + j + 1 once.
// Ends with a barrier: the whole factor is visible to every thread of the group.
template <int P2_THREADS, int BAR>
__device__ __forceinline__ void potf2_factor(double* sm, int tid, int nv, long long r0, int* info) {
  double* l8 = sm + PACKED;
  double* rd = l8 + 64;
  constexpr int P2_WARPS = P2_THREADS / 32;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int c_in = g * 8 + ((2 * q) ^ (((g >> 1) & 1) << 2));
  const int x_in = g * 8 + (q ^ (((g >> 1) & 1) << 2));          // panel fragment: row g, column q
  const int x4 = (x_in ^ 4) - x_in;                               // ... and column q + 4
  constexpr int UW = P2_WARPS - 1;                                // warp UW looks ahead (factors the next diagonal tile)
  // The look-ahead factorisation is one thread's serial chain of ~100 dependent FP64 operations, and DMMA shares the
  // FP64 pipe of its scheduler (16 cycles per DMMA): with eight or more warps, the warps that sit on the look-ahead
  // warp's scheduler (warp index = 3 mod 4) stay out of the rank-8 update, so that the chain never queues behind them.
  constexpr int NU = P2_WARPS >= 8 ? P2_WARPS - P2_WARPS / 4 : UW;  // update warps
  const bool upd_warp = P2_WARPS >= 8 ? ((warp & 3) != 3) : (warp != UW);
  const int uidx = P2_WARPS >= 8 ? warp - (warp >> 2) : warp;
  if (tid == 0) {
    const int bad = chol8_tile(sm, tile_base(0, 0), l8, rd);
    if (bad && bad - 1 < nv) atomicCAS(info, 0, (int)(r0 + bad));
  }
  p2_sync<P2_THREADS, BAR>();
  for (int p = 0; p < 16; p++) {
    const int c0 = 8 * p;
    P2_MARK(2 + 3 * p)
    // (1) panel solve by forward substitution against the factored 8x8 diagonal tile (l8, rd), one thread
    //     per row below it:  x_c = (a_c - sum_{k<c} x_k L8[c][k]) / L8[c][c]
    if (tid < NB && tid >= c0 + 8) {
      // the row as four 16-byte pairs (a pair keeps its order under the in-tile swizzle): the eight rows of a tile are
      // 64 bytes apart, so 8-byte accesses run at eight wavefronts per instruction and 16-byte ones at two per quarter
      // warp -- half the shared-memory traffic of this step (the batched kernel is bound by shared-memory wavefronts)
      double x[8];
      double* row = sm + tile_base(tid >> 3, p) + (tid & 7) * 8;
      const int sw = ((tid >> 1) & 1) << 2;
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        const double2 v = *reinterpret_cast<const double2*>(row + (c ^ sw));
        x[c] = v.x; x[c + 1] = v.y;
      }
#pragma unroll
      for (int c = 0; c < 8; c++) {
        double v = x[c];
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k < c) v = fma(-x[k], l8[c * (c + 1) / 2 + k], v);
        x[c] = v * rd[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < 8; c += 2) *reinterpret_cast<double2*>(row + (c ^ sw)) = make_double2(x[c], x[c + 1]);
    }
    p2_sync<P2_THREADS, BAR>();
    P2_MARK(3 + 3 * p)
    if (p == 15) break;
    // (2) rank-8 update of the trailing 8x8 tiles, C[ti][tj] -= X_ti X_tj^T (DMMA), with look-ahead: the last
    //     warp updates the next diagonal tile first and factors it (one lane) while the other warps update
    //     the remaining tiles, so the serial 8x8 factorisation leaves the critical path of the big panels.
    const int nt = 15 - p, rbt = p + 1;
    const int T = nt * (nt + 1) / 2;             // trailing tiles in row-major lower order; tile 0 = (rbt, rbt)
    if (warp == UW) {
      const double* xa = sm + tile_base(rbt, p) + x_in;
      double2* cp = reinterpret_cast<double2*>(sm + tile_base(rbt, rbt) + c_in);
      const double b0 = xa[0], b1 = xa[x4];
      double2 c = *cp;
      dmma(c.x, c.y, -b0, b0);
      dmma(c.x, c.y, -b1, b1);
      *cp = c;
      __syncwarp();
      if (lane == 0) {
        P2_MARK_LA(p, 0)
        const int bad = chol8_tile(sm, tile_base(rbt, rbt), l8, rd + c0 + 8);
        if (bad && c0 + 8 + bad - 1 < nv) atomicCAS(info, 0, (int)(r0 + c0 + 8 + bad));
        P2_MARK_LA(p, 1)
      }
    } else if (upd_warp) {
      // contiguous chunk of tiles per warp, walked row by row: inside a row the A fragments stay in registers and the
      // C / B addresses advance by constant / linearly growing strides (packed rows), two tiles (two independent
      // DMMA chains) per iteration and no per-tile branching -- the first version of this loop spent 41 instructions
      // per tile on its bookkeeping and ran at a third of the DMMA rate of its scheduler (ncu instruction counts,
      // clock stamps: 600 cycles per pair of tiles)
      const int per = (T - 1 + NU - 1) / NU;
      int t = 1 + uidx * per;
      int left = min(T, t + per) - t;
      if (left > 0) {
        int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
        while (ti * (ti + 1) / 2 > t) ti--;
        int tj = t - ti * (ti + 1) / 2;
        int arow = tile_base(rbt + ti, p) + x_in;           // panel fragment of tile-row ti (the A operand)
        while (left > 0) {
          const int n = min(left, ti + 1 - tj);             // tiles of this row inside the chunk
          const double a0 = -sm[arow], a1 = -sm[arow + x4];
          int cb = tile_base(rbt + ti, rbt + tj) + c_in;    // C fragment of tile (ti, tj); +64 per tile
          int bb = tile_base(rbt + tj, p) + x_in;           // B fragment: panel rows of tile-row tj; next row: + bstep
          int bstep = (rbt + tj + 1) * 64;
          int j = 0;
          for (; j + 1 < n; j += 2) {
            double2 cA = *reinterpret_cast<const double2*>(sm + cb), cB = *reinterpret_cast<const double2*>(sm + cb + 64);
            const double bA0 = sm[bb], bA1 = sm[bb + x4], bB0 = sm[bb + bstep], bB1 = sm[bb + bstep + x4];
            dmma(cA.x, cA.y, a0, bA0);
            dmma(cB.x, cB.y, a0, bB0);
            dmma(cA.x, cA.y, a1, bA1);
            dmma(cB.x, cB.y, a1, bB1);
            *reinterpret_cast<double2*>(sm + cb) = cA;
            *reinterpret_cast<double2*>(sm + cb + 64) = cB;
            cb += 128; bb += 2 * bstep + 64; bstep += 128;
          }
          if (j < n) {
            double2 cA = *reinterpret_cast<const double2*>(sm + cb);
            const double bA0 = sm[bb], bA1 = sm[bb + x4];
            dmma(cA.x, cA.y, a0, bA0);
            dmma(cA.x, cA.y, a1, bA1);
            *reinterpret_cast<double2*>(sm + cb) = cA;
          }
          left -= n;
          ti++; tj = 0;
          arow += (rbt + ti) * 64;                           // tile_base(r + 1, p) - tile_base(r, p) = (r + 1) 64
        }
      }
    }
    p2_sync<P2_THREADS, BAR>();
    P2_MARK(4 + 3 * p)
  }

}

// In-place inverse of the factor in sm (the diagonal reciprocals come from rd).  Ends with a barrier.
template <int P2_THREADS, int BAR>
__device__ __forceinline__ void potf2_invert(double* sm, int tid) {
  double* rd = sm + PACKED + 64;
  constexpr int P2_WARPS = P2_THREADS / 32;
  const int warp = tid >> 5, lane = tid & 31;
  // ---- inverse, level 0: the sixteen 8x8 diagonal blocks, one thread per column ----
  double xcol[8];
  {
    const int pt = tid >> 3, j = tid & 7;
    if (tid < NB) {
      const double* Lb = sm + tile_base(pt, pt);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k < i) v = fma(Lb[in_tile(i, k)], (k >= j) ? xcol[k] : 0.0, v);
        xcol[i] = (i < j) ? 0.0 : ((i == j) ? rd[pt * 8 + i] : -v * rd[pt * 8 + i]);
      }
    }
  }
  p2_sync<P2_THREADS, BAR>();
  if (tid < NB) {       // each column of each diagonal tile gets its inverse (zeros above the diagonal)
    const int pt = tid >> 3, j = tid & 7;
    double* Lb = sm + tile_base(pt, pt);
#pragma unroll
    for (int i = 0; i < 8; i++) Lb[in_tile(i, j)] = xcol[i];
  }
  p2_sync<P2_THREADS, BAR>();
  P2_MARK(51)
  inv_level_dmma<8, P2_WARPS, BAR>(sm, warp, lane);
  P2_MARK(52)
  inv_level_dmma<16, P2_WARPS, BAR>(sm, warp, lane);
  P2_MARK(53)
  inv_level_dmma<32, P2_WARPS, BAR>(sm, warp, lane);
  P2_MARK(54)
  inv_level_dmma<64, P2_WARPS, BAR>(sm, warp, lane);
  P2_MARK(55)
}

// z_k = inv(L_kk) r_k with the inverse in sm: rk, zk point at row 0 of the block's right-hand sides (nv valid rows,
// R columns, row-major).  Four right-hand sides at a time, staged in shared memory as the B operand.
// zk may alias rk (the single-matrix fit solves in place): every entry of a four-column chunk is staged in shared
// memory before any entry of that chunk is written.
// The first chunk of right-hand sides into the stage (no barrier): callers whose r_k is final before the
// factorisation starts call this early, so that the global-load latency hides behind the factorisation.
template <int P2_THREADS>
__device__ __forceinline__ void potf2_stage_rhs(double* sm, int tid, int nv, const double* rk, int R) {
  double* rs = sm + PACKED + 64 + NB;           // = rd + NB (potf2_fwd_z)
  const int nc = min(4, R);
  for (int idx = tid; idx < NB * 4; idx += P2_THREADS) {
    const int c = idx >> 2, j = idx & 3;
    rs[idx] = (c < nv && j < nc) ? rk[c * R + j] : 0.0;
  }
}

template <int P2_THREADS, int BAR, bool STAGED = false>
__device__ __forceinline__ void potf2_fwd_z(double* sm, int tid, int nv, const double* rk, double* zk, int R) {
  double* rd = sm + PACKED + 64;
  constexpr int P2_WARPS = P2_THREADS / 32;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int x_in = g * 8 + (q ^ (((g >> 1) & 1) << 2));
  const int x4 = (x_in ^ 4) - x_in;
  // fused forward substitution (batched fits): z_k = inv(L_kk) r_k on DMMA tiles, four right-hand sides at a
  // time (staged in shared memory as the B operand, zero-padded to the 8 columns of a tile).  A warp owns the
  // row tiles {w, 15 - w} (8 warps: 17 contraction tiles each) or {w} (16 warps).
  {
    double* rs = rd + NB;                // [128][4]
    for (int rc = 0; rc < R; rc += 4) {
      const int nc = min(4, R - rc);
      if (!(STAGED && rc == 0)) {
        for (int idx = tid; idx < NB * 4; idx += P2_THREADS) {
          const int c = idx >> 2, j = idx & 3;
          rs[idx] = (c < nv && j < nc) ? rk[c * R + rc + j] : 0.0;
        }
      }
      p2_sync<P2_THREADS, BAR>();
#pragma unroll
      for (int hh = 0; hh < 16 / P2_WARPS; hh++) {
        // balanced row tiles: {w, 15 - w} for 8 warps, {w, 7 - w, 8 + w, 15 - w} for 4, {w} for 16
        const int ti = P2_WARPS == 4 ? (hh >> 1) * 8 + ((hh & 1) ? 7 - warp : warp) : (hh == 0 ? warp : 15 - warp);
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
        for (int kt = 0; kt <= ti; kt++) {
          const double* ap = sm + tile_base(ti, kt) + x_in;
          const double b0 = g < 4 ? rs[(kt * 8 + q) * 4 + g] : 0.0;
          const double b1 = g < 4 ? rs[(kt * 8 + q + 4) * 4 + g] : 0.0;
          dmma(c0, c1, ap[0], b0);
          dmma(d0, d1, ap[x4], b1);
        }
        c0 += d0; c1 += d1;
        const int row = ti * 8 + g;
        if (row < nv) {
          if (2 * q < nc) zk[row * R + rc + 2 * q] = c0;
          if (2 * q + 1 < nc) zk[row * R + rc + 2 * q + 1] = c1;
        }
      }
      p2_sync<P2_THREADS, BAR>();
    }
  }
}

}  // namespace gpm
