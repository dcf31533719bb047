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

namespace gpm {

constexpr int NT8 = NB / 8;                         // 16 tiles per block edge
constexpr int PACKED = NT8 * (NT8 + 1) / 2 * 64;    // doubles in the packed lower triangle (136 tiles)
constexpr int POTF2_SMEM = (PACKED + 64 + NB + 4 * NB) * 8;  // 69.5 KB + 5.5 KB: three CTAs per SM (<= 75 KB each)

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
template <int S, int P2_WARPS>
__device__ __forceinline__ void inv_level_dmma(double* sm, int warp, int lane) {
  constexpr int TB = S / 8;                 // 8x8 tiles per block edge
  constexpr int TILES = (NB / (2 * S)) * TB * TB;
  constexpr int PER_WARP = (TILES + P2_WARPS - 1) / P2_WARPS;
  static_assert((TB * TB) % PER_WARP == 0, "a warp's tiles must belong to one pair of blocks");
  const int g = lane >> 2, q = lane & 3;
  const int sw = ((g >> 1) & 1) << 2;
  const int a_in = g * 8 + (q ^ sw);                              // A fragment: row g, col q
  const int a4 = ((q ^ sw) ^ 4) - (q ^ sw);                       // ... col q + 4
  const int b_in = q * 8 + (g ^ (((q >> 1) & 1) << 2));           // B fragment: row q, col g; row q + 4 is +32
  const int c_in = g * 8 + ((2 * q) ^ sw);                        // C fragment: row g, cols 2q, 2q+1
  double c0[PER_WARP], c1[PER_WARP];
  int ta[PER_WARP], tb[PER_WARP];
  const int tw = warp * PER_WARP;
  const bool active = tw < TILES;
  const int t1 = (active ? tw / (TB * TB) : 0) * 2 * TB;          // first tile row/column of this warp's pair
#pragma unroll
  for (int e = 0; e < PER_WARP; e++) {
    const int u = (tw + e) % (TB * TB);
    if (TB >= 2) {
      constexpr int H = TB >= 2 ? TB / 2 : 1;
      const int quad = u >> 2, m = u & 3, a = quad / H, b = quad % H;
      ta[e] = (m & 1) ? TB - 1 - a : a;
      tb[e] = (m == 1 || m == 2) ? TB - 1 - b : b;
    } else {
      ta[e] = 0; tb[e] = 0;
    }
  }
  // phase 1: T = L21 * X11   (X11 lower: contraction tiles kt >= tb)
  if (active) {
#pragma unroll
    for (int e = 0; e < PER_WARP; e++) {
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
  __syncthreads();
  if (active) {
#pragma unroll
    for (int e = 0; e < PER_WARP; e++)
      *reinterpret_cast<double2*>(sm + tile_base(t1 + TB + ta[e], t1 + tb[e]) + c_in) = make_double2(c0[e], c1[e]);
  }
  __syncthreads();
  // phase 2: X21 = -X22 * T   (X22 lower: contraction tiles kt <= ta)
  if (active) {
#pragma unroll
    for (int e = 0; e < PER_WARP; e++) {
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
  __syncthreads();
  if (active) {
#pragma unroll
    for (int e = 0; e < PER_WARP; e++)
      *reinterpret_cast<double2*>(sm + tile_base(t1 + TB + ta[e], t1 + tb[e]) + c_in) = make_double2(c0[e], c1[e]);
  }
  __syncthreads();
}

#ifdef GPM_POTF2_TIMING
// phase stamps of CTA 0 (cycles): the branch on a value loaded from shared memory after the barrier keeps
// the clock read behind the barrier's completion (BAR.SYNC.DEFER_BLOCKING lets independent work issue early)
__device__ long long g_p2_marks[64];
#define P2_MARK(slot)                                                                        \
  if (tid == 0 && blockIdx.x == 0) {                                                         \
    const double pv_ = *reinterpret_cast<volatile double*>(sm + PACKED + 63);                \
    if (__double_as_longlong(pv_) != 0x7ff8dead0000beefLL) g_p2_marks[slot] = clock64();     \
  }
#else
#define P2_MARK(slot)
#endif

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
                 const double* __restrict__ rhs_r, double* __restrict__ rhs_z, int R, long long batch_rhs_rows) {
  extern __shared__ __align__(16) double sm[];
  double* l8 = sm + PACKED;            // [36] current 8x8 factor (packed lower)
  double* rd = l8 + 64;                // [128] reciprocals of the diagonal of L
  constexpr int P2_WARPS = P2_THREADS / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  K += blockIdx.x * batch_k;
  invD += blockIdx.x * batch_inv + (long long)kblk * NB * NB;
  info += blockIdx.x;
  const long long r0 = (long long)kblk * NB;
  const int nv = (int)((N - r0) < NB ? (N - r0) : NB);

#ifdef GPM_POTF2_TIMING
  if (tid == 0 && blockIdx.x == 0) g_p2_marks[58] = clock64();
#endif
  // load the lower triangle as 16-byte pairs (LU independent loads in flight per thread); identity padding
  // beyond nv.  A pair never straddles a tile and keeps its order under the in-tile swizzle (bit 2 only).
  constexpr int LU = P2_THREADS == 256 ? 8 : 16;   // loads in flight per thread (register budget)
  for (int u0 = 0; u0 < NB * NB / 2 / P2_THREADS; u0 += LU) {
    double2 v[LU];
#pragma unroll
    for (int u = 0; u < LU; u++) {
      const int idx = tid + P2_THREADS * (u0 + u), i = idx >> 6, c = (idx & 63) * 2;
      if (i < nv && c <= i) v[u] = *reinterpret_cast<const double2*>(K + (r0 + i) * ldk + r0 + c);
      else v[u] = make_double2((i == c && i >= nv) ? 1.0 : 0.0, (i == c + 1 && i >= nv) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < LU; u++) {
      const int idx = tid + P2_THREADS * (u0 + u), i = idx >> 6, c = (idx & 63) * 2;
      if ((c >> 3) <= (i >> 3)) *reinterpret_cast<double2*>(sm + toff(i, c)) = v[u];
    }
  }
  P2_MARK(0)
  __syncthreads();
  P2_MARK(1)

  const int g = lane >> 2, q = lane & 3;
  const int c_in = g * 8 + ((2 * q) ^ (((g >> 1) & 1) << 2));
  const int x_in = g * 8 + (q ^ (((g >> 1) & 1) << 2));          // panel fragment: row g, column q
  const int x4 = (x_in ^ 4) - x_in;                               // ... and column q + 4
  constexpr int UW = P2_WARPS - 1;                                // update warps; warp UW looks ahead
  if (tid == 0) {
    const int bad = chol8_tile(sm, tile_base(0, 0), l8, rd);
    if (bad && bad - 1 < nv) atomicCAS(info, 0, (int)(r0 + bad));
  }
  __syncthreads();
  for (int p = 0; p < 16; p++) {
    const int c0 = 8 * p;
    P2_MARK(2 + 3 * p)
    // (1) panel solve by forward substitution against the factored 8x8 diagonal tile (l8, rd), one thread
    //     per row below it:  x_c = (a_c - sum_{k<c} x_k L8[c][k]) / L8[c][c]
    if (tid < NB && tid >= c0 + 8) {
      double x[8];
      double* row = sm + tile_base(tid >> 3, p) + (tid & 7) * 8;
      const int sw = ((tid >> 1) & 1) << 2;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        double v = row[c ^ sw];
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k < c) v = fma(-x[k], l8[c * (c + 1) / 2 + k], v);
        x[c] = v * rd[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < 8; c++) row[c ^ sw] = x[c];
    }
    __syncthreads();
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
        const int bad = chol8_tile(sm, tile_base(rbt, rbt), l8, rd + c0 + 8);
        if (bad && c0 + 8 + bad - 1 < nv) atomicCAS(info, 0, (int)(r0 + c0 + 8 + bad));
      }
    } else {
      // contiguous chunk of tiles per warp: consecutive tiles share their row, so the A fragments are
      // reloaded only at a row change and the C / B addresses advance by constant strides
      const int per = (T - 1 + UW - 1) / UW;
      int t = 1 + warp * per;
      const int tend = min(T, t + per);
      if (t < tend) {
        int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
        while (ti * (ti + 1) / 2 > t) ti--;
        int tj = t - ti * (ti + 1) / 2;
        const int bb0 = tile_base(rbt, p) + x_in;
        int cb = tile_base(rbt + ti, rbt + tj) + c_in;      // C fragment of tile (ti, tj)
        int bb = tile_base(rbt + tj, p) + x_in;             // B fragment: panel rows of tile-row tj
        double a0, a1;
        { const double* xa = sm + tile_base(rbt + ti, p) + x_in; a0 = -xa[0]; a1 = -xa[x4]; }
        auto step = [&]() {
          tj++; cb += 64; bb += (rbt + tj) * 64;
          if (tj > ti) {
            ti++; tj = 0;
            cb = tile_base(rbt + ti, rbt) + c_in; bb = bb0;
            const double* xa = sm + tile_base(rbt + ti, p) + x_in; a0 = -xa[0]; a1 = -xa[x4];
          }
        };
        while (t < tend) {
          const bool two = t + 1 < tend;
          const int cbA = cb, bbA = bb;
          const double aA0 = a0, aA1 = a1;
          if (two) step();
          const int cbB = cb, bbB = bb;
          const double aB0 = a0, aB1 = a1;
          if (t + 2 < tend) step();
          t += 2;
          double2 cA = *reinterpret_cast<const double2*>(sm + cbA), cB = *reinterpret_cast<const double2*>(sm + cbB);
          const double bA0 = sm[bbA], bA1 = sm[bbA + x4], bB0 = sm[bbB], bB1 = sm[bbB + x4];
          dmma(cA.x, cA.y, aA0, bA0);
          dmma(cB.x, cB.y, aB0, bB0);
          dmma(cA.x, cA.y, aA1, bA1);
          dmma(cB.x, cB.y, aB1, bB1);
          *reinterpret_cast<double2*>(sm + cbA) = cA;
          if (two) *reinterpret_cast<double2*>(sm + cbB) = cB;
        }
      }
    }
    __syncthreads();
    P2_MARK(4 + 3 * p)
  }

  // ---- write L_kk (lower part, valid rows), 16-byte pairs ----
  for (int u0 = 0; u0 < NB * NB / 2 / P2_THREADS; u0 += 16) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int idx = tid + P2_THREADS * (u0 + u), i = idx >> 6, c = (idx & 63) * 2;
      if (i < nv && c <= i) {
        const double2 v = *reinterpret_cast<const double2*>(sm + toff(i, c));
        double* dst = K + (r0 + i) * ldk + r0 + c;
        if (c < i) *reinterpret_cast<double2*>(dst) = v; else *dst = v.x;
      }
    }
  }
  P2_MARK(50)
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
  __syncthreads();
  if (tid < NB) {       // each column of each diagonal tile gets its inverse (zeros above the diagonal)
    const int pt = tid >> 3, j = tid & 7;
    double* Lb = sm + tile_base(pt, pt);
#pragma unroll
    for (int i = 0; i < 8; i++) Lb[in_tile(i, j)] = xcol[i];
  }
  __syncthreads();
  P2_MARK(51)
  inv_level_dmma<8, P2_WARPS>(sm, warp, lane);
  P2_MARK(52)
  inv_level_dmma<16, P2_WARPS>(sm, warp, lane);
  P2_MARK(53)
  inv_level_dmma<32, P2_WARPS>(sm, warp, lane);
  P2_MARK(54)
  inv_level_dmma<64, P2_WARPS>(sm, warp, lane);
  P2_MARK(55)
  // fused forward substitution (batched fits): z_k = inv(L_kk) r_k on DMMA tiles, four right-hand sides at a
  // time (staged in shared memory as the B operand, zero-padded to the 8 columns of a tile).  A warp owns the
  // row tiles {w, 15 - w} (8 warps: 17 contraction tiles each) or {w} (16 warps).
  if (rhs_r != nullptr) {
    double* rs = rd + NB;                // [128][4]
    const double* rk = rhs_r + (blockIdx.x * batch_rhs_rows + r0) * R;
    double* zk = rhs_z + (blockIdx.x * batch_rhs_rows + r0) * R;
    for (int rc = 0; rc < R; rc += 4) {
      const int nc = min(4, R - rc);
      for (int idx = tid; idx < NB * 4; idx += P2_THREADS) {
        const int c = idx >> 2, j = idx & 3;
        rs[idx] = (c < nv && j < nc) ? rk[c * R + rc + j] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int hh = 0; hh < 16 / P2_WARPS; hh++) {
        const int ti = hh == 0 ? warp : 15 - warp;
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
      __syncthreads();
    }
  }
  P2_MARK(57)
  for (int u0 = 0; u0 < NB * NB / 2 / P2_THREADS; u0 += 16) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int idx = tid + P2_THREADS * (u0 + u), i = idx >> 6, c = (idx & 63) * 2;
      reinterpret_cast<double2*>(invD)[idx] =
          ((c >> 3) <= (i >> 3)) ? *reinterpret_cast<const double2*>(sm + toff(i, c)) : make_double2(0.0, 0.0);
    }
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
    potf2_inv_kernel<256><<<batch, 256, POTF2_SMEM, stream>>>(K, ldk, N, kblk, invD, info, batch_k, batch_inv, rhs_r, rhs_z, R, batch_rhs_rows);
  else
    potf2_inv_kernel<512><<<batch, 512, POTF2_SMEM, stream>>>(K, ldk, N, kblk, invD, info, batch_k, batch_inv, rhs_r, rhs_z, R, batch_rhs_rows);
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
  rc = ensure_events(h, 2 * nblk + 2);
  if (rc) return rc;

  rc = launch_zero_info(info, batch, s0);
  if (rc) return rc;

  auto panel = [&](int k, cudaStream_t st) -> int {
    int r = launch_potf2(h, K, ldk, N, k, invD, info, batch, batch_k, batch_inv, st, rhs_r, rhs_z, R, batch_rhs_rows);
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
  // trailing update with block columns [k0, k0+kw) of L (contraction length kw*NB) restricted to tile
  // columns [jlo, jhi) (block indices), rows >= column
  auto update = [&](int k0, int kw, int jlo, int jhi, cudaStream_t st) -> int {
    if (jlo >= jhi || jlo >= nblk) return 0;
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
  // hidden behind the rest-update: width 8 while >= 96 block columns remain, 4 while >= 64, 2 while >= 32,
  // then 1 (thresholds swept on B200 at N = 8192 and 16384, tools/potrf_sweep.sh and potrf_sweep2.sh; the response
  // is flat within 1 % around these values).
  // Large batches are throughput-bound in every launch, so they use width 2 and no look-ahead.
  const int wide_env = h->opt.wide_min, wide4_env = h->opt.wide4_min, wide8_env = h->opt.wide8_min;
  std::vector<int> pb(nblk + 1), pw(nblk + 1);
  int npanel = 0;
  for (int b = 0; b < nblk;) {
    const int left = nblk - b;
    int w = 1;
    if (batch >= 32) w = left >= 2 ? 2 : 1;
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
  if ((rc = outer_panel(pb[0], pw[0], s0))) return rc;
  GPM_CUDA(cudaEventRecord(ev_panel[0], s0));
  for (int P = 0; P + 1 < npanel; P++) {
    const int b0 = pb[P], w = pw[P], n0 = pb[P + 1], wn = pw[P + 1];
    // helper stream: the block columns of the next outer panel first, then that panel
    GPM_CUDA(cudaStreamWaitEvent(s1, ev_panel[P], 0));
    if (P > 0) GPM_CUDA(cudaStreamWaitEvent(s1, ev_rest[P - 1], 0));
    for (int c = 0; c < wn; c++)
      if ((rc = update(b0, w, n0 + c, n0 + c + 1, s1))) return rc;
    if ((rc = outer_panel(n0, wn, s1))) return rc;
    GPM_CUDA(cudaEventRecord(ev_panel[P + 1], s1));
    // caller's stream: the rest of the step-P update (tile columns beyond the next panel)
    GPM_CUDA(cudaStreamWaitEvent(s0, ev_panel[P], 0));
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
  return (int)cudaMemcpyFromSymbol(out, g_p2_marks, sizeof(long long) * 64);
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
