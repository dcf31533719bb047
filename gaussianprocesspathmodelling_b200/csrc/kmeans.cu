// SURVEY.md section 8f "next" #1: the reference's real hot loop, trajectories.kmeansclustering (GPmap.py:36-93), as a
// kernel pair that keeps the whole Lloyd iteration on the device:
//
//   assign  (GPmap.py:72-80, 114-121)  dist[p,c] = sum_i sqrt(dx^2 + dy^2), summed sequentially in sample order as the
//                                      reference's Python float accumulation does; first minimum with strict '<'.
//   update  (GPmap.py:83-84, 95-112)   centroid c = point-wise mean of its members' xs, ys, timestamp, the members
//                                      summed in path order (the order the reference appends them to the cluster
//                                      list), then divided by the member count: bit-exact with calc_mean_traj.
//   finish  (GPmap.py:87-90)           shift = sum_c calc_distance(new_c, old_c);  converged when shift < threshold.
//
// The paths are uploaded once per clustering call in two layouts: sample-major [n][P] for the assignment (one
// thread per path: a warp's loads are 256 contiguous bytes) and path-major [P][n] for the update (one thread per
// sample: same).  Every kernel of an iteration returns at once when the device-side `converged` flag is set, so
// the host can enqueue several iterations back to back and look at the 16-byte state only once per batch.
#include <algorithm>

#include "common.cuh"

namespace gpm {

struct KmState {          // device-resident state of a clustering call
  int iters;              // finished Lloyd iterations
  int converged;          // set by the finish kernel when shift < threshold
  double shift;           // summed centroid shift of the last finished iteration
};

// pxT, pyT: [n][P].  cx, cy: [k][n].  One thread per path; centroids staged in shared memory.
__global__ void __launch_bounds__(128)
kmeans_assign_kernel(const double* __restrict__ pxT, const double* __restrict__ pyT, long long P, int n,
                     const double* __restrict__ cx, const double* __restrict__ cy, int k,
                     double* __restrict__ dist, int* __restrict__ assign, const KmState* __restrict__ state) {
  if (state && state->converged) return;
  extern __shared__ double sc[];          // centroids: cx[k][n] then cy[k][n]
  for (int e = threadIdx.x; e < k * n; e += blockDim.x) { sc[e] = cx[e]; sc[k * n + e] = cy[e]; }
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double best = 0.0;
  int best_c = 0;
  // eight centroids at a time: a sample of the path is loaded once for the eight (the path's 2n values were re-read
  // from L2 for every centroid: 422 MB per assignment at P = 100000, k = 8), the eight sqrt chains interleave; every
  // distance is still summed sequentially in sample order, and the minimum scans the centroids in order
  constexpr int KC = 8;
  for (int c0 = 0; c0 < k; c0 += KC) {
    double s[KC];
#pragma unroll
    for (int u = 0; u < KC; u++) s[u] = 0.0;
    const double* ccx = sc + c0 * n;
    const double* ccy = sc + k * n + c0 * n;
    for (int i = 0; i < n; i++) {
      const double x = pxT[(long long)i * P + p], y = pyT[(long long)i * P + p];
#pragma unroll
      for (int u = 0; u < KC; u++)
        if (c0 + u < k) {
          const double dx = x - ccx[u * n + i];
          const double dy = y - ccy[u * n + i];
          s[u] = __dadd_rn(s[u], sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
        }
    }
#pragma unroll
    for (int u = 0; u < KC; u++)
      if (c0 + u < k) {
        if (dist) dist[p * k + c0 + u] = s[u];
        if (c0 + u == 0 || s[u] < best) { best = s[u]; best_c = c0 + u; }
      }
  }
  assign[p] = best_c;
}

// ------------------------------------------------------------------------------------------------------------------
// Centroid update (calc_mean_traj, GPmap.py:95-112).  The additions of a centroid's members are a serial chain by
// construction -- path order, the order the reference appends them to the cluster list: the mean is bit-exact that way
// -- but finding the members and fetching their rows need not be serial, nor run on k SMs only:
//   rank    (all SMs)   every path gets its rank among the paths of its 2048-path chunk that share its centroid
//                       (__match_any_sync inside a warp, per-warp tables across the warps of a round); per-chunk counts
//   scan    (one CTA)   exclusive scan of the counts over the chunks, per centroid; row offset of every centroid's region
//   gather  (all SMs)   row p of (xs, ys, timestamp) -> row offset[c] + chunk_offset + rank of ONE ordered buffer
//   sum     (~148 CTAs) the 3n running sums of a centroid are independent chains: the row entries are cut into slices of W
//                       (the ordered buffer is slice-major, [slice][row][W]) and every (centroid, slice) CTA streams its
//                       own contiguous region through shared memory (TMA bulk copies, two stages) while its W summing
//                       threads add the rows in order
// History (P = 100000, n = 33, k = 8, per Lloyd iteration): each member's value loaded right before its addition
// 5.1 ms (ncu: 13 GB/s, issue slots 2 %: one DRAM round trip per member); in-kernel ballot compaction + staged gather on
// k SMs 1.1 ms; one CTA per centroid 0.56 ms = assignment 0.083 + rank 0.011 + scan 0.004 + gather 0.03 + sum 0.43.  That sum
// streamed 10 MB per centroid from DRAM through ONE SM at about 30 GB/s (stage sizes of 64 and 128 rows give 2.7 and
// 4.4 us per stage: 1 us of latency + bytes / 30 GB/s).  Measured on top and not kept, all at 0.55 - 0.65 ms: four
// stages of 64 rows (three bulk copies in flight), bulk copies in 4 KB pieces (slower) or one per stage, sixteen loads
// ahead of their additions in the inner loop, evict-first loads in the gather so that the ordered buffer stays in L2,
// and a chain of 16 CTAs per centroid passing the running sums on (each CTA can prefetch only its two stages, so the
// later stages of its segment still load on the chain).
// Last session of round 2: the stream is spread over the SMs by COLUMNS instead -- the sums of different row entries never
// meet, so (centroid, slice) CTAs need no hand-off at all: sum 0.43 -> 0.24 ms, iteration 0.56 -> 0.40 ms (1 M paths: 5.1 ->
// 3.5 ms).  What is left of the sum is the serial chain of the LARGEST cluster (37549 of the 100000 paths in the bench:
// 12.5 cycles per addition); the shift of the centroids moved to the finish kernel (distances in parallel, their sum in order).
// ------------------------------------------------------------------------------------------------------------------
constexpr int KM_CH = 2048;              // paths per chunk (one CTA of the rank / gather kernels)
constexpr int KM_THREADS = 256, KM_WARPS = KM_THREADS / 32;
constexpr int KM_ROUNDS = KM_CH / KM_THREADS;

__device__ __forceinline__ void km_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// lrank[p] = number of paths q < p of p's chunk with assign[q] == assign[p];  counts[chunk][c] = members of c in the chunk
__global__ void __launch_bounds__(KM_THREADS)
kmeans_rank_kernel(const int* __restrict__ assign, long long P, int k, int* __restrict__ lrank,
                   int* __restrict__ counts, const KmState* __restrict__ state) {
  if (state->converged) return;
  extern __shared__ int km_tab[];          // cnt[k] running counts of the chunk, then wtab[KM_WARPS][k] of the round
  int* cnt = km_tab;
  int* wtab = km_tab + k;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long p0 = (long long)blockIdx.x * KM_CH;
  for (int e = tid; e < k; e += KM_THREADS) cnt[e] = 0;
  for (int r = 0; r < KM_ROUNDS; r++) {
    const long long p = p0 + r * KM_THREADS + tid;
    const int a = p < P ? assign[p] : -1;
    for (int e = tid; e < KM_WARPS * k; e += KM_THREADS) wtab[e] = 0;
    __syncthreads();
    const unsigned same = __match_any_sync(0xffffffffu, a);
    const int before = __popc(same & ((1u << lane) - 1u));
    if (a >= 0 && before == 0) wtab[warp * k + a] = __popc(same);      // the group's first lane
    __syncthreads();
    if (a >= 0) {
      int base = cnt[a];
      for (int w = 0; w < warp; w++) base += wtab[w * k + a];
      lrank[p] = base + before;
    }
    __syncthreads();
    for (int e = tid; e < k; e += KM_THREADS) {
      int t = cnt[e];
#pragma unroll
      for (int w = 0; w < KM_WARPS; w++) t += wtab[w * k + e];
      cnt[e] = t;
    }
    __syncthreads();
  }
  for (int e = tid; e < k; e += KM_THREADS) counts[(long long)blockIdx.x * k + e] = cnt[e];
}

// choff[chunk][c] = members of c in earlier chunks; total[c]; coff[c] = first row of c's region (even: 16-byte aligned rows)
__global__ void __launch_bounds__(1024)
kmeans_scan_kernel(const int* __restrict__ counts, int nchunks, int k, int* __restrict__ choff,
                   long long* __restrict__ total, long long* __restrict__ coff, const KmState* __restrict__ state) {
  if (state->converged) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < k; c += 32) {
    long long run = 0;
    for (int t0 = 0; t0 < nchunks; t0 += 32) {
      const int ch = t0 + lane;
      const int v = ch < nchunks ? counts[(long long)ch * k + c] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
      if (ch < nchunks) choff[(long long)ch * k + c] = (int)(run + inc - v);
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) total[c] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long r = 0;
    for (int c = 0; c < k; c++) { coff[c] = r; r += (total[c] + 1) & ~1ll; }
  }
}

// row p of xs / ys / timestamp -> row coff[c] + choff[chunk][c] + lrank[p] of the ordered buffer G ([rows][3n])
__global__ void __launch_bounds__(KM_THREADS)
kmeans_gather_kernel(const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pt,
                     long long P, int n, int k, const int* __restrict__ assign, const int* __restrict__ lrank,
                     const int* __restrict__ choff, const long long* __restrict__ coff, double* __restrict__ G,
                     const KmState* __restrict__ state, int W, long long rows_total) {
  if (state->converged) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * KM_WARPS + warp;       // a warp per path: every copy of the launch is independent
  if (p >= P) return;
  const int c = assign[p];
  const long long row = coff[c] + choff[(p / KM_CH) * k + c] + lrank[p];
  const long long src = p * n;
  // entry w of the row (xs | ys | timestamp) -> column w % W of slice w / W
  auto put = [&](int w, double v) {
    const int sl = w / W;
    G[((long long)sl * rows_total + row) * W + (w - sl * W)] = v;
  };
  for (int i = lane; i < n; i += 32) {
    put(i, px[src + i]);
    put(n + i, py[src + i]);
    put(2 * n + i, pt[src + i]);
  }
}

// One CTA per (centroid, slice of W row entries): the 3n running sums of a centroid are independent chains, so the
// entries (xs | ys | timestamp) are cut into slices and every slice streams its own [members][W] region on its own SM
// (the ordered buffer is slice-major).  Thread w < W owns entry slice * W + w.  cold / cnew: [3][k][n].
// An empty cluster keeps its previous centroid.
__global__ void __launch_bounds__(KM_THREADS, 1)
kmeans_update_kernel(const double* __restrict__ G, int n, int k, const long long* __restrict__ total,
                     const long long* __restrict__ coff, const double* __restrict__ cold, double* __restrict__ cnew,
                     const KmState* __restrict__ state, int stage_rows, int W, long long rows_total) {
  if (state->converged) return;
  extern __shared__ __align__(16) double km_stage[];      // [2][stage_rows][W]
  const int c = blockIdx.x, sl = blockIdx.y, tid = threadIdx.x;
  const int work = W;                                      // row length of the slice (even: rows are 16-byte multiples)
  const int w0 = sl * W;                                   // first entry of the slice
  const int live = 3 * n - w0 < W ? 3 * n - w0 : W;        // entries of the slice that exist (the last slice may be padded)
  const long long m = total[c];
  const double* region = G + ((long long)sl * rows_total + coff[c]) * work;
  // each thread carries up to SLOTS running sums (entries tid, tid + KM_THREADS, ...)
  constexpr int SLOTS = 2048 / KM_THREADS;                 // 3 n <= 2048
  double sum[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; s++) sum[s] = 0.0;
  // rows [g0, g0 + cnt) of the region into stage `buf`: contiguous, so ONE thread moves them with bulk asynchronous
  // copies (TMA, completion counted in bytes on the stage's mbarrier).  Plain or cp.async loads top out near 20 GB/s
  // for a single SM (ncu on the one-CTA-per-centroid version: 164 GB/s over the 8 busy SMs).
  __shared__ __align__(8) unsigned long long km_bar[2];
  const uint32_t bar0 = smem_u32(&km_bar[0]);
  if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_mbar_init(); }
  __syncthreads();
  auto gather = [&](int buf, long long g0, int cnt) {       // thread 0 only
    const uint32_t dst0 = smem_u32(km_stage + (size_t)buf * stage_rows * work), bar = bar0 + 8u * buf;
    const char* src = reinterpret_cast<const char*>(region + g0 * work);
    const uint32_t bytes = ((uint32_t)(cnt * work) * 8u + 15u) & ~15u;
    mbar_arrive_expect_tx(bar, bytes);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
      const uint32_t sz = bytes - o < 32768u ? bytes - o : 32768u;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst0 + o), "l"(src + o), "r"(sz), "r"(bar) : "memory");
    }
  };
  uint32_t par = 0;                                          // bit b: parity of stage b's next completed phase
  const long long ngroups = (m + stage_rows - 1) / stage_rows;
  if (ngroups > 0 && tid == 0) gather(0, 0, (int)(m < stage_rows ? m : stage_rows));
  for (long long g = 0; g < ngroups; g++) {
    if (g + 1 < ngroups && tid == 0) {                       // stage (g + 1) & 1 was released by the barrier that ended group g - 1
      const long long left = m - (g + 1) * stage_rows;
      gather((int)((g + 1) & 1), (g + 1) * stage_rows, (int)(left < stage_rows ? left : stage_rows));
    }
    mbar_wait(bar0 + 8u * (uint32_t)(g & 1), (par >> (g & 1)) & 1u);
    par ^= 1u << (g & 1);
    const long long left = m - g * stage_rows;
    const int cnt = (int)(left < stage_rows ? left : stage_rows);
    const double* sb = km_stage + (size_t)(g & 1) * stage_rows * work;
#pragma unroll 1
    for (int s = 0; s < SLOTS; s++) {
      const int w = tid + KM_THREADS * s;
      if (w >= live) break;
      double acc = 0.0;
#pragma unroll
      for (int t = 0; t < SLOTS; t++) if (t == s) acc = sum[t];
      const double* col = sb + w;
      // the additions are one dependent chain (member order: bit-exact with the reference; 8.3 cycles per DADD), so a value
      // is reloaded for the row eight further on right after its addition: loads, address arithmetic and loop control
      // issue in the shadow of the chain instead of between two groups of additions (12.4 -> ~9 cycles per row)
      int r = 0;
      if (cnt >= 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = col[(size_t)u * work];
        const double* nx = col + (size_t)8 * work;
        for (; r + 16 <= cnt; r += 8) {
#pragma unroll
          for (int u = 0; u < 8; u++) {
            acc = __dadd_rn(acc, v[u]);
            v[u] = *nx;
            nx += work;
          }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) acc = __dadd_rn(acc, v[u]);
        r += 8;
      }
      for (; r < cnt; r++) acc = __dadd_rn(acc, col[(size_t)r * work]);
#pragma unroll
      for (int t = 0; t < SLOTS; t++) if (t == s) sum[t] = acc;
    }
    __syncthreads();                       // stage g & 1 may be refilled (by the gather of group g + 2)
  }
  const long long cnt = m;
#pragma unroll
  for (int s = 0; s < SLOTS; s++) {
    const int w = tid + KM_THREADS * s;
    if (w >= live) break;
    const int wg = w0 + w, a = wg / n, i = wg - a * n;
    const long long off = ((long long)a * k + c) * n + i;
    cnew[off] = cnt > 0 ? sum[s] / (double)cnt : cold[off];
  }
}

// shift = sum_c calc_distance(new_c, old_c): a thread per centroid walks its samples sequentially, as the reference
// does (GPmap.py:114-121), then thread 0 adds the k distances in centroid order.
__global__ void __launch_bounds__(1024)
kmeans_finish_kernel(const double* __restrict__ cold, const double* __restrict__ cnew, int n, int k,
                     double* __restrict__ shift_c, double threshold, KmState* state) {
  if (state->converged) return;
  extern __shared__ double km_d[];         // [warps][n]: the point distances of one centroid per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  double* d = km_d + (size_t)warp * n;
  for (int c = warp; c < k; c += nwarp) {  // the distances in parallel, their sum in sample order
    for (int i = lane; i < n; i += 32) {
      const double dx = cnew[(long long)c * n + i] - cold[(long long)c * n + i];
      const double dy = cnew[((long long)k + c) * n + i] - cold[((long long)k + c) * n + i];
      d[i] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    }
    __syncwarp();
    if (lane == 0) {
      double s = 0.0;
      for (int i = 0; i < n; i++) s = __dadd_rn(s, d[i]);
      shift_c[c] = s;
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double s = 0.0;
  for (int c = 0; c < k; c++) s = __dadd_rn(s, shift_c[c]);
  state->shift = s;
  state->iters += 1;
  if (s < threshold) state->converged = 1;
}

}  // namespace gpm

using namespace gpm;

extern "C" int gpm_kmeans_assign(gpm_handle_t h, const double* pxT, const double* pyT, int64_t P, int32_t n,
                                 const double* cx, const double* cy, int32_t k, double* dist,
                                 int32_t* assign, gpm_stream_t stream) {
  GPM_ARG(h != nullptr, 1);
  GPM_ARG(pxT != nullptr, 2);
  GPM_ARG(pyT != nullptr, 3);
  GPM_ARG(P > 0, 4);
  GPM_ARG(n > 0, 5);
  GPM_ARG(cx != nullptr, 6);
  GPM_ARG(cy != nullptr, 7);
  GPM_ARG(k > 0 && (size_t)k * n * 16 <= 200 * 1024, 8);
  GPM_ARG(assign != nullptr, 10);
  DeviceGuard guard(reinterpret_cast<gpm_handle_impl*>(h)->device);
  const size_t smem = (size_t)k * n * 16;
  if (smem > 48 * 1024)
    GPM_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kmeans_assign_kernel<<<(unsigned)((P + 127) / 128), 128, smem, (cudaStream_t)stream>>>(pxT, pyT, P, n, cx, cy, k, dist, assign, nullptr);
  GPM_LAUNCH_CHECK();
  return 0;
}

// workspace layout: second centroid buffer [3][k][n] | per-centroid shifts [k] | state (2 doubles) |
//   total [k], coff [k] (int64) | lrank [P], counts [nchunks][k], choff [nchunks][k] (int32, padded to 16 bytes) |
//   ordered row buffer G [(P + 2k + 2)][3n]
struct KmLayout {
  size_t shift_off, state_off, total_off, coff_off, lrank_off, counts_off, choff_off, g_off, bytes;
  long long nchunks, rows_total;
  int slices, W;            // the 3n row entries in `slices` slices of W (even) entries: slices * k CTAs ~ one per SM
};
static KmLayout km_layout(long long P, int n, int k) {
  KmLayout L;
  L.nchunks = (P + KM_CH - 1) / KM_CH;
  {
    const int work = 3 * n, want = std::max(1, 148 / k);   // B200: 148 SMs (k = 8, P = 100000: 8 / 32 / 64 / 148 / 296 / 600 CTAs give 0.70 / 0.47 / 0.42 / 0.40 / 0.39 / 0.39 ms per iteration)
    int W = (work + want - 1) / want;
    W = std::max(4, (W + 3) & ~3);                         // rows of whole 32-byte sectors (the gather writes them one by one)
    L.W = W;
    L.slices = (work + W - 1) / W;
    L.rows_total = P + 2 * (long long)k + 2;
  }
  size_t o = (size_t)3 * k * n * 8;
  L.shift_off = o; o += (size_t)k * 8;
  L.state_off = o; o += 16;
  L.total_off = o; o += (size_t)k * 8;
  L.coff_off = o; o += (size_t)k * 8;
  L.lrank_off = o; o += (size_t)P * 4;
  L.counts_off = o; o += (size_t)L.nchunks * k * 4;
  L.choff_off = o; o += (size_t)L.nchunks * k * 4;
  o = (o + 15) & ~(size_t)15;
  L.g_off = o; o += (size_t)L.rows_total * L.slices * L.W * 8;
  L.bytes = o;
  return L;
}

extern "C" size_t gpm_kmeans_workspace_bytes(int64_t P, int32_t n, int32_t k) {
  if (P <= 0 || n <= 0 || k <= 0) return 0;
  return km_layout(P, n, k).bytes;
}

extern "C" int gpm_kmeans_lloyd(gpm_handle_t h, const double* px, const double* py, const double* pt,
                                const double* pxT, const double* pyT, int64_t P, int32_t n, int32_t k,
                                double* centroids, int32_t* assign, double threshold, int32_t iters,
                                int32_t first, void* ws, gpm_stream_t stream) {
  GPM_ARG(h != nullptr, 1);
  GPM_ARG(px != nullptr && py != nullptr && pt != nullptr, 2);
  GPM_ARG(pxT != nullptr && pyT != nullptr, 5);
  GPM_ARG(P > 0, 7);
  GPM_ARG(n > 0 && 3 * n <= 2048, 8);
  GPM_ARG(k > 0 && (size_t)k * n * 16 <= 200 * 1024 && k <= 4096, 9);
  GPM_ARG(centroids != nullptr, 10);
  GPM_ARG(assign != nullptr, 11);
  GPM_ARG(iters >= 0, 13);
  GPM_ARG(ws != nullptr && ((uintptr_t)ws & 15) == 0, 15);
  DeviceGuard guard(reinterpret_cast<gpm_handle_impl*>(h)->device);
  cudaStream_t st = (cudaStream_t)stream;
  const KmLayout L = km_layout(P, n, k);
  char* wsb = reinterpret_cast<char*>(ws);
  double* buf1 = reinterpret_cast<double*>(wsb);
  double* shift_c = reinterpret_cast<double*>(wsb + L.shift_off);
  KmState* state = reinterpret_cast<KmState*>(wsb + L.state_off);
  long long* total = reinterpret_cast<long long*>(wsb + L.total_off);
  long long* coff = reinterpret_cast<long long*>(wsb + L.coff_off);
  int* lrank = reinterpret_cast<int*>(wsb + L.lrank_off);
  int* counts = reinterpret_cast<int*>(wsb + L.counts_off);
  int* choff = reinterpret_cast<int*>(wsb + L.choff_off);
  double* G = reinterpret_cast<double*>(wsb + L.g_off);
  if (first) GPM_CUDA(cudaMemsetAsync(state, 0, sizeof(KmState), st));
  const size_t smem = (size_t)k * n * 16;
  if (smem > 48 * 1024)
    GPM_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t plane = (size_t)k * n;
  // staged rows of the centroid sum: two stages of about 48 KB (an even number of rows of W doubles)
  int stage_rows = (int)std::max<size_t>(2, std::min<size_t>(4096, (48 * 1024) / ((size_t)L.W * 8)));
  stage_rows &= ~1;
  const size_t upd_smem = (size_t)2 * stage_rows * L.W * sizeof(double) + 16;
  GPM_CUDA(cudaFuncSetAttribute(kmeans_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  const size_t rank_smem = (size_t)(KM_WARPS + 1) * k * sizeof(int);
  if (rank_smem > 48 * 1024)
    GPM_CUDA(cudaFuncSetAttribute(kmeans_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rank_smem));
  const int fin_threads = 32 * std::min(32, k);             // a warp per centroid (up to 32 at a time)
  const size_t fin_smem = (size_t)(fin_threads / 32) * n * sizeof(double);
  if (fin_smem > 48 * 1024)
    GPM_CUDA(cudaFuncSetAttribute(kmeans_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
  const unsigned nchunks = (unsigned)L.nchunks;
  // Iterations ping-pong between `centroids` (even) and the workspace buffer (odd).  Once `converged` is set every
  // later kernel is a no-op, so state->iters tells the caller which buffer holds the final centroids; an even
  // number of enqueued iterations per call keeps the parity bookkeeping on the caller's side trivial.
  GPM_ARG((iters & 1) == 0, 13);
  for (int it = 0; it < iters; it++) {
    double* cur = (it & 1) ? buf1 : centroids;
    double* nxt = (it & 1) ? centroids : buf1;
    kmeans_assign_kernel<<<(unsigned)((P + 127) / 128), 128, smem, st>>>(pxT, pyT, P, n, cur, cur + plane, k, nullptr, assign, state);
    GPM_LAUNCH_CHECK();
    kmeans_rank_kernel<<<nchunks, KM_THREADS, rank_smem, st>>>(assign, P, k, lrank, counts, state);
    GPM_LAUNCH_CHECK();
    kmeans_scan_kernel<<<1, 1024, 0, st>>>(counts, (int)nchunks, k, choff, total, coff, state);
    GPM_LAUNCH_CHECK();
    kmeans_gather_kernel<<<(unsigned)((P + KM_WARPS - 1) / KM_WARPS), KM_THREADS, 0, st>>>(px, py, pt, P, n, k, assign, lrank, choff, coff, G, state, L.W, L.rows_total);
    GPM_LAUNCH_CHECK();
    kmeans_update_kernel<<<dim3(k, L.slices), KM_THREADS, upd_smem, st>>>(G, n, k, total, coff, cur, nxt, state, stage_rows, L.W, L.rows_total);
    GPM_LAUNCH_CHECK();
    kmeans_finish_kernel<<<1, fin_threads, fin_smem, st>>>(cur, nxt, n, k, shift_c, threshold, state);
    GPM_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int gpm_kmeans_state(gpm_handle_t h, const void* ws, int32_t n, int32_t k, int32_t* iters,
                                int32_t* converged, double* shift, gpm_stream_t stream) {
  GPM_ARG(h != nullptr, 1);
  GPM_ARG(ws != nullptr, 2);
  GPM_ARG(n > 0 && k > 0, 3);
  DeviceGuard guard(reinterpret_cast<gpm_handle_impl*>(h)->device);
  // the state sits right behind the second centroid buffer and the shifts, whatever P is
  const KmState* state = reinterpret_cast<const KmState*>(reinterpret_cast<const char*>(ws) + ((size_t)3 * k * n + k) * 8);
  KmState hs;
  GPM_CUDA(cudaMemcpyAsync(&hs, state, sizeof(KmState), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  GPM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (iters) *iters = hs.iters;
  if (converged) *converged = hs.converged;
  if (shift) *shift = hs.shift;
  return 0;
}
