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
  for (int c = 0; c < k; c++) {
    const double* ccx = sc + c * n;
    const double* ccy = sc + k * n + c * n;
    double s = 0.0;
#pragma unroll 4
    for (int i = 0; i < n; i++) {
      const double dx = pxT[(long long)i * P + p] - ccx[i];
      const double dy = pyT[(long long)i * P + p] - ccy[i];
      s = __dadd_rn(s, sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
    }
    if (dist) dist[p * k + c] = s;
    if (c == 0 || s < best) { best = s; best_c = c; }
  }
  assign[p] = best_c;
}

// One CTA per centroid; thread (a, i) owns sample i of array a in {xs, ys, timestamp}.  px, py, pt: [P][n].
// cold / cnew: [3][k][n] (xs, ys, timestamp planes).  An empty cluster keeps its previous centroid.
//
// The additions of a centroid's members are a serial chain by construction (path order, the order the reference
// appends them to the cluster list: calc_mean_traj is bit-exact that way); the LOADS need not be.  Paths are scanned in
// chunks of 2048: the eight warps compact the chunk's members with ballots into one ordered list, the members' rows are
// then gathered into shared memory by all threads (a warp per row) with asynchronous 8-byte copies (stage_rows rows per stage, two
// stages: the gather of group g + 1 flies while group g is summed), and the 3n summing threads add the staged rows in
// order.  The first version loaded each member's value right before its addition: one DRAM round trip per member
// (ncu: 13 GB/s, issue slots 2 % busy), 5.1 ms per Lloyd iteration at P = 100000, k = 8; this one 1.1 ms (assignment
// 0.085 ms of it).  What is left is the gather itself on k SMs; gathering with all SMs into an ordered per-centroid
// buffer first (a global order-preserving compaction) is the next step.
constexpr int KM_CH = 2048;

__device__ __forceinline__ void km_cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int KM_THREADS = 256, KM_WARPS = KM_THREADS / 32;    // all warps gather, the first 3n threads also sum (1024 threads measured no faster)

__global__ void __launch_bounds__(KM_THREADS, 1)
kmeans_update_kernel(const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pt,
                     long long P, int n, int k, const int* __restrict__ assign, const double* __restrict__ cold,
                     double* __restrict__ cnew, double* __restrict__ shift_c, const KmState* __restrict__ state,
                     int stage_rows) {
  if (state->converged) return;
  extern __shared__ __align__(16) double km_stage[];      // [2][stage_rows][3n]
  constexpr int WCH = KM_CH / KM_WARPS;
  __shared__ int members[KM_CH];
  __shared__ int wcnt[KM_WARPS];
  __shared__ long long total;
  const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int work = 3 * n;
  if (tid == 0) total = 0;
  // each thread carries up to SLOTS running sums (sample slots tid, tid + KM_THREADS, ...)
  constexpr int SLOTS = 2048 / KM_THREADS; // 3 n <= 2048
  double sum[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; s++) sum[s] = 0.0;
  for (long long p0 = 0; p0 < P; p0 += KM_CH) {
    unsigned masks[WCH / 32];
    int mw = 0;
    const long long base = p0 + (long long)warp * WCH;
#pragma unroll
    for (int r = 0; r < WCH / 32; r++) {
      const long long pp = base + r * 32 + lane;
      masks[r] = __ballot_sync(0xffffffffu, pp < P && assign[pp] == c);
      mw += __popc(masks[r]);
    }
    __syncthreads();                       // the previous chunk's list and stages have been consumed
    if (lane == 0) wcnt[warp] = mw;
    __syncthreads();
    int off = 0, m = 0;
#pragma unroll
    for (int w = 0; w < KM_WARPS; w++) { if (w < warp) off += wcnt[w]; m += wcnt[w]; }
#pragma unroll
    for (int r = 0; r < WCH / 32; r++) {
      if (masks[r] & (1u << lane)) members[off + __popc(masks[r] & ((1u << lane) - 1u))] = warp * WCH + r * 32 + lane;
      off += __popc(masks[r]);
    }
    __syncthreads();
    if (tid == 0) total += m;
    // gather rows [g0, g0 + cnt) of the member list into stage `buf` (asynchronous copies, one commit group)
    auto gather = [&](int buf, int g0, int cnt) {
      const uint32_t dst0 = smem_u32(km_stage + (size_t)buf * stage_rows * work);
      for (int row = warp; row < cnt; row += KM_WARPS) {   // a warp per member row: no index divisions in the loop
        const long long roff = (p0 + members[g0 + row]) * (long long)n;
        const uint32_t d = dst0 + (uint32_t)(row * work) * 8u;
        for (int i = lane; i < n; i += 32) {
          km_cp_async8(d + (uint32_t)i * 8u, px + roff + i);
          km_cp_async8(d + (uint32_t)(n + i) * 8u, py + roff + i);
          km_cp_async8(d + (uint32_t)(2 * n + i) * 8u, pt + roff + i);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int ngroups = (m + stage_rows - 1) / stage_rows;
    if (ngroups > 0) gather(0, 0, min(stage_rows, m));
    for (int g = 0; g < ngroups; g++) {
      if (g + 1 < ngroups) {
        gather((g + 1) & 1, (g + 1) * stage_rows, min(stage_rows, m - (g + 1) * stage_rows));
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();                     // every thread's copies of group g have landed
      const int cnt = min(stage_rows, m - g * stage_rows);
      const double* sb = km_stage + (size_t)(g & 1) * stage_rows * work;
#pragma unroll 1
      for (int s = 0; s < SLOTS; s++) {
        const int w = tid + KM_THREADS * s;
        if (w >= work) break;
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < SLOTS; t++) if (t == s) acc = sum[t];
        const double* col = sb + w;
#pragma unroll 8
        for (int r = 0; r < cnt; r++) acc = __dadd_rn(acc, col[(size_t)r * work]);
#pragma unroll
        for (int t = 0; t < SLOTS; t++) if (t == s) sum[t] = acc;
      }
      __syncthreads();                     // stage g & 1 may be refilled (by the gather of group g + 2)
    }
  }
  __syncthreads();
  const long long cnt = total;
#pragma unroll
  for (int s = 0; s < SLOTS; s++) {
    const int w = tid + KM_THREADS * s;
    if (w >= work) break;
    const int a = w / n, i = w - a * n;
    const long long off = ((long long)a * k + c) * n + i;
    cnew[off] = cnt > 0 ? sum[s] / (double)cnt : cold[off];
  }
  __syncthreads();
  if (tid == 0) {                          // calc_distance(new_c, old_c): sequential, as the reference
    double s = 0.0;
    for (int i = 0; i < n; i++) {
      const double dx = cnew[(long long)c * n + i] - cold[(long long)c * n + i];
      const double dy = cnew[((long long)k + c) * n + i] - cold[((long long)k + c) * n + i];
      s = __dadd_rn(s, sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
    }
    shift_c[c] = s;
  }
}

__global__ void kmeans_finish_kernel(const double* __restrict__ shift_c, int k, double threshold, KmState* state) {
  if (threadIdx.x != 0 || state->converged) return;
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

extern "C" size_t gpm_kmeans_workspace_bytes(int64_t P, int32_t n, int32_t k) {
  if (P <= 0 || n <= 0 || k <= 0) return 0;
  // second centroid buffer [3][k][n] + per-centroid shifts [k] + state
  return ((size_t)3 * k * n + (size_t)k + 2) * sizeof(double);
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
  GPM_ARG(k > 0 && (size_t)k * n * 16 <= 200 * 1024, 9);
  GPM_ARG(centroids != nullptr, 10);
  GPM_ARG(assign != nullptr, 11);
  GPM_ARG(iters >= 0, 13);
  GPM_ARG(ws != nullptr && ((uintptr_t)ws & 7) == 0, 15);
  DeviceGuard guard(reinterpret_cast<gpm_handle_impl*>(h)->device);
  cudaStream_t st = (cudaStream_t)stream;
  double* buf1 = reinterpret_cast<double*>(ws);
  double* shift_c = buf1 + (size_t)3 * k * n;
  KmState* state = reinterpret_cast<KmState*>(shift_c + k);
  if (first) GPM_CUDA(cudaMemsetAsync(state, 0, sizeof(KmState), st));
  const size_t smem = (size_t)k * n * 16;
  if (smem > 48 * 1024)
    GPM_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t plane = (size_t)k * n;
  // staged member rows of the centroid update: two stages of up to 128 rows of 3n doubles, within 200 KB
  const int stage_rows = (int)std::max<size_t>(1, std::min<size_t>(128, (100 * 1024) / ((size_t)3 * n * 8)));
  const size_t upd_smem = (size_t)2 * stage_rows * 3 * n * sizeof(double);
  GPM_CUDA(cudaFuncSetAttribute(kmeans_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  // Iterations ping-pong between `centroids` (even) and the workspace buffer (odd).  Once `converged` is set every
  // later kernel is a no-op, so state->iters tells the caller which buffer holds the final centroids; an even
  // number of enqueued iterations per call keeps the parity bookkeeping on the caller's side trivial.
  GPM_ARG((iters & 1) == 0, 13);
  for (int it = 0; it < iters; it++) {
    double* cur = (it & 1) ? buf1 : centroids;
    double* nxt = (it & 1) ? centroids : buf1;
    kmeans_assign_kernel<<<(unsigned)((P + 127) / 128), 128, smem, st>>>(pxT, pyT, P, n, cur, cur + plane, k, nullptr, assign, state);
    GPM_LAUNCH_CHECK();
    kmeans_update_kernel<<<k, KM_THREADS, upd_smem, st>>>(px, py, pt, P, n, k, assign, cur, nxt, shift_c, state, stage_rows);
    GPM_LAUNCH_CHECK();
    kmeans_finish_kernel<<<1, 32, 0, st>>>(shift_c, k, threshold, state);
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
  const double* buf1 = reinterpret_cast<const double*>(ws);
  const KmState* state = reinterpret_cast<const KmState*>(buf1 + (size_t)3 * k * n + k);
  KmState hs;
  GPM_CUDA(cudaMemcpyAsync(&hs, state, sizeof(KmState), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  GPM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (iters) *iters = hs.iters;
  if (converged) *converged = hs.converged;
  if (shift) *shift = hs.shift;
  return 0;
}
