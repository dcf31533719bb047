// SURVEY.md section 8f "next" #1: the reference's real hot loop, trajectories.calc_distance inside
// kmeansclustering (GPmap.py:72-80, 114-121).  dist[p,c] = sum_i hypot-free sqrt(dx^2+dy^2) summed
// sequentially in sample order (as the reference's Python float accumulation does), and the
// first-minimum assignment with strict '<' (GPmap.py:76).  One thread per (path, centroid) pair.
#include "common.cuh"

namespace gpm {

__global__ void __launch_bounds__(128)
kmeans_assign_kernel(const double* __restrict__ px, const double* __restrict__ py, long long P, int n,
                     const double* __restrict__ cx, const double* __restrict__ cy, int k,
                     double* __restrict__ dist, int* __restrict__ assign) {
  extern __shared__ double sc[];          // centroids: cx[k][n] then cy[k][n]
  for (int e = threadIdx.x; e < k * n; e += blockDim.x) { sc[e] = cx[e]; sc[k * n + e] = cy[e]; }
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double best = 0.0;
  int best_c = 0;
  for (int c = 0; c < k; c++) {
    double s = 0.0;
    for (int i = 0; i < n; i++) {
      const double dx = px[p * n + i] - sc[c * n + i];
      const double dy = py[p * n + i] - sc[k * n + c * n + i];
      s = __dadd_rn(s, sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
    }
    if (dist) dist[p * k + c] = s;
    if (c == 0 || s < best) { best = s; best_c = c; }
  }
  if (assign) assign[p] = best_c;
}

}  // namespace gpm

using namespace gpm;

extern "C" int gpm_kmeans_assign(gpm_handle_t h, const double* px, const double* py, int64_t P, int32_t n,
                                 const double* cx, const double* cy, int32_t k, double* dist,
                                 int32_t* assign, gpm_stream_t stream) {
  GPM_ARG(h != nullptr, 1);
  GPM_ARG(px != nullptr, 2);
  GPM_ARG(py != nullptr, 3);
  GPM_ARG(P > 0, 4);
  GPM_ARG(n > 0, 5);
  GPM_ARG(cx != nullptr, 6);
  GPM_ARG(cy != nullptr, 7);
  GPM_ARG(k > 0 && (size_t)k * n * 16 <= 200 * 1024, 8);
  const size_t smem = (size_t)k * n * 16;
  if (smem > 48 * 1024)
    GPM_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kmeans_assign_kernel<<<(unsigned)((P + 127) / 128), 128, smem, (cudaStream_t)stream>>>(px, py, P, n, cx, cy, k, dist, assign);
  GPM_LAUNCH_CHECK();
  return 0;
}
