// Dependent-issue latency of DFMA / DMMA.8x8x4 / rsqrt on sm_100a: one warp per SM, NACC independent chains.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC> __global__ void k_dmma(double* out, long long* cyc, int iters, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = i; c1[i] = -i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NACC> __global__ void k_dfma(double* out, long long* cyc, int iters, double a, double b) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rsqrt(double* out, long long* cyc, int iters, double a) {
  double c = a + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) c = rsqrt(c) + 2.0;
  long long t1 = clock64();
  out[threadIdx.x] = c;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_chain(double* out, long long* cyc, int iters) {
  __shared__ int nxt[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) nxt[i] = (i + 33) & 255;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) p = nxt[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
  const int it = 4096;
#define RUN(name, kern, threads, per) kern<<<1, threads>>>(out, cyc, it, 1.0000001, 1e-9); cudaDeviceSynchronize(); printf("%s: %.1f cycles per step\n", name, (double)cyc[0] / it / per);
  RUN("dmma chain, 1 warp, 1 acc ", k_dmma<1>, 32, 1)
  RUN("dmma chain, 1 warp, 2 acc (per dmma)", k_dmma<2>, 32, 2)
  RUN("dmma chain, 1 warp, 4 acc (per dmma)", k_dmma<4>, 32, 4)
  RUN("dmma chain, 1 warp, 8 acc (per dmma)", k_dmma<8>, 32, 8)
  RUN("dmma chain, 4 warps (1/SMSP), 1 acc", k_dmma<1>, 128, 1)
  RUN("dmma chain, 16 warps (4/SMSP), 1 acc", k_dmma<1>, 512, 1)
  RUN("dmma chain, 16 warps (4/SMSP), 2 acc (per dmma)", k_dmma<2>, 512, 2)
  RUN("dfma chain, 1 warp, 1 acc", k_dfma<1>, 32, 1)
  RUN("dfma chain, 1 warp, 4 acc (per fma)", k_dfma<4>, 32, 4)
  RUN("dfma chain, 1 thread, 1 acc", k_dfma<1>, 1, 1)
  k_rsqrt<<<1, 1>>>(out, cyc, it, 3.0); cudaDeviceSynchronize(); printf("rsqrt+add chain, 1 thread: %.1f cycles per step\n", (double)cyc[0] / it);
  k_lds_chain<<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); printf("LDS pointer chase: %.1f cycles per step\n", (double)cyc[0] / it);
  return 0;
}
