// FP64 peak microbenchmarks for B200 (sm_100a): DFMA, DMMA (m8n8k4 / m16n8k16), mixed, exp().
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_peaks tools/fp64_peaks.cu
// Prints one JSON object on stdout. Measurement tool only; not on the product path.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma884(double *out, int iters, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma16816(double *out, int iters, double a, double b) {
  double c[NACC][4];
  double af[8], bf[4];
#pragma unroll
  for (int i = 0; i < 8; i++) af[i] = a + i;
#pragma unroll
  for (int i = 0; i < 4; i++) bf[i] = b - i;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma16816(c[i], af, bf);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: per iteration NACC DMMA.884 + NF DFMA, to see whether they share a pipe
template <int NACC, int NF>
__global__ void k_mixed(double *out, int iters, double a, double b) {
  double c0[NACC], c1[NACC], f[NF];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = i; c1[i] = -i; }
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = fma(f[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double *out, int iters, double x0) {
  double s = 0, x = -x0 * (threadIdx.x + 1) * 1e-3;
  for (int it = 0; it < iters; it++) {
    s += exp(x); x -= 1e-6;
    s += exp(x * 1.5); 
    s += exp(x * 2.5);
    s += exp(x * 3.5);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_write(double2 *out, size_t n, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double2 w = make_double2(v, v + 1);
  for (; i < n; i += stride) out[i] = w;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM (one block per SM... use blocks of 128 thr)
    int threads = 128, blocks = sms * wps / 4;
    { float ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 8 * iters * (double)blocks * threads; printf(", \"dfma_tflops_w%d\": %.3f", wps, fl / ms * 1e-9); }
    { float ms = time_ms([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 256 * 8 * iters * (double)blocks * threads / 32; printf(", \"dmma884_tflops_w%d\": %.3f", wps, fl / ms * 1e-9); }
    { float ms = time_ms([&] { k_dmma16816<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 16 * 8 * 16 * 4 * iters * (double)blocks * threads / 32; printf(", \"dmma16816_tflops_w%d\": %.3f", wps, fl / ms * 1e-9); }
  }
  { int threads = 128, blocks = sms * 4;
    float ms = time_ms([&] { k_mixed<8, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fl = (2.0 * 256 * 8 / 32 + 2.0 * 8) * iters * (double)blocks * threads; printf(", \"mixed_8dmma_8dfma_tflops_w16\": %.3f", fl / ms * 1e-9);
    ms = time_ms([&] { k_mixed<8, 32><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    fl = (2.0 * 256 * 8 / 32 + 2.0 * 32) * iters * (double)blocks * threads; printf(", \"mixed_8dmma_32dfma_tflops_w16\": %.3f", fl / ms * 1e-9); }
  { int threads = 256, blocks = sms * 8;
    float ms = time_ms([&] { k_exp<<<blocks, threads>>>(out, 1024, 1.0); });
    double ev = 4.0 * 1024 * (double)blocks * threads; printf(", \"exp_gevals_per_s\": %.2f", ev / ms * 1e-6); }
  { size_t bytes = (size_t)4 << 30; double2 *buf; CK(cudaMalloc(&buf, bytes));
    float ms = time_ms([&] { k_write<<<sms * 16, 512>>>(buf, bytes / 16, 1.0); });
    printf(", \"hbm_write_gbs\": %.1f", bytes / ms * 1e-6);
    ms = time_ms([&] { CK(cudaMemsetAsync(buf, 0, bytes)); });
    printf(", \"memset_gbs\": %.1f", bytes / ms * 1e-6);
    CK(cudaFree(buf)); }
  printf("}\n");
  return 0;
}
