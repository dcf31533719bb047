// Handle, error text and TMA tensor-map construction for libgpmap_b200.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gpm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  return (int)e;
}

int make_tmap(gpm_handle_impl* h, CUtensorMap* map, const double* base, int64_t rows, int64_t cols,
              int64_t ld, int rows_box) {
  if (((uintptr_t)base & 15) != 0 || (ld & 1) != 0 || rows <= 0 || cols <= 0) {
    set_error("tensor map: base must be 16-byte aligned and ld even (base=%p ld=%lld rows=%lld cols=%lld)",
              (const void*)base, (long long)ld, (long long)rows, (long long)cols);
    return -1;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {SLAB_K, (cuuint32_t)rows_box};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
              (long long)rows, (long long)cols, (long long)ld);
    return 999;
  }
  return 0;
}

}  // namespace gpm

using namespace gpm;

extern "C" {

int gpm_version(void) { return GPM_VERSION; }

long long gpm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* gpm_last_error(void) { return g_err; }

int gpm_create(gpm_handle_t* handle, int device) {
  GPM_ARG(handle != nullptr, 1);
  *handle = nullptr;
  int count = 0;
  GPM_CUDA(cudaGetDeviceCount(&count));
  GPM_ARG(device >= 0 && device < count, 2);
  GPM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  GPM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("libgpmap_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return (int)cudaErrorNoKernelImageForDevice;
  }
  gpm_handle_impl* h = new gpm_handle_impl();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->ev = nullptr;
  h->n_ev = 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete h;
    set_error("cuTensorMapEncodeTiled entry point not available");
    return e != cudaSuccess ? (int)e : 999;
  }
  h->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  int lo = 0, hi = 0;
  GPM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  GPM_CUDA(cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, hi));
  h->n_ev = 4;
  h->ev = new cudaEvent_t[h->n_ev];
  for (int i = 0; i < h->n_ev; i++) GPM_CUDA(cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming));
  h->n_flags = 8192;                  // up to N = 2^20
  GPM_CUDA(cudaMalloc(&h->flags, 2 * h->n_flags * sizeof(int)));
  GPM_CUDA(cudaMemset(h->flags, 0, 2 * h->n_flags * sizeof(int)));
  *handle = reinterpret_cast<gpm_handle_t>(h);
  return 0;
}

int gpm_destroy(gpm_handle_t handle) {
  if (!handle) return 0;
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  cudaSetDevice(h->device);
  for (int i = 0; i < h->n_ev; i++) cudaEventDestroy(h->ev[i]);
  delete[] h->ev;
  cudaStreamDestroy(h->aux);
  cudaFree(h->flags);
  delete h;
  return 0;
}

int gpm_sm_count(gpm_handle_t handle) {
  if (!handle) return 0;
  return reinterpret_cast<gpm_handle_impl*>(handle)->sm_count;
}

}  // extern "C"
