// Handle, error text and TMA tensor-map construction for libgpmap_b200.
#include <atomic>
#include <stdarg.h>
#include <ctype.h>
#include <stdlib.h>
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

// name -> member table of the switches (gpm_set_option / gpm_get_option; GPM_<NAME> in the environment at gpm_create)
struct OptEntry { const char* name; int Options::*member; };
static const OptEntry kOptions[] = {
    {"no_lookahead", &Options::no_lookahead},     {"tpc_wide", &Options::tpc_wide},
    {"tpc_narrow", &Options::tpc_narrow},         {"wide_min", &Options::wide_min},
    {"wide4_min", &Options::wide4_min},           {"wide8_min", &Options::wide8_min},
    {"no_separable", &Options::no_separable},     {"no_small_fused", &Options::no_small_fused},
    {"no_small_tiles", &Options::no_small_tiles}, {"no_fused_fwd", &Options::no_fused_fwd},
    {"no_fused_mean", &Options::no_fused_mean},   {"var_steps", &Options::var_steps},
    {"solve_steps", &Options::solve_steps},       {"grad_sweep", &Options::grad_sweep},
    {"no_path_fused", &Options::no_path_fused},   {"no_fused_solve", &Options::no_fused_solve},
    {"path_fused", &Options::path_fused},         {"no_half_tiles", &Options::no_half_tiles},
    {"half_warps", &Options::half_warps},          {"half_stages", &Options::half_stages},
    {"no_split_column", &Options::no_split_column}, {"no_scratch_factor", &Options::no_scratch_factor},
    {"batch_width", &Options::batch_width},      {"small_two_max", &Options::small_two_max},
};

static void options_from_env(Options* o) {
  for (const OptEntry& e : kOptions) {
    char env[64] = "GPM_";
    size_t n = 4;
    for (const char* c = e.name; *c && n + 1 < sizeof(env); c++) env[n++] = (char)toupper((unsigned char)*c);
    env[n] = 0;
    const char* v = getenv(env);
    if (v && *v) o->*(e.member) = atoi(v);
  }
}

}  // namespace gpm

using namespace gpm;

extern "C" {

int gpm_version(void) { return GPM_VERSION; }

int gpm_set_option(gpm_handle_t handle, const char* name, int value) {
  GPM_ARG(handle != nullptr, 1);
  GPM_ARG(name != nullptr, 2);
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  for (const OptEntry& e : kOptions)
    if (strcmp(e.name, name) == 0) { h->opt.*(e.member) = value; return 0; }
  set_error("unknown option '%s'", name);
  return -2;
}

int gpm_get_option(gpm_handle_t handle, const char* name, int* value) {
  GPM_ARG(handle != nullptr, 1);
  GPM_ARG(name != nullptr, 2);
  GPM_ARG(value != nullptr, 3);
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  for (const OptEntry& e : kOptions)
    if (strcmp(e.name, name) == 0) { *value = h->opt.*(e.member); return 0; }
  set_error("unknown option '%s'", name);
  return -2;
}

long long gpm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* gpm_last_error(void) { return g_err; }

static int create_impl(gpm_handle_impl* h, int device) {
  cudaDeviceProp prop;
  GPM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("libgpmap_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return (int)cudaErrorNoKernelImageForDevice;
  }
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  options_from_env(&h->opt);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return e != cudaSuccess ? (int)e : 999;
  }
  h->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  int lo = 0, hi = 0;
  GPM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  GPM_CUDA(cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, hi));
  h->ev = new cudaEvent_t[4];
  for (int i = 0; i < 4; i++) {
    GPM_CUDA(cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming));
    h->n_ev = i + 1;
  }
  h->n_flags = 8192;                  // up to N = 2^20
  // (no memset here: every solve clears the flags on its own stream, and a synchronous memset would invalidate a
  //  CUDA-graph capture in progress -- the Python binding creates the handle of a new stream on first use)
  GPM_CUDA(cudaMalloc(&h->flags, 3 * h->n_flags * sizeof(int)));
  GPM_CUDA(cudaMalloc(&h->lml_part, (size_t)h->n_flags * 9 * sizeof(double)));
  return 0;
}

static void destroy_impl(gpm_handle_impl* h) {
  for (int i = 0; i < h->n_ev; i++) cudaEventDestroy(h->ev[i]);
  delete[] h->ev;
  if (h->aux) cudaStreamDestroy(h->aux);
  if (h->flags) cudaFree(h->flags);
  if (h->lml_part) cudaFree(h->lml_part);
  delete h;
}

// A handle is SINGLE-STREAM and SINGLE-THREAD state: its helper stream, event pool and the flag arrays of the
// chained solves are reused by every call, so calls that may overlap in time (other streams, other host threads)
// need handles of their own.  The caller's current device is left as it was.
int gpm_create(gpm_handle_t* handle, int device) {
  GPM_ARG(handle != nullptr, 1);
  *handle = nullptr;
  int count = 0;
  GPM_CUDA(cudaGetDeviceCount(&count));
  GPM_ARG(device >= 0 && device < count, 2);
  int prev = -1;
  GPM_CUDA(cudaGetDevice(&prev));
  GPM_CUDA(cudaSetDevice(device));
  gpm_handle_impl* h = new gpm_handle_impl();
  h->ev = nullptr; h->n_ev = 0; h->aux = nullptr; h->flags = nullptr; h->lml_part = nullptr;
  // The binding creates the handle of a stream on first use, which may be inside a CUDA-graph capture; allocations
  // are "potentially unsafe" calls under the default (global) capture mode, so this thread is switched to relaxed
  // mode for the duration (as PyTorch's allocator does).  Nothing here enqueues work on a capturing stream.
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  const bool swapped = cudaThreadExchangeStreamCaptureMode(&mode) == cudaSuccess;
  const int rc = create_impl(h, device);
  if (swapped) cudaThreadExchangeStreamCaptureMode(&mode);
  if (rc) destroy_impl(h);
  else *handle = reinterpret_cast<gpm_handle_t>(h);
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  return rc;
}

int gpm_destroy(gpm_handle_t handle) {
  if (!handle) return 0;
  gpm_handle_impl* h = reinterpret_cast<gpm_handle_impl*>(handle);
  DeviceGuard guard(h->device);
  destroy_impl(h);
  return 0;
}

int gpm_sm_count(gpm_handle_t handle) {
  if (!handle) return 0;
  return reinterpret_cast<gpm_handle_impl*>(handle)->sm_count;
}

}  // extern "C"
