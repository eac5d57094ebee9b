// Library-level entry points: version, thread-local error text, launch counter.
#include <cublas_v2.h>
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace gloria {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

// Statistics only (bench.py's gpu_launches): process-wide because autograd runs the backward on its own thread.
std::atomic<long long>& launch_counter() {
  static std::atomic<long long> n{0};
  return n;
}

// Caller-owned cudaEvent_t pairs recorded around the named kernels (bench.py's live roofline timing).
std::atomic<void*> g_timer[GLORIA_TIMER_SLOTS][2];

void timer_record(int slot, int which, cudaStream_t st) {
  void* ev = g_timer[slot][which].load(std::memory_order_relaxed);
  if (ev) cudaEventRecord((cudaEvent_t)ev, st);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

void* cublas_handle_opaque() {
  static thread_local cublasHandle_t h[16] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (!h[dev] && cublasCreate(&h[dev]) != CUBLAS_STATUS_SUCCESS) h[dev] = nullptr;
  return h[dev];
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(err_buf(), 512, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return GLORIA_ERR_CUDA_BASE + (int)e;
}

}  // namespace gloria

extern "C" int gloria_b200_version(void) { return 200; }

#ifndef GLORIA_BUILD_ID
#define GLORIA_BUILD_ID "unknown"
#endif
// sha256 prefix of the sources this binary was compiled from (build.py: source_id()); _lib.py compares it with the
// sources on disk, so a stale prebuilt library is refused instead of silently used
extern "C" const char* gloria_b200_build_id(void) { return GLORIA_BUILD_ID; }

extern "C" const char* gloria_b200_last_error(void) { return gloria::err_buf(); }

extern "C" long long gloria_b200_launch_count(int reset) {
  return reset ? gloria::launch_counter().exchange(0) : gloria::launch_counter().load();
}

extern "C" int gloria_b200_set_timer_events(int slot, void* start_event, void* stop_event) {
  if (slot < 0 || slot >= GLORIA_TIMER_SLOTS) return gloria::fail(GLORIA_ERR_BAD_ARG, "timer slot %d", slot);
  gloria::g_timer[slot][0].store(start_event);
  gloria::g_timer[slot][1].store(stop_event);
  return GLORIA_OK;
}

// cudaEventRecord for callers that hold only raw handles (the Python shim marks "d_ctx is final" on paths that do
// not go through gloria_b200_tc_local_sim_bwd_train_ev)
extern "C" int gloria_b200_record_event(void* event, void* stream) {
  GLORIA_CHECK_ARG(event != nullptr, "null event");
  GLORIA_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return GLORIA_OK;
}

// Small host int arrays (caption lengths) to the device THROUGH THE KERNEL PARAMETER BUFFER: no copy engine is involved,
// so the upload neither blocks the host (a pageable cudaMemcpyAsync does) nor queues behind a large asynchronous
// host-to-device copy of the next batch on another stream (measured: 10 ms per step at B = 48 in bench.py's e2e loop).
namespace gloria {
struct IntPack { int32_t v[960]; };
__global__ void upload_ints_kernel(int32_t* dst, IntPack p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = p.v[i];
}
}  // namespace gloria

extern "C" int gloria_b200_upload_ints(const int32_t* host, int n, int32_t* dev, void* stream) {
  GLORIA_CHECK_ARG(host != nullptr && dev != nullptr && n >= 0, "null pointer / negative count");
  for (int o = 0; o < n; o += 960) {
    gloria::IntPack p;
    const int m = n - o < 960 ? n - o : 960;
    memcpy(p.v, host + o, (size_t)m * sizeof(int32_t));
    gloria::upload_ints_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dev + o, p, m);
    GLORIA_LAUNCHED("upload_ints_kernel");
  }
  return GLORIA_OK;
}
