// Shared host/device helpers for libgloria_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gloria_b200.h"

namespace gloria {

// thread-local error text + launch counter (no mutable process-global state: entry points stay re-entrant)
char* err_buf();
std::atomic<long long>& launch_counter();
void timer_record(int slot, int which, cudaStream_t st);   // which: 0 = before, 1 = after the kernel
// one cuBLAS handle per host thread and device, created on first use (plain library GEMMs only); returned as void*
// so that only the translation units that call cuBLAS include its header
void* cublas_handle_opaque();

// tc_f32.cu: scores of the diagonal pairs on the split-precision tensor-core GEMM (0 bytes = shape not covered)
size_t f32tc_diag_scores_workspace(int B, int D, int S, int Lcap);
int f32tc_diag_scores(const float* ctx, const float* wt32, const int32_t* cap_lens, int B, int D, int S, int Lw, int Lcap, int off,
                      float* sc, void* ws, size_t ws_bytes, cudaStream_t st);

int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define GLORIA_CHECK_ARG(cond, ...)                                   \
  do {                                                                \
    if (!(cond)) return ::gloria::fail(GLORIA_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define GLORIA_CUDA(expr)                                             \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) return ::gloria::cuda_fail(_e, #expr);     \
  } while (0)

// call after every kernel launch
#define GLORIA_LAUNCHED(name)                                         \
  do {                                                                \
    ++::gloria::launch_counter();                                     \
    cudaError_t _e = cudaPeekAtLastError();                           \
    if (_e != cudaSuccess) return ::gloria::cuda_fail(_e, name);      \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace gloria
