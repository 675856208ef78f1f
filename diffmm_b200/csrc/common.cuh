// Shared host/device helpers for the diffmm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/diffmm_b200.h"

struct dmm_ctx {
  int device;
  int num_sms;
  int max_smem_optin;
  // cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency)
  void* encode_tiled;
};

void dmm_set_error(const char* fmt, ...);

#define DMM_CHECK_ARG(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      dmm_set_error(__VA_ARGS__);             \
      return DMM_ERR_INVALID;                 \
    }                                         \
  } while (0)

#define DMM_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      dmm_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DMM_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define DMM_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      dmm_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DMM_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// One-time per-device setup (cudaFuncSetAttribute and friends are per device): thread-safe, keyed by ctx->device.
// The guarded calls are idempotent, so two threads racing through `need` at worst repeat them.
struct DmmPerDeviceOnce {
  std::atomic<uint64_t> done{0};
  bool need(const dmm_ctx* c) const { return !((done.load(std::memory_order_acquire) >> (c->device & 63)) & 1ull); }
  void mark(const dmm_ctx* c) { done.fetch_or(1ull << (c->device & 63), std::memory_order_release); }
};

static inline int64_t dmm_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------
// bf16 split helpers: x ~= hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t dmm_bf16_bits(float x) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float dmm_bf16_to_f32(uint16_t b) {
  return __uint_as_float(((uint32_t)b) << 16);
}
__device__ __forceinline__ void dmm_split_bf16(float x, uint16_t& hi, uint16_t& lo) {
  hi = dmm_bf16_bits(x);
  lo = dmm_bf16_bits(x - dmm_bf16_to_f32(hi));
}

__device__ __forceinline__ float dmm_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dmm_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif
