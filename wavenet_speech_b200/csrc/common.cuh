// Shared helpers for libwnb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/wnb200.h"

namespace wnb {

void set_error(const char* fmt, ...);

#define WNB_CHECK_ARG(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      wnb::set_error(__VA_ARGS__);    \
      return 1;                       \
    }                                 \
  } while (0)

#define WNB_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      wnb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                        \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define WNB_LAUNCH_OK()                                                                    \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      wnb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                            \
      return 3;                                                                            \
    }                                                                                      \
  } while (0)

// Every argument struct of the C-ABI starts with `uint32_t struct_size` = sizeof(the struct) as the CALLER compiled it:
// a host built against an older header (shorter struct) is refused instead of having garbage read as pointers.
#define WNB_CHECK_STRUCT(a, type, what)                                                                          \
  WNB_CHECK_ARG((a)->struct_size == sizeof(type), what ": struct_size %u != sizeof(" #type ") = %u (header / "   \
                "binding out of date?)", (unsigned)(a)->struct_size, (unsigned)sizeof(type))

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel instantiation,
// device), thread-safely (one process may drive several GPUs from several threads).
#define WNB_SET_SMEM_ATTR(bytes, ...)                                                                          \
  do {                                                                                                         \
    static std::atomic<unsigned long long> done_{0};                                                           \
    int dev_ = 0;                                                                                              \
    cudaGetDevice(&dev_);                                                                                      \
    const unsigned long long bit_ = (dev_ >= 0 && dev_ < 64) ? (1ull << dev_) : 0ull;                          \
    if (!bit_ || !(done_.load(std::memory_order_acquire) & bit_)) {                                            \
      WNB_CUDA_OK(cudaFuncSetAttribute((__VA_ARGS__), cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)));  \
      done_.fetch_or(bit_, std::memory_order_release);                                                         \
    }                                                                                                          \
  } while (0)

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : 0.01f * v; }
__device__ __forceinline__ float sigmoid_precise(float v) { return 1.f / (1.f + expf(-v)); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace wnb
