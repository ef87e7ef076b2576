// Shared helpers for libgwen_b200: error convention, small device utilities.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "gwen_b200.h"

namespace gwen {

// Thread-local last-error text (gwen_last_error()).
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define GWEN_CHECK_ARG(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) return ::gwen::set_err(GWEN_E_BADARG, __VA_ARGS__); \
  } while (0)

#define GWEN_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return ::gwen::set_err(GWEN_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

// After a kernel launch: report launch-configuration errors without synchronising.
#define GWEN_LAUNCH_CHECK(name)                                                           \
  do {                                                                                    \
    cudaError_t e_ = cudaPeekAtLastError();                                               \
    if (e_ != cudaSuccess) {                                                              \
      cudaGetLastError();                                                                 \
      return ::gwen::set_err(GWEN_E_CUDA, "launch of %s failed: %s", name,                \
                             cudaGetErrorString(e_));                                     \
    }                                                                                     \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count();  // cached multiProcessorCount of the current device (148 on B200)
int sm_reserve();  // SMs the persistent kernels leave free (gwen_set_sm_reserve)

template <typename T>
struct DType;
template <>
struct DType<float> {
  static constexpr int code = GWEN_F32;
};
template <>
struct DType<__nv_bfloat16> {
  static constexpr int code = GWEN_BF16;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

}  // namespace gwen
