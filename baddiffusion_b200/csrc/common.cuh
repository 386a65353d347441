// Shared helpers for libb200bd (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200bd.h"

namespace bd {

// thread-local last-error string behind bd_last_error()
void set_error(const char* fmt, ...);
const char* get_error();
void count_launch(int n);
unsigned long long launches();

#define BD_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      bd::set_error(__VA_ARGS__);               \
      return BD_ERR_INVALID;                    \
    }                                           \
  } while (0)

#define BD_CHECK_LAUNCH()                                                        \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      bd::set_error("%s:%d CUDA launch failed: %s", __FILE__, __LINE__,          \
                    cudaGetErrorString(e__));                                    \
      return BD_ERR_CUDA;                                                        \
    }                                                                            \
  } while (0)

__host__ __device__ static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

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
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16-byte vector of 8 halfs
struct __align__(16) half8 {
  __half2 a, b, c, d;
};

__device__ __forceinline__ void unpack8(const half8& v, float* f) {
  float2 t;
  t = __half22float2(v.a); f[0] = t.x; f[1] = t.y;
  t = __half22float2(v.b); f[2] = t.x; f[3] = t.y;
  t = __half22float2(v.c); f[4] = t.x; f[5] = t.y;
  t = __half22float2(v.d); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ half8 pack8(const float* f) {
  half8 v;
  v.a = __floats2half2_rn(f[0], f[1]);
  v.b = __floats2half2_rn(f[2], f[3]);
  v.c = __floats2half2_rn(f[4], f[5]);
  v.d = __floats2half2_rn(f[6], f[7]);
  return v;
}

}  // namespace bd
