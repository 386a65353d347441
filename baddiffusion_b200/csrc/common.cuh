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

// 16-byte vector of 8 halfs.  The payload is a uint4 on purpose: __half2 has user-provided copy operations, so a struct of
// four __half2 is copied member by member and every load / store of it becomes four 32-bit accesses (LDG.E x4) instead
// of one LDG.E.128 -- 4x the load/store instructions and L1 wavefronts in every streaming kernel.
struct __align__(16) half8 {
  uint4 u;
};

__device__ __forceinline__ float2 h2f2_bits(unsigned int w) {
  __half2 h;
  *reinterpret_cast<unsigned int*>(&h) = w;
  return __half22float2(h);
}
__device__ __forceinline__ unsigned int f2h2_bits(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const unsigned int*>(&h);
}
__device__ __forceinline__ void unpack8(const half8& v, float* f) {
  float2 t;
  t = h2f2_bits(v.u.x); f[0] = t.x; f[1] = t.y;
  t = h2f2_bits(v.u.y); f[2] = t.x; f[3] = t.y;
  t = h2f2_bits(v.u.z); f[4] = t.x; f[5] = t.y;
  t = h2f2_bits(v.u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ half8 pack8(const float* f) {
  half8 v;
  v.u.x = f2h2_bits(f[0], f[1]);
  v.u.y = f2h2_bits(f[2], f[3]);
  v.u.z = f2h2_bits(f[4], f[5]);
  v.u.w = f2h2_bits(f[6], f[7]);
  return v;
}

}  // namespace bd
