// GroupNorm (+SiLU) forward / backward on fp16 NHWC views with fp32/fp64 statistics (K4/K5).
// HBM-bound: forward reads x twice (second read is L2-resident at the UNet's sizes) and writes y once.
// Reductions are two-stage with a fixed summation order => bitwise reproducible run to run.
#include "common.cuh"

namespace bd {
void count_launch(int n);

constexpr int kGnMaxSplits = 32;

static inline int gn_splits(int B, int HW) {
  int want = ceil_div(2 * num_sms(), B > 0 ? B : 1);
  int cap = HW / 64 > 0 ? HW / 64 : 1;  // >= 64 pixels per block
  int s = want < cap ? want : cap;
  if (s > kGnMaxSplits) s = kGnMaxSplits;
  return s < 1 ? 1 : s;
}

// block = C8 * rows threads (rows = max(1, 256 / C8)); thread owns channel-vector v = tid % C8
// smem: red[rows][C][2]
__global__ void gn_stats_kernel(const __half* __restrict__ x, int64_t ldx, float* __restrict__ work, int HW, int C,
                                int G, int splits) {
  extern __shared__ float red[];
  const int C8 = C / 8, rows = blockDim.x / C8;
  const int v = threadIdx.x % C8, r = threadIdx.x / C8;
  const int b = blockIdx.y, sp = blockIdx.x;
  const int p0 = (int)((int64_t)HW * sp / splits), p1 = (int)((int64_t)HW * (sp + 1) / splits);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0}, f[8];
  const __half* xb = x + (int64_t)b * HW * ldx + v * 8;
  for (int p = p0 + r; p < p1; p += rows) {
    unpack8(*reinterpret_cast<const half8*>(xb + (int64_t)p * ldx), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] += f[k] * f[k]; }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[(r * C + v * 8 + k) * 2 + 0] = s[k];
    red[(r * C + v * 8 + k) * 2 + 1] = q[k];
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int rr = 0; rr < rows; ++rr)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        ds += (double)red[(rr * C + c) * 2 + 0];
        dq += (double)red[(rr * C + c) * 2 + 1];
      }
    float* w = work + (((int64_t)b * splits + sp) * G + g) * 2;
    w[0] = (float)ds;
    w[1] = (float)dq;
  }
}

__device__ __forceinline__ void gn_finalize_stats(const float* __restrict__ work, float* sm_mean, float* sm_rstd,
                                                  int b, int G, int splits, int HW, int cpg, float eps) {
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int sp = 0; sp < splits; ++sp) {
      const float* w = work + (((int64_t)b * splits + sp) * G + g) * 2;
      ds += (double)w[0];
      dq += (double)w[1];
    }
    double n = (double)HW * cpg;
    double mean = ds / n;
    double var = dq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    sm_mean[g] = (float)mean;
    sm_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const __half* __restrict__ x, int64_t ldx,
                                                       __half* __restrict__ y, int64_t ldy,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ work, float* __restrict__ stats,
                                                       int HW, int C, int G, int splits, int asplits, float eps,
                                                       int apply_silu) {
  extern __shared__ float sm[];
  float* sm_mean = sm;
  float* sm_rstd = sm + G;
  const int b = blockIdx.y, cpg = C / G, C8 = C / 8;
  gn_finalize_stats(work, sm_mean, sm_rstd, b, G, splits, HW, cpg, eps);
  __syncthreads();
  if (blockIdx.x == 0 && stats)
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      stats[((int64_t)b * G + g) * 2 + 0] = sm_mean[g];
      stats[((int64_t)b * G + g) * 2 + 1] = sm_rstd[g];
    }
  const int p0 = (int)((int64_t)HW * blockIdx.x / asplits), p1 = (int)((int64_t)HW * (blockIdx.x + 1) / asplits);
  const int64_t total = (int64_t)(p1 - p0) * C8;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int v = (int)(i % C8);
    const int64_t p = p0 + i / C8;
    float f[8];
    unpack8(*reinterpret_cast<const half8*>(x + ((int64_t)b * HW + p) * ldx + v * 8), f);
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8), b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int g = (v * 8 + k) / cpg;
      float z = (f[k] - sm_mean[g]) * sm_rstd[g] * gm[k] + bt[k];
      f[k] = apply_silu ? silu_f(z) : z;
    }
    *reinterpret_cast<half8*>(y + ((int64_t)b * HW + p) * ldy + v * 8) = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// backward pass 1: per (b, split, c): s1 = sum dz, s2 = sum dz * xhat    (dz = dy * silu'(z))
// block = C8 * rows threads, smem red[rows][C][2] + mean/rstd [2G]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dsilu(float z) {
  float sg = sigmoid_f(z);
  return sg * (1.0f + z * (1.0f - sg));
}

__global__ void gn_bwd_reduce_kernel(const __half* __restrict__ x, int64_t ldx, const __half* __restrict__ dy,
                                     int64_t lddy, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ stats, float* __restrict__ work, int HW, int C, int G,
                                     int splits, int apply_silu) {
  extern __shared__ float red[];
  const int C8 = C / 8, rows = blockDim.x / C8, cpg = C / G;
  const int v = threadIdx.x % C8, r = threadIdx.x / C8;
  const int b = blockIdx.y, sp = blockIdx.x;
  const int p0 = (int)((int64_t)HW * sp / splits), p1 = (int)((int64_t)HW * (sp + 1) / splits);
  float mean[8], rstd[8], gm[8], bt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k, g = c / cpg;
    mean[k] = stats[((int64_t)b * G + g) * 2 + 0];
    rstd[k] = stats[((int64_t)b * G + g) * 2 + 1];
    gm[k] = gamma[c];
    bt[k] = beta[c];
  }
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0}, fx[8], fd[8];
  for (int p = p0 + r; p < p1; p += rows) {
    unpack8(*reinterpret_cast<const half8*>(x + ((int64_t)b * HW + p) * ldx + v * 8), fx);
    unpack8(*reinterpret_cast<const half8*>(dy + ((int64_t)b * HW + p) * lddy + v * 8), fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float xh = (fx[k] - mean[k]) * rstd[k];
      float dz = fd[k];
      if (apply_silu) dz *= dsilu(xh * gm[k] + bt[k]);
      s1[k] += dz;
      s2[k] += dz * xh;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[(r * C + v * 8 + k) * 2 + 0] = s1[k];
    red[(r * C + v * 8 + k) * 2 + 1] = s2[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int rr = 0; rr < rows; ++rr) {
      a += red[(rr * C + c) * 2 + 0];
      q += red[(rr * C + c) * 2 + 1];
    }
    float* w = work + (((int64_t)b * splits + sp) * 2) * C;
    w[c] = a;
    w[C + c] = q;
  }
}

// backward pass 2: dx = rstd * (dz*gamma - A/n - xhat * Bq/n) (+ add_dx)
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(
    const __half* __restrict__ x, int64_t ldx, const __half* __restrict__ dy, int64_t lddy,
    const __half* __restrict__ add, int64_t ldadd, __half* __restrict__ dx, int64_t lddx,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    const float* __restrict__ work, int HW, int C, int G, int splits, int asplits, int apply_silu) {
  extern __shared__ float sm[];
  float* gA = sm;       // [G] sum_c gamma*s1 / n
  float* gB = sm + G;   // [G] sum_c gamma*s2 / n
  const int b = blockIdx.y, cpg = C / G, C8 = C / 8;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      double s1 = 0.0, s2 = 0.0;
      for (int sp = 0; sp < splits; ++sp) {
        const float* w = work + (((int64_t)b * splits + sp) * 2) * C;
        s1 += (double)w[c];
        s2 += (double)w[C + c];
      }
      a += (double)gamma[c] * s1;
      q += (double)gamma[c] * s2;
    }
    double n = (double)HW * cpg;
    gA[g] = (float)(a / n);
    gB[g] = (float)(q / n);
  }
  __syncthreads();
  const int p0 = (int)((int64_t)HW * blockIdx.x / asplits), p1 = (int)((int64_t)HW * (blockIdx.x + 1) / asplits);
  const int64_t total = (int64_t)(p1 - p0) * C8;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int v = (int)(i % C8);
    const int64_t p = p0 + i / C8;
    const int64_t row = (int64_t)b * HW + p;
    float fx[8], fd[8], fa[8];
    unpack8(*reinterpret_cast<const half8*>(x + row * ldx + v * 8), fx);
    unpack8(*reinterpret_cast<const half8*>(dy + row * lddy + v * 8), fd);
    if (add) unpack8(*reinterpret_cast<const half8*>(add + row * ldadd + v * 8), fa);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = v * 8 + k, g = c / cpg;
      const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
      const float gm = gamma[c];
      float xh = (fx[k] - mean) * rstd;
      float dz = fd[k];
      if (apply_silu) dz *= dsilu(xh * gm + beta[c]);
      float r = rstd * (dz * gm - gA[g] - xh * gB[g]);
      if (add) r += fa[k];
      fx[k] = r;
    }
    *reinterpret_cast<half8*>(dx + row * lddx + v * 8) = pack8(fx);
  }
}

// dgamma[c] (+)= sum_{b,split} s2 ; dbeta[c] (+)= sum s1   (fixed order)
__global__ void gn_bwd_params_kernel(const float* __restrict__ work, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int BS, int C, int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < BS; ++i) {
    s1 += (double)work[(int64_t)i * 2 * C + c];
    s2 += (double)work[(int64_t)i * 2 * C + C + c];
  }
  dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s2;
  dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s1;
}

}  // namespace bd

using namespace bd;

extern "C" {

size_t bd_gn_workspace_floats(int B, int C) {
  // forward needs B*splits*G*2 (G <= C), backward B*splits*2*C
  return (size_t)(B > 0 ? B : 1) * kGnMaxSplits * 2 * (size_t)C;
}

int bd_groupnorm_fwd(const void* x, int64_t ld_x, void* y, int64_t ld_y, const float* gamma, const float* beta,
                     float* stats, float* work, int B, int HW, int C, int G, float eps, int apply_silu, void* stream) {
  BD_CHECK_ARG(x && y && gamma && beta && work, "bd_groupnorm_fwd: null pointer");
  BD_CHECK_ARG(C % 8 == 0 && C % G == 0 && ld_x % 8 == 0 && ld_y % 8 == 0 && C <= 2048,
               "bd_groupnorm_fwd: need C %% 8 == 0, C %% G == 0, ld %% 8 == 0, C <= 2048 (C=%d G=%d)", C, G);
  if (B == 0) return BD_OK;
  const int C8 = C / 8, rows = C8 >= 256 ? 1 : 256 / C8, threads = C8 * rows;
  const int splits = gn_splits(B, HW);
  gn_stats_kernel<<<dim3(splits, B), threads, (size_t)rows * C * 2 * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, work, HW, C, G, splits);
  gn_apply_kernel<<<dim3(splits, B), 256, 2 * G * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, (__half*)y, ld_y, gamma, beta, work, stats, HW, C, G, splits, splits, eps, apply_silu);
  count_launch(2);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_groupnorm_bwd(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, const void* add_dx, int64_t ld_add,
                     void* dx, int64_t ld_dx, const float* gamma, const float* beta, const float* stats, float* dgamma,
                     float* dbeta, float* work, int B, int HW, int C, int G, int apply_silu, void* stream) {
  BD_CHECK_ARG(x && dy && dx && gamma && beta && stats && dgamma && dbeta && work, "bd_groupnorm_bwd: null pointer");
  BD_CHECK_ARG(C % 8 == 0 && C % G == 0 && ld_x % 8 == 0 && ld_dy % 8 == 0 && ld_dx % 8 == 0 && C <= 2048 &&
                   (!add_dx || ld_add % 8 == 0),
               "bd_groupnorm_bwd: bad shape (C=%d G=%d)", C, G);
  if (B == 0) return BD_OK;
  const int C8 = C / 8, rows = C8 >= 256 ? 1 : 256 / C8, threads = C8 * rows;
  const int splits = gn_splits(B, HW);
  gn_bwd_reduce_kernel<<<dim3(splits, B), threads, (size_t)rows * C * 2 * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, (const __half*)dy, ld_dy, gamma, beta, stats, work, HW, C, G, splits, apply_silu);
  gn_bwd_apply_kernel<<<dim3(splits, B), 256, 2 * G * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, (const __half*)dy, ld_dy, (const __half*)add_dx, ld_add, (__half*)dx, ld_dx, gamma, beta,
      stats, work, HW, C, G, splits, splits, apply_silu);
  gn_bwd_params_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(work, dgamma, dbeta, B * splits, C, 1);
  count_launch(3);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

}  // extern "C"
