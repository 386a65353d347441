// GroupNorm (+SiLU) forward / backward on fp16 NHWC views with fp32/fp64 statistics (K4/K5).
// HBM-bound: forward reads x twice (the second read is L2-resident at the UNet's sizes) and writes y once.
// Statistics are reduced in two stages with a fixed summation order (bitwise reproducible); the parameter gradients
// dgamma / dbeta are accumulated across samples with fp32 atomics (one per sample and channel).
// Thread mapping (all kernels): block = C8 * rows threads, C8 = C/8; a thread owns ONE 8-channel vector
// (v = tid % C8) for its whole life, so per-channel parameters live in registers and every warp reads
// consecutive 16-byte vectors of a pixel row (coalesced); pixels are strided by `rows` with 4 loads in flight.
#include "common.cuh"

namespace bd {

constexpr int kGnMaxSplits = 32;
constexpr int UNR = 4;

// Splits per sample: B * splits blocks must ALL be resident at once (`per_sm` blocks fit on an SM) -- these kernels are
// bandwidth bound, so an uneven spread over the SMs costs nothing, but a second, nearly empty wave costs a whole pass.
static inline int gn_splits(int B, int HW, int rows, int per_sm) {
  int want = per_sm * num_sms() / (B > 0 ? B : 1);
  if (want < 1) want = 1;
  int cap = HW / (rows * UNR) > 0 ? HW / (rows * UNR) : 1;
  int s = want < cap ? want : cap;
  if (s > kGnMaxSplits) s = kGnMaxSplits;
  return s < 1 ? 1 : s;
}

// smem: red[rows][C][2]
__global__ void __launch_bounds__(256, 4) gn_stats_kernel(const __half* __restrict__ x, int64_t ldx, float* __restrict__ work, int HW, int C,
                                int G, int splits) {
  extern __shared__ float red[];
  const int C8 = C / 8, rows = blockDim.x / C8;
  const int v = threadIdx.x % C8, r = threadIdx.x / C8;
  const int b = blockIdx.y, sp = blockIdx.x;
  const int p0 = (int)((int64_t)HW * sp / splits), p1 = (int)((int64_t)HW * (sp + 1) / splits);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const __half* xb = x + (int64_t)b * HW * ldx + v * 8;
  for (int p = p0 + r; p < p1; p += rows * UNR) {
    half8 hv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p + u * rows < p1) hv[u] = *reinterpret_cast<const half8*>(xb + (int64_t)(p + u * rows) * ldx);
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p + u * rows < p1) {
        float f[8];
        unpack8(hv[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] += f[k] * f[k]; }
      }
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

// one block per sample: (mean, rstd) per group from the split partials, fixed order, double accumulation
__global__ void gn_finalize_kernel(const float* __restrict__ work, float* __restrict__ stats, int G, int splits, int HW,
                                   int cpg, float eps) {
  const int b = blockIdx.x;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int sp = 0; sp < splits; ++sp) {
      const float* w = work + (((int64_t)b * splits + sp) * G + g) * 2;
      ds += (double)w[0];
      dq += (double)w[1];
    }
    const double n = (double)HW * cpg;
    const double mean = ds / n;
    double var = dq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((int64_t)b * G + g) * 2 + 0] = (float)mean;
    stats[((int64_t)b * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// block = C8 * rows threads.  y = x * a + c  with a = rstd*gamma, c = beta - mean*a (per channel, in registers)
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(const __half* __restrict__ x, int64_t ldx,
                                                          __half* __restrict__ y, int64_t ldy,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ stats, int HW, int C, int G,
                                                          int asplits, int apply_silu) {
  const int b = blockIdx.y, cpg = C / G, C8 = C / 8, rows = blockDim.x / C8;
  const int v = threadIdx.x % C8, r = threadIdx.x / C8;
  float a[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k, g = ch / cpg;
    a[k] = stats[((int64_t)b * G + g) * 2 + 1] * gamma[ch];
    c[k] = beta[ch] - stats[((int64_t)b * G + g) * 2 + 0] * a[k];
  }
  const int p0 = (int)((int64_t)HW * blockIdx.x / asplits), p1 = (int)((int64_t)HW * (blockIdx.x + 1) / asplits);
  const __half* xb = x + (int64_t)b * HW * ldx + v * 8;
  __half* yb = y + (int64_t)b * HW * ldy + v * 8;
  for (int p = p0 + r; p < p1; p += rows * UNR) {
    half8 hv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p + u * rows < p1) hv[u] = *reinterpret_cast<const half8*>(xb + (int64_t)(p + u * rows) * ldx);
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p + u * rows < p1) {
        float f[8];
        unpack8(hv[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float z = fmaf(f[k], a[k], c[k]);
          f[k] = apply_silu ? silu_f(z) : z;
        }
        *reinterpret_cast<half8*>(yb + (int64_t)(p + u * rows) * ldy) = pack8(f);
      }
  }
}

__device__ __forceinline__ float dsilu(float z) {
  float sg = __fdividef(1.0f, 1.0f + __expf(-z));
  return sg * (1.0f + z * (1.0f - sg));
}

// backward pass 1: per (b, split, c): s1 = sum dz, s2 = sum dz * xhat    (dz = dy * silu'(z))
// In the loop only  z = x*a + c  (a = rstd*gamma, c = beta - mean*a) is formed; sum dz*xhat is recovered from
// sum dz*x afterwards, so the per-thread state is 4 x 8 registers.
__global__ void __launch_bounds__(256, 3) gn_bwd_reduce_kernel(
    const __half* __restrict__ x, int64_t ldx, const __half* __restrict__ dy, int64_t lddy,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    float* __restrict__ work, int HW, int C, int G, int splits, int apply_silu) {
  extern __shared__ float red[];
  const int C8 = C / 8, rows = blockDim.x / C8, cpg = C / G;
  const int v = threadIdx.x % C8, r = threadIdx.x / C8;
  const int b = blockIdx.y, sp = blockIdx.x;
  const int p0 = (int)((int64_t)HW * sp / splits), p1 = (int)((int64_t)HW * (sp + 1) / splits);
  float a[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k, g = ch / cpg;
    const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
    a[k] = rstd * gamma[ch];
    c[k] = beta[ch] - mean * a[k];
  }
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const __half* xb = x + (int64_t)b * HW * ldx + v * 8;
  const __half* db = dy + (int64_t)b * HW * lddy + v * 8;
  for (int p = p0 + r; p < p1; p += rows * 2) {
    const bool two = p + rows < p1;
    half8 hx0 = *reinterpret_cast<const half8*>(xb + (int64_t)p * ldx);
    half8 hd0 = *reinterpret_cast<const half8*>(db + (int64_t)p * lddy);
    half8 hx1 = hx0, hd1 = hd0;
    if (two) {
      hx1 = *reinterpret_cast<const half8*>(xb + (int64_t)(p + rows) * ldx);
      hd1 = *reinterpret_cast<const half8*>(db + (int64_t)(p + rows) * lddy);
    }
    float fx[8], fd[8];
    unpack8(hx0, fx);
    unpack8(hd0, fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float dz = fd[k];
      if (apply_silu) dz *= dsilu(fmaf(fx[k], a[k], c[k]));
      s1[k] += dz;
      s2[k] = fmaf(dz, fx[k], s2[k]);
    }
    if (two) {
      unpack8(hx1, fx);
      unpack8(hd1, fd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float dz = fd[k];
        if (apply_silu) dz *= dsilu(fmaf(fx[k], a[k], c[k]));
        s1[k] += dz;
        s2[k] = fmaf(dz, fx[k], s2[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    // sum dz*xhat = rstd * (sum dz*x - mean * sum dz)
    const int ch = v * 8 + k, g = ch / cpg;
    const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
    red[(r * C + ch) * 2 + 0] = s1[k];
    red[(r * C + ch) * 2 + 1] = rstd * (s2[k] - mean * s1[k]);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float sa = 0.f, sq = 0.f;
    for (int rr = 0; rr < rows; ++rr) {
      sa += red[(rr * C + ch) * 2 + 0];
      sq += red[(rr * C + ch) * 2 + 1];
    }
    float* w = work + (((int64_t)b * splits + sp) * 2) * C;
    w[ch] = sa;
    w[C + ch] = sq;
  }
}

// per sample b: (1) channel sums over the split partials -> dgamma / dbeta contribution (fp32 atomics: one per
// (sample, channel)), (2) per-group gA = sum_c gamma*s1 / n, gB = sum_c gamma*s2 / n -> gab (B, G, 2).
// smem: cs[2][C]
__global__ void __launch_bounds__(256) gn_bwd_group_kernel(const float* __restrict__ work, const float* __restrict__ gamma,
                                                           float* __restrict__ gab, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int HW, int C, int G, int splits) {
  extern __shared__ float cs[];
  const int b = blockIdx.x, cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int sp = 0; sp < splits; ++sp) {
      const float* w = work + (((int64_t)b * splits + sp) * 2) * C;
      s1 += w[c];
      s2 += w[C + c];
    }
    cs[c] = s1;
    cs[C + c] = s2;
    atomicAdd(dbeta + c, s1);
    atomicAdd(dgamma + c, s2);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += (double)gamma[c] * (double)cs[c];
      q += (double)gamma[c] * (double)cs[C + c];
    }
    const double n = (double)HW * cpg;
    gab[((int64_t)b * G + g) * 2 + 0] = (float)(a / n);
    gab[((int64_t)b * G + g) * 2 + 1] = (float)(q / n);
  }
}

// backward pass 2:  dx = rstd*(dz*gamma - gA - xhat*gB) (+ add)  ==  k1*dz + c1*x + c0 (+ add)  with per-channel
// k1 = rstd*gamma, c1 = -rstd^2*gB, c0 = rstd*(mean*rstd*gB - gA);  z = x*k1 + cz for the SiLU derivative.
__global__ void __launch_bounds__(256, 3) gn_bwd_apply_kernel(
    const __half* __restrict__ x, int64_t ldx, const __half* __restrict__ dy, int64_t lddy,
    const __half* __restrict__ add, int64_t ldadd, __half* __restrict__ dx, int64_t lddx,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    const float* __restrict__ gab, int HW, int C, int G, int asplits, int apply_silu) {
  const int b = blockIdx.y, cpg = C / G, C8 = C / 8, rows = blockDim.x / C8;
  const int v = threadIdx.x % C8, r = threadIdx.x / C8;
  float k1[8], cz[8], c1[8], c0[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k, g = ch / cpg;
    const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
    const float gA = gab[((int64_t)b * G + g) * 2 + 0], gB = gab[((int64_t)b * G + g) * 2 + 1];
    k1[k] = rstd * gamma[ch];
    cz[k] = beta[ch] - mean * k1[k];
    c1[k] = -rstd * rstd * gB;
    c0[k] = rstd * (mean * rstd * gB - gA);
  }
  const int p0 = (int)((int64_t)HW * blockIdx.x / asplits), p1 = (int)((int64_t)HW * (blockIdx.x + 1) / asplits);
  const int64_t rb = (int64_t)b * HW;
  for (int p = p0 + r; p < p1; p += rows * 2) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int pp = p + u * rows;
      if (pp >= p1) break;
      const int64_t row = rb + pp;
      float fx[8], fd[8], fa[8];
      unpack8(*reinterpret_cast<const half8*>(x + row * ldx + v * 8), fx);
      unpack8(*reinterpret_cast<const half8*>(dy + row * lddy + v * 8), fd);
      if (add) unpack8(*reinterpret_cast<const half8*>(add + row * ldadd + v * 8), fa);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float dz = fd[k];
        if (apply_silu) dz *= dsilu(fmaf(fx[k], k1[k], cz[k]));
        float o = fmaf(k1[k], dz, fmaf(c1[k], fx[k], c0[k]));
        if (add) o += fa[k];
        fx[k] = o;
      }
      *reinterpret_cast<half8*>(dx + row * lddx + v * 8) = pack8(fx);
    }
  }
}

}  // namespace bd

using namespace bd;

extern "C" {

size_t bd_gn_workspace_floats(int B, int C) {
  // forward needs B*splits*G*2 (G <= C), backward B*splits*2*C
  return (size_t)(B > 0 ? B : 1) * (kGnMaxSplits * 2 * (size_t)C + 2 * (size_t)C);
}

static inline void gn_geometry(int B, int HW, int C, int per_sm, int* threads, int* rows, int* splits, int* asplits) {
  const int C8 = C / 8;
  *rows = C8 >= 256 ? 1 : 256 / C8;
  *threads = C8 * (*rows);
  *splits = gn_splits(B, HW, *rows, per_sm);
  int want = per_sm * num_sms() / B;
  if (want < 1) want = 1;
  int cap = HW / (*rows) > 0 ? HW / (*rows) : 1;
  *asplits = want < cap ? want : cap;
  if (*asplits < 1) *asplits = 1;
}

int bd_groupnorm_fwd(const void* x, int64_t ld_x, void* y, int64_t ld_y, const float* gamma, const float* beta,
                     float* stats, float* work, int B, int HW, int C, int G, float eps, int apply_silu, void* stream) {
  BD_CHECK_ARG(x && y && gamma && beta && work, "bd_groupnorm_fwd: null pointer");
  BD_CHECK_ARG(C % 8 == 0 && C % G == 0 && ld_x % 8 == 0 && ld_y % 8 == 0 && C <= 2048,
               "bd_groupnorm_fwd: need C %% 8 == 0, C %% G == 0, ld %% 8 == 0, C <= 2048 (C=%d G=%d)", C, G);
  if (B == 0) return BD_OK;
  int threads, rows, splits, asplits;
  gn_geometry(B, HW, C, 4, &threads, &rows, &splits, &asplits);
  gn_stats_kernel<<<dim3(splits, B), threads, (size_t)rows * C * 2 * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, work, HW, C, G, splits);
  // stats may be omitted by inference callers: park them behind the partials
  float* st = stats ? stats : work + (size_t)B * splits * 2 * C;
  gn_finalize_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(work, st, G, splits, HW, C / G, eps);
  gn_apply_kernel<<<dim3(asplits, B), threads, 0, (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, (__half*)y, ld_y, gamma, beta, st, HW, C, G, asplits, apply_silu);
  count_launch(3);
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
  int threads, rows, splits, asplits;
  gn_geometry(B, HW, C, 3, &threads, &rows, &splits, &asplits);
  gn_bwd_reduce_kernel<<<dim3(splits, B), threads, (size_t)rows * C * 2 * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, (const __half*)dy, ld_dy, gamma, beta, stats, work, HW, C, G, splits, apply_silu);
  // group sums live right behind the per-split partials in the workspace
  float* gab = work + (size_t)B * splits * 2 * C;
  gn_bwd_group_kernel<<<B, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(work, gamma, gab, dgamma, dbeta, HW, C, G, splits);
  gn_bwd_apply_kernel<<<dim3(asplits, B), threads, 0, (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, (const __half*)dy, ld_dy, (const __half*)add_dx, ld_add, (__half*)dx, ld_dx, gamma, beta,
      stats, gab, HW, C, G, asplits, apply_silu);
  count_launch(3);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

}  // extern "C"
