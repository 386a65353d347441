// GroupNorm (+SiLU) forward / backward on fp16 NHWC views with fp32/fp64 statistics (K4/K5).
// HBM-bound: forward reads x twice (the second read is L2-resident at the UNet's sizes) and writes y once.
// Statistics are reduced in two stages with a fixed summation order (bitwise reproducible); the parameter gradients
// dgamma / dbeta are accumulated across samples with fp32 atomics (one per sample and channel).
// Thread mapping (all kernels): block = C8 * rows threads, C8 = C/8; a thread owns ONE 8-channel vector
// (v = tid % C8) for its whole life, so per-channel parameters live in registers and every warp reads
// consecutive 16-byte vectors of a pixel row (coalesced); pixels are strided by `rows` with 4 loads in flight.
#include "common.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace bd {

// upper bound of the per-sample split count of the three-kernel path: 32 at training batch sizes, more when few large
// samples (CelebA-HQ: 4 samples of 256 x 256 per GPU) would otherwise leave most SMs without a block
static inline int gn_max_splits(int B) {
  int s = ceil_div(4 * num_sms(), B > 0 ? B : 1);
  return s < 32 ? 32 : (s > 256 ? 256 : s);
}
constexpr int UNR = 4;

// SiLU and its derivative from ONE special-function op.  These kernels are bound by the 16-lane MUFU pipe, not by HBM:
// exp + reciprocal is 2 MUFU ops per element (28 kcycles for a 128 x 32 x 32 x 128 tensor = the whole measured run
// time of the apply pass); tanh.approx.f32 is one.  Callers pass h = z/2 (the 1/2 is folded into the per-channel
// affine coefficients):  sigmoid(z) = (1 + tanh h)/2,  silu(z) = h + h*tanh h,
// silu'(z) = s*(1 + z*(1-s)) = s + s*h*(1 - tanh h).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float silu_h(float h) { return fmaf(h, tanh_approx(h), h); }
__device__ __forceinline__ float dsilu_h(float h) {
  const float t = tanh_approx(h);
  const float u = fmaf(-h, t, h);        // z*(1-s)
  const float sg = fmaf(0.5f, t, 0.5f);  // s
  return fmaf(sg, u, sg);
}

// Splits per sample: B * splits blocks must ALL be resident at once (`per_sm` blocks fit on an SM) -- these kernels are
// bandwidth bound, so an uneven spread over the SMs costs nothing, but a second, nearly empty wave costs a whole pass.
static inline int gn_splits(int B, int HW, int rows, int per_sm) {
  int want = per_sm * num_sms() / (B > 0 ? B : 1);
  if (want < 1) want = 1;
  int cap = HW / (rows * UNR) > 0 ? HW / (rows * UNR) : 1;
  int s = want < cap ? want : cap;
  if (s > gn_max_splits(B)) s = gn_max_splits(B);
  return s < 1 ? 1 : s;
}

// one block per sample: (mean, rstd) per group from the split partials, fixed order, double accumulation
// Runs in the CTA of gn_stats_kernel that finishes sample b last (partials come from other SMs: ld.global.cg).
__device__ __forceinline__ void gn_finalize_body(const float* work, float* __restrict__ stats, int G, int splits, int HW,
                                                 int cpg, float eps, int b) {
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    int sp = 0;
    for (; sp + 8 <= splits; sp += 8) {   // 8 independent loads in flight per sum (same fixed order of the adds)
      float2 w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = __ldcg(reinterpret_cast<const float2*>(work + (((int64_t)b * splits + sp + u) * G + g) * 2));
#pragma unroll
      for (int u = 0; u < 8; ++u) { ds += (double)w[u].x; dq += (double)w[u].y; }
    }
    for (; sp < splits; ++sp) {
      const float* w = work + (((int64_t)b * splits + sp) * G + g) * 2;
      ds += (double)__ldcg(w);
      dq += (double)__ldcg(w + 1);
    }
    const double n = (double)HW * cpg;
    const double mean = ds / n;
    double var = dq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((int64_t)b * G + g) * 2 + 0] = (float)mean;
    stats[((int64_t)b * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// smem: red[rows][C][2]
__global__ void __launch_bounds__(256, 4) gn_stats_kernel(const __half* __restrict__ x, int64_t ldx, float* __restrict__ work, int HW, int C,
                                int G, int splits, float* __restrict__ stats, int* __restrict__ counters, float eps) {
  extern __shared__ float red[];
  __shared__ int s_last;
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
  // last CTA of the sample: (mean, rstd) per group (used to be a separate one-warp-per-sample launch)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counters + b, 1) == splits - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  gn_finalize_body(work, stats, G, splits, HW, cpg, eps, b);
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
    if (apply_silu) { a[k] *= 0.5f; c[k] *= 0.5f; }   // h = z/2 for silu_h
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
          f[k] = apply_silu ? silu_h(z) : z;
        }
        *reinterpret_cast<half8*>(yb + (int64_t)(p + u * rows) * ldy) = pack8(f);
      }
  }
}


// GroupNorm (+SiLU) forward whose statistics were accumulated by the PRODUCER of x: the 3x3 conv that wrote x added the
// per-(sample, channel) sum and sum of squares of its fp16-rounded outputs to `sums` (B, ld_sums) = [channel][2] in its
// epilogue (umma_conv3.cu, Conv3Params::gn_sums), so this launch is a pure streaming pass -- 4 B/element, no reduction,
// no cluster barrier, any number of CTAs per sample.  Every thread derives (mean, rstd) of the groups its 8 channels
// belong to from the channel sums (<= 2 * C/G floats per group, L2-resident, fp64 combine); block (0, b) also writes
// stats (B, G, 2) for the backward.  Same affine-then-silu_h arithmetic as gn_apply_kernel.
constexpr int GN_APPLY_MAXG = 256;
__global__ void __launch_bounds__(256, 4) gn_apply_sums_kernel(const __half* __restrict__ x, int64_t ldx,
                                                               __half* __restrict__ y, int64_t ldy,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const float* __restrict__ sums, int64_t ld_sums,
                                                               float* __restrict__ stats, int HW, int C, int G, float eps,
                                                               int asplits, int apply_silu) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched with programmatic stream serialization
  const int b = blockIdx.y, cpg = C / G, C8 = C / 8, rows = blockDim.x / C8;
  const int v = threadIdx.x % C8, r = threadIdx.x / C8;
  const float* sb = sums + (int64_t)b * ld_sums;
  // (mean, rstd) of every group ONCE per CTA (G threads, fp64 combine), then broadcast through shared memory: doing the
  // fp64 divide / sqrt in every thread made this pass a flat 24 us whatever its size (ncu: xu 42 %, dram 17 %)
  __shared__ float sm_mean[GN_APPLY_MAXG], sm_rstd[GN_APPLY_MAXG];
  const double n = (double)HW * cpg;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float2 sq = *reinterpret_cast<const float2*>(sb + 2 * c);
      ds += (double)sq.x;
      dq += (double)sq.y;
    }
    const double mean = ds / n;
    double var = dq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float m = (float)mean, rs = (float)(1.0 / sqrt(var + (double)eps));
    sm_mean[g] = m;
    sm_rstd[g] = rs;
    if (blockIdx.x == 0 && stats) {
      stats[((int64_t)b * G + g) * 2 + 0] = m;
      stats[((int64_t)b * G + g) * 2 + 1] = rs;
    }
  }
  __syncthreads();
  float a[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k, g = ch / cpg;
    a[k] = sm_rstd[g] * gamma[ch];
    c[k] = beta[ch] - sm_mean[g] * a[k];
    if (apply_silu) { a[k] *= 0.5f; c[k] *= 0.5f; }
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
          f[k] = apply_silu ? silu_h(z) : z;
        }
        *reinterpret_cast<half8*>(yb + (int64_t)(p + u * rows) * ldy) = pack8(f);
      }
  }
}


// backward pass 1: per (b, split, c): s1 = sum dz, s2 = sum dz * xhat    (dz = dy * silu'(z))
// In the loop only  z = x*a + c  (a = rstd*gamma, c = beta - mean*a) is formed; sum dz*xhat is recovered from
// sum dz*x afterwards, so the per-thread state is 4 x 8 registers.
// per sample b: (1) channel sums over the split partials -> dgamma / dbeta contribution (fp32 atomics: one per
// (sample, channel)), (2) per-group gA = sum_c gamma*s1 / n, gB = sum_c gamma*s2 / n -> gab (B, G, 2).
// Runs in the last CTA of sample b of gn_bwd_reduce_kernel; the partials were written by other SMs, so they are read with
// ld.global.cg (L2).  cs: shared scratch, [2][C] floats.
__device__ __forceinline__ void gn_bwd_group_body(const float* work, const float* __restrict__ gamma,
                                                  float* __restrict__ gab, float* __restrict__ dgamma,
                                                  float* __restrict__ dbeta, float* __restrict__ dgb_parts, int HW, int C,
                                                  int G, int splits, int b, float* cs) {
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    // 8 loads per sum in flight: with few samples (CelebA-HQ: B = 4 -> 4 CTAs) the split count is ~100 and a single
    // dependent chain of L2 round trips made this tiny kernel 24 us; the summation order stays fixed
    float a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* w0 = work + ((int64_t)b * splits * 2) * C + c;
    int sp = 0;
    for (; sp + 8 <= splits; sp += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a1[u] += __ldcg(w0 + (int64_t)(sp + u) * 2 * C);
        a2[u] += __ldcg(w0 + (int64_t)(sp + u) * 2 * C + C);
      }
    }
    for (; sp < splits; ++sp) {
      a1[0] += __ldcg(w0 + (int64_t)sp * 2 * C);
      a2[0] += __ldcg(w0 + (int64_t)sp * 2 * C + C);
    }
    const float s1 = ((a1[0] + a1[1]) + (a1[2] + a1[3])) + ((a1[4] + a1[5]) + (a1[6] + a1[7]));
    const float s2 = ((a2[0] + a2[1]) + (a2[2] + a2[3])) + ((a2[4] + a2[5]) + (a2[6] + a2[7]));
    cs[c] = s1;
    cs[C + c] = s2;
    if (dgb_parts) {
      dgb_parts[(int64_t)b * 2 * C + c] = s1;
      dgb_parts[(int64_t)b * 2 * C + C + c] = s2;
    } else {
      atomicAdd(dbeta + c, s1);
      atomicAdd(dgamma + c, s2);
    }
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

__global__ void __launch_bounds__(256, 3) gn_bwd_reduce_kernel(
    const __half* __restrict__ x, int64_t ldx, const __half* __restrict__ dy, int64_t lddy,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    float* __restrict__ work, int HW, int C, int G, int splits, int apply_silu, float* __restrict__ gab,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dgb_parts, int* __restrict__ counters) {
  extern __shared__ float red[];
  __shared__ int s_last;
  const int C8 = C / 8, rows = blockDim.x / C8, cpg = C / G;
  const int v = threadIdx.x % C8, r = threadIdx.x / C8;
  const int b = blockIdx.y, sp = blockIdx.x;
  const int p0 = (int)((int64_t)HW * sp / splits), p1 = (int)((int64_t)HW * (sp + 1) / splits);
  float a[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k, g = ch / cpg;
    const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
    a[k] = 0.5f * rstd * gamma[ch];            // h = z/2 for dsilu_h
    c[k] = 0.5f * (beta[ch] - mean * rstd * gamma[ch]);
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
      if (apply_silu) dz *= dsilu_h(fmaf(fx[k], a[k], c[k]));
      s1[k] += dz;
      s2[k] = fmaf(dz, fx[k], s2[k]);
    }
    if (two) {
      unpack8(hx1, fx);
      unpack8(hd1, fd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float dz = fd[k];
        if (apply_silu) dz *= dsilu_h(fmaf(fx[k], a[k], c[k]));
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
  // The CTA that finishes a sample last also folds the split partials into the per-group terms (what used to be a
  // separate 4-CTA launch of ~20 us between the two passes at CelebA-HQ sizes).  counters[] is zeroed by the launcher.
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counters + b, 1) == splits - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  gn_bwd_group_body(work, gamma, gab, dgamma, dbeta, dgb_parts, HW, C, G, splits, b, red);
}

// backward pass 2:  dx = rstd*(dz*gamma - gA - xhat*gB) (+ add)  ==  k1*dz + c1*x + c0 (+ add)  with per-channel
// k1 = rstd*gamma, c1 = -rstd^2*gB, c0 = rstd*(mean*rstd*gB - gA);  z = x*k1 + cz for the SiLU derivative.
template <bool GSUM>
__global__ void __launch_bounds__(256, GSUM ? 2 : 3) gn_bwd_apply_kernel(
    const __half* __restrict__ x, int64_t ldx, const __half* __restrict__ dy, int64_t lddy,
    const __half* __restrict__ add, int64_t ldadd, const __half* __restrict__ add2, int64_t ldadd2,
    __half* __restrict__ dx, int64_t lddx,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    const float* __restrict__ gab, int HW, int C, int G, int asplits, int apply_silu, float* __restrict__ gsum,
    int64_t ld_gsum) {
  extern __shared__ float gsred[];   // [rows][C] when gsum is requested
  const int b = blockIdx.y, cpg = C / G, C8 = C / 8, rows = blockDim.x / C8;
  const int v = threadIdx.x % C8, r = threadIdx.x / C8;
  float k1[8], cz[8], c1[8], c0[8];
  float gs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // channel sums of this thread's (rounded) dx rows
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k, g = ch / cpg;
    const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
    const float gA = gab[((int64_t)b * G + g) * 2 + 0], gB = gab[((int64_t)b * G + g) * 2 + 1];
    k1[k] = rstd * gamma[ch];
    cz[k] = 0.5f * (beta[ch] - mean * k1[k]);  // h = z/2 = x*(k1/2) + cz for dsilu_h
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
      if (add2) {   // second fan-in operand (only with add)
        float fb[8];
        unpack8(*reinterpret_cast<const half8*>(add2 + row * ldadd2 + v * 8), fb);
#pragma unroll
        for (int k = 0; k < 8; ++k) fa[k] += fb[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float dz = fd[k];
        if (apply_silu) dz *= dsilu_h(fmaf(fx[k], 0.5f * k1[k], cz[k]));
        float o = fmaf(k1[k], dz, fmaf(c1[k], fx[k], c0[k]));
        if (add) o += fa[k];
        fx[k] = o;
      }
      const half8 hv = pack8(fx);
      *reinterpret_cast<half8*>(dx + row * lddx + v * 8) = hv;
      if (GSUM) {
        float fr[8];
        unpack8(hv, fr);
#pragma unroll
        for (int k = 0; k < 8; ++k) gs[k] += fr[k];
      }
    }
  }
  if (GSUM) {
    // per-sample channel sums of dx (the producer's bias gradient / the temb gradient) without a second pass over dx:
    // block-level fold through shared memory, one atomic per (CTA, channel) into the zeroed (B, C) view
#pragma unroll
    for (int k = 0; k < 8; ++k) gsred[r * C + v * 8 + k] = gs[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float t = 0.f;
      for (int rr = 0; rr < rows; ++rr) t += gsred[rr * C + c];
      atomicAdd(gsum + (int64_t)b * ld_gsum + c, t);
    }
  }
}


// =============================================================================================================
// Single-launch GroupNorm: ONE thread-block cluster per sample.  Each CTA of the cluster owns a pixel range; a thread
// keeps up to VMAX of its 16-byte vectors in registers between the statistics pass and the apply pass (what does
// not fit is re-read -- it was touched microseconds ago and is L2-resident), the per-CTA partial sums meet through
// distributed shared memory in a fixed order (bitwise reproducible), and every CTA normalises its own range.
// Against the three-kernel path above this is 1 launch instead of 3 and 4 B/element instead of 6 (forward),
// 6-8 instead of 10-12 (backward).
// =============================================================================================================
constexpr int GN_MAX_CS = 8;    // portable cluster limit; 16 (non-portable) measured 13 % slower per step
constexpr int GN_VMAX = 4;   // default register-cache depth (vectors per thread); 8 is also compiled (BD_GN_VMAX)

// Block-level reduction of per-thread, per-channel partials (thread (r, v) holds NV values for each of its 8 channels)
// to per-channel sums chs[j][C] (fp32), bank-conflict free in both directions:
//   red[((j*8 + k) * rows + r) * C8 + v]  -- consecutive lanes are consecutive v (and r): stride-1 stores;
//   the summing thread takes (j, k, v) with v fastest: stride-1 loads, `rows` (<= 64) serial adds in 4 chains.
template <int NV>
__device__ __forceinline__ void gn_block_channel_sums(float* red, float* chs, const float (*vals)[8], int rows, int C8,
                                                      int C, int r, int v) {
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) red[((j * 8 + k) * rows + r) * C8 + v] = vals[j][k];
  __syncthreads();
  for (int idx = threadIdx.x; idx < NV * C; idx += blockDim.x) {
    const int j = idx / C, rem = idx - j * C, k = rem / C8, vv = rem - k * C8;
    const float* src = red + (size_t)((j * 8 + k) * rows) * C8 + vv;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int rr = 0;
    for (; rr + 4 <= rows; rr += 4) {
      a0 += src[(rr + 0) * C8];
      a1 += src[(rr + 1) * C8];
      a2 += src[(rr + 2) * C8];
      a3 += src[(rr + 3) * C8];
    }
    for (; rr < rows; ++rr) a0 += src[rr * C8];
    chs[j * C + vv * 8 + k] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
}

// dynamic smem: red[threads][16] f32 | chs[2][C] f32 | grp[G][2] f64 | mr[G][2] f32
template <int VMAX>
__global__ void __launch_bounds__(512, 2) gn_fwd_fused_kernel(const __half* __restrict__ x, int64_t ldx,
                                                              __half* __restrict__ y, int64_t ldy,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float* __restrict__ stats_out,
                                                              int HW, int C, int G, float eps, int apply_silu) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched with programmatic stream serialization
  cg::cluster_group cl = cg::this_cluster();
  const int CS = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  extern __shared__ __align__(16) unsigned char gsm[];
  float* red = reinterpret_cast<float*>(gsm);
  float* chs = red + (size_t)blockDim.x * 16;
  double* grp = reinterpret_cast<double*>(chs + 2 * C);
  float* mr = reinterpret_cast<float*>(grp + 2 * G);
  const int C8 = C / 8, rows = blockDim.x / C8, cpg = C / G;
  const int tid = threadIdx.x, v = tid % C8, r = tid / C8;
  const int b = blockIdx.y;
  const int p0 = (int)((int64_t)HW * rank / CS), p1 = (int)((int64_t)HW * (rank + 1) / CS);
  const __half* xb = x + (int64_t)b * HW * ldx + v * 8;
  __half* yb = y + (int64_t)b * HW * ldy + v * 8;
  half8 hv[VMAX > 0 ? VMAX : 1];
  float sq[2][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) sq[0][k] = sq[1][k] = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int p = p0 + r + i * rows;
    if (p < p1) hv[i] = *reinterpret_cast<const half8*>(xb + (int64_t)p * ldx);
  }
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int p = p0 + r + i * rows;
    if (p < p1) {
      float f[8];
      unpack8(hv[i], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { sq[0][k] += f[k]; sq[1][k] = fmaf(f[k], f[k], sq[1][k]); }
    }
  }
  for (int p = p0 + r + VMAX * rows; p < p1; p += rows * UNR) {  // what the register cache does not hold: UNR loads in flight
    half8 t[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p + u * rows < p1) t[u] = *reinterpret_cast<const half8*>(xb + (int64_t)(p + u * rows) * ldx);
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p + u * rows < p1) {
        float f[8];
        unpack8(t[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { sq[0][k] += f[k]; sq[1][k] = fmaf(f[k], f[k], sq[1][k]); }
      }
  }
  gn_block_channel_sums<2>(red, chs, sq, rows, C8, C, r, v);
  // per-channel affine parameters: issued now so their latency hides behind the cluster exchange
  float ga[8], be[8];
  {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8), b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
  }
  for (int g = tid; g < G; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { ds += (double)chs[c]; dq += (double)chs[C + c]; }
    grp[2 * g] = ds;
    grp[2 * g + 1] = dq;
  }
  cl.sync();
  // gather every peer's group partials with ONE remote (DSMEM) load per thread, then reduce locally in rank order:
  // a single round trip instead of CS dependent ones
  double* rgrp = reinterpret_cast<double*>(red);   // red[] is free again: [CS][2G] doubles
  for (int i = tid; i < CS * 2 * G; i += blockDim.x) {
    const int rk = i / (2 * G), j = i - rk * 2 * G;
    rgrp[i] = cl.map_shared_rank(grp, rk)[j];
  }
  cl.barrier_arrive();  // this CTA no longer reads its peers' shared memory
  __syncthreads();
  for (int g = tid; g < G; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int rk = 0; rk < CS; ++rk) { ds += rgrp[rk * 2 * G + 2 * g]; dq += rgrp[rk * 2 * G + 2 * g + 1]; }
    const double n = (double)HW * cpg;
    const double mean = ds / n;
    double var = dq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float fm = (float)mean, fr = (float)(1.0 / sqrt(var + (double)eps));
    mr[2 * g] = fm;
    mr[2 * g + 1] = fr;
    if (rank == 0 && stats_out) {
      stats_out[((int64_t)b * G + g) * 2 + 0] = fm;
      stats_out[((int64_t)b * G + g) * 2 + 1] = fr;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int g = (v * 8 + k) / cpg;
    ga[k] *= mr[2 * g + 1];
    be[k] -= mr[2 * g] * ga[k];
    if (apply_silu) { ga[k] *= 0.5f; be[k] *= 0.5f; }   // h = z/2 for silu_h
  }
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int p = p0 + r + i * rows;
    if (p < p1) {
      float f[8];
      unpack8(hv[i], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float z = fmaf(f[k], ga[k], be[k]);
        f[k] = apply_silu ? silu_h(z) : z;
      }
      *reinterpret_cast<half8*>(yb + (int64_t)p * ldy) = pack8(f);
    }
  }
  for (int p = p0 + r + VMAX * rows; p < p1; p += rows * UNR) {   // second read: L2-resident (this CTA read it microseconds ago)
    half8 t[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p + u * rows < p1) t[u] = *reinterpret_cast<const half8*>(xb + (int64_t)(p + u * rows) * ldx);
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p + u * rows < p1) {
        float f[8];
        unpack8(t[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float z = fmaf(f[k], ga[k], be[k]);
          f[k] = apply_silu ? silu_h(z) : z;
        }
        *reinterpret_cast<half8*>(yb + (int64_t)(p + u * rows) * ldy) = pack8(f);
      }
  }
  cl.barrier_wait();  // peers may still be reading grp[]: do not retire before they are done
}

// Backward in one launch.  Phase 1: per-channel s1 = sum dz, s2 = sum dz*x over the CTA's pixels -> cluster-wide
// totals (every CTA sums all peers' partials, fixed order) -> dgamma/dbeta (atomics, rank 0) and the per-group
// gA, gB.  Phase 2: dx = k1*dz + c1*x + c0 (+ add).  Optionally the per-(sample, channel) sums of dx are produced
// (gsum): they are the bias gradients of the convolution that made x and, for norm2, the time_emb_proj gradient
// (D/models/resnet.py:574-580), which otherwise cost one extra pass over dx each.
// dynamic smem: red[threads][16] f32 | chs[2][C] f32 | chs2[C] f32 | tot[2][C] f32 | gab[G][2] f32
template <int VMAX, int MINB, bool ADD2 = false>   // ADD2: a second gradient fan-in operand (dx += add + add2), own instantiation
__global__ void __launch_bounds__(256, MINB) gn_bwd_fused_kernel(
    const __half* __restrict__ x, int64_t ldx, const __half* __restrict__ dy, int64_t lddy,
    const __half* __restrict__ add, int64_t ldadd, const __half* __restrict__ add2, int64_t ldadd2,
    __half* __restrict__ dx, int64_t lddx,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dgb_parts, float* __restrict__ gsum,
    int64_t ld_gsum, int HW, int C, int G, int apply_silu, int cache_dz) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched with programmatic stream serialization
  cg::cluster_group cl = cg::this_cluster();
  const int CS = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  extern __shared__ __align__(16) unsigned char gsm[];
  float* red = reinterpret_cast<float*>(gsm);
  float* chs = red + (size_t)blockDim.x * 16;
  float* chs2 = chs + 2 * C;
  float* tot = chs2 + C;
  float* gab = tot + 2 * C;
  // cache_dz: dz = dy * silu'(z) of the rows that do not fit the register cache is parked here (fp16, [row][thread]) by
  // phase 1, so phase 2 neither reads dy again nor evaluates silu' a second time for them
  half8* dzc = reinterpret_cast<half8*>(gsm + (((size_t)blockDim.x * 64 + (size_t)5 * C * 4 + (size_t)2 * G * 4 + 15) & ~(size_t)15));
  const int C8 = C / 8, rows = blockDim.x / C8, cpg = C / G;
  const int tid = threadIdx.x, v = tid % C8, r = tid / C8;
  const int b = blockIdx.y;
  const int p0 = (int)((int64_t)HW * rank / CS), p1 = (int)((int64_t)HW * (rank + 1) / CS);
  const int64_t rb = (int64_t)b * HW;
  const __half* xb = x + rb * ldx + v * 8;
  const __half* db = dy + rb * lddy + v * 8;
  half8 hx[VMAX > 0 ? VMAX : 1], hd[VMAX > 0 ? VMAX : 1];
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int p = p0 + r + i * rows;
    if (p < p1) {
      hx[i] = *reinterpret_cast<const half8*>(xb + (int64_t)p * ldx);
      hd[i] = *reinterpret_cast<const half8*>(db + (int64_t)p * lddy);
    }
  }
  float k1[8], hk[8], cz[8];   // k1 = rstd*gamma ; hk = k1/2 ; z/2 = x*hk + cz
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k, g = ch / cpg;
    k1[k] = stats[((int64_t)b * G + g) * 2 + 1] * gamma[ch];
    hk[k] = 0.5f * k1[k];
    cz[k] = 0.5f * (beta[ch] - stats[((int64_t)b * G + g) * 2 + 0] * k1[k]);   // h = z/2 = x*(k1/2) + cz
  }
  {
    float ss[2][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ss[0][k] = ss[1][k] = 0.f;
#pragma unroll
    for (int i = 0; i < VMAX; ++i) {
      const int p = p0 + r + i * rows;
      if (p < p1) {
        float fx[8], fd[8];
        unpack8(hx[i], fx);
        unpack8(hd[i], fd);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float dz = fd[k];
          if (apply_silu) dz *= dsilu_h(fmaf(fx[k], hk[k], cz[k]));
          fd[k] = dz;
          ss[0][k] += dz;
          ss[1][k] = fmaf(dz, fx[k], ss[1][k]);
        }
        hd[i] = pack8(fd);   // keep dz, not dy: phase 2 does not evaluate silu' again
      }
    }
    constexpr int TU = VMAX == 0 ? 4 : 2;   // 2*TU loads in flight
    int jt = 0;                             // index of the tail row (per thread)
    for (int p = p0 + r + VMAX * rows; p < p1; p += rows * TU, jt += TU) {
      half8 tx[TU], td[TU];
#pragma unroll
      for (int u = 0; u < TU; ++u)
        if (p + u * rows < p1) {
          tx[u] = *reinterpret_cast<const half8*>(xb + (int64_t)(p + u * rows) * ldx);
          td[u] = *reinterpret_cast<const half8*>(db + (int64_t)(p + u * rows) * lddy);
        }
#pragma unroll
      for (int u = 0; u < TU; ++u)
        if (p + u * rows < p1) {
          float fx[8], fd[8];
          unpack8(tx[u], fx);
          unpack8(td[u], fd);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float dz = fd[k];
            if (apply_silu) dz *= dsilu_h(fmaf(fx[k], hk[k], cz[k]));
            fd[k] = dz;
            ss[0][k] += dz;
            ss[1][k] = fmaf(dz, fx[k], ss[1][k]);
          }
          if (cache_dz) dzc[(size_t)(jt + u) * blockDim.x + tid] = pack8(fd);
        }
    }
    gn_block_channel_sums<2>(red, chs, ss, rows, C8, C, r, v);
  }
  cl.sync();
  for (int ch = tid; ch < C; ch += blockDim.x) {
    float ra[GN_MAX_CS], rx[GN_MAX_CS];   // remote loads first, adds after (one DSMEM round trip)
#pragma unroll
    for (int rk = 0; rk < GN_MAX_CS; ++rk)
      if (rk < CS) {
        const float* rc = cl.map_shared_rank(chs, rk);
        ra[rk] = rc[ch];
        rx[rk] = rc[C + ch];
      }
    float sa = 0.f, sx = 0.f;
#pragma unroll
    for (int rk = 0; rk < GN_MAX_CS; ++rk)
      if (rk < CS) { sa += ra[rk]; sx += rx[rk]; }
    // sum dz*xhat = rstd * (sum dz*x - mean * sum dz)
    const int g = ch / cpg;
    const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
    const float sq = rstd * (sx - mean * sa);
    tot[ch] = sa;
    tot[C + ch] = sq;
    if (rank == 0) {
      if (dgb_parts) {   // per-sample partials, summed over the batch by bd_bias_from_gsum (no atomics, fixed order)
        dgb_parts[(int64_t)b * 2 * C + ch] = sa;
        dgb_parts[(int64_t)b * 2 * C + C + ch] = sq;
      } else {
        atomicAdd(dbeta + ch, sa);
        atomicAdd(dgamma + ch, sq);
      }
    }
  }
  cl.barrier_arrive();
  __syncthreads();
  for (int g = tid; g < G; g += blockDim.x) {
    double ga = 0.0, gq = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      ga += (double)gamma[c] * (double)tot[c];
      gq += (double)gamma[c] * (double)tot[C + c];
    }
    const double n = (double)HW * cpg;
    gab[2 * g] = (float)(ga / n);
    gab[2 * g + 1] = (float)(gq / n);
  }
  __syncthreads();
  float c1[8], c0[8], so[1][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k, g = ch / cpg;
    const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
    const float gA = gab[2 * g], gB = gab[2 * g + 1];
    c1[k] = -rstd * rstd * gB;
    c0[k] = rstd * (mean * rstd * gB - gA);
    so[0][k] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int p = p0 + r + i * rows;
    if (p < p1) {
      const int64_t row = rb + p;
      float fx[8], fd[8], fa[8];
      unpack8(hx[i], fx);
      unpack8(hd[i], fd);
      if (add) unpack8(*reinterpret_cast<const half8*>(add + row * ldadd + v * 8), fa);
      if (ADD2) {
        float fb[8];
        unpack8(*reinterpret_cast<const half8*>(add2 + row * ldadd2 + v * 8), fb);
#pragma unroll
        for (int k = 0; k < 8; ++k) fa[k] += fb[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float dz = fd[k];
        float o = fmaf(k1[k], dz, fmaf(c1[k], fx[k], c0[k]));
        if (add) o += fa[k];
        fx[k] = o;
        so[0][k] += o;
      }
      *reinterpret_cast<half8*>(dx + row * lddx + v * 8) = pack8(fx);
    }
  }
  constexpr int TU2 = VMAX == 0 ? 3 : 2;
  int jt2 = 0;
  for (int p = p0 + r + VMAX * rows; p < p1; p += rows * TU2, jt2 += TU2) {
    half8 tx[TU2], td[TU2], ta[TU2], tb[ADD2 ? TU2 : 1];
#pragma unroll
    for (int u = 0; u < TU2; ++u)
      if (p + u * rows < p1) {
        const int64_t row = rb + p + u * rows;
        tx[u] = *reinterpret_cast<const half8*>(x + row * ldx + v * 8);
        td[u] = cache_dz ? dzc[(size_t)(jt2 + u) * blockDim.x + tid] : *reinterpret_cast<const half8*>(dy + row * lddy + v * 8);
        if (add) ta[u] = *reinterpret_cast<const half8*>(add + row * ldadd + v * 8);
        if (ADD2) tb[u] = *reinterpret_cast<const half8*>(add2 + row * ldadd2 + v * 8);
      }
#pragma unroll
    for (int u = 0; u < TU2; ++u)
      if (p + u * rows < p1) {
        const int64_t row = rb + p + u * rows;
        float fx[8], fd[8], fa[8];
        unpack8(tx[u], fx);
        unpack8(td[u], fd);
        if (add) unpack8(ta[u], fa);
        if (ADD2) {
          float fb[8];
          unpack8(tb[u], fb);
#pragma unroll
          for (int k = 0; k < 8; ++k) fa[k] += fb[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float dz = fd[k];
          if (apply_silu && !cache_dz) dz *= dsilu_h(fmaf(fx[k], hk[k], cz[k]));
          float o = fmaf(k1[k], dz, fmaf(c1[k], fx[k], c0[k]));
          if (add) o += fa[k];
          fx[k] = o;
          so[0][k] += o;
        }
        *reinterpret_cast<half8*>(dx + row * lddx + v * 8) = pack8(fx);
      }
  }
  cl.barrier_wait();  // every CTA has finished reading chs[]
  if (gsum) {         // uniform across the cluster
    gn_block_channel_sums<1>(red, chs2, so, rows, C8, C, r, v);  // red[] was last read before the cluster barrier above
    cl.sync();
    for (int ch = rank * blockDim.x + tid; ch < C; ch += CS * blockDim.x) {
      float ra[GN_MAX_CS];
#pragma unroll
      for (int rk = 0; rk < GN_MAX_CS; ++rk)
        if (rk < CS) ra[rk] = cl.map_shared_rank(chs2, rk)[ch];
      float sa = 0.f;
#pragma unroll
      for (int rk = 0; rk < GN_MAX_CS; ++rk)
        if (rk < CS) sa += ra[rk];
      gsum[(int64_t)b * ld_gsum + ch] = sa;
    }
    cl.sync();
  }
}


// ---------------------------------------------------------------------------------------------------------------
// GroupNorm backward, shared-memory resident.  Same cluster-per-sample decomposition and the same arithmetic as
// gn_bwd_fused_kernel, but the CTA's slice of x and dy is brought into shared memory by bulk async copies
// (cp.async.bulk, one elected warp, completion on an mbarrier) instead of through registers:
//   * every byte the CTA will ever read from x / dy is in flight from the first instruction, independent of the register
//     budget (the register kernel keeps 4 rows per thread and RE-READS the other 12 of 16 in its second phase);
//   * nothing is held in registers across the cluster exchange, so 3 CTAs of 256 threads fit an SM and one CTA's
//     arithmetic (MUFU-bound silu') and cluster round trip overlap the other CTAs' copies;
//   * dz = dy * silu'(z) is written back over dy in shared memory (fp16, exactly what the register kernel keeps), so the
//     second phase does not evaluate silu' again.
// Wide tensors are split at GROUP boundaries into `gridDim.z` independent channel chunks of Cs channels (groups do not
// interact), so a slice is at most ~64 KB: 256 -> 2 x 128, 384 -> 4 x 96 channels at 32 x 32.
// dynamic smem: sx[npx][Cs] f16 | sd[npx][Cs] f16 | red f32 (tot[2Cs], gab[2Gs] alias its head) | chs[2][Cs] | chs2[Cs] | mbarrier
__device__ __forceinline__ uint32_t gn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gn_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gn_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(gn_smem_u32(bar))
               : "memory");
}

// Block-wide per-channel sums of NV x 8 per-thread partials.  When a warp holds several rows of the same channel vectors
// (32 % C8 == 0) they are folded by shuffles first, so the scratch is [NV*8][warps][C8] instead of [NV*8][rows][C8].
template <int NV>
__device__ __forceinline__ void gn_block_channel_sums_w(float* red, float* chs, float (*vals)[8], int rows, int C8, int C,
                                                        int r, int v, bool fold) {
  if (fold) {
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k)
        for (int off = C8; off < 32; off <<= 1) vals[j][k] += __shfl_xor_sync(0xffffffffu, vals[j][k], off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane < C8) {
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) red[((j * 8 + k) * nw + warp) * C8 + v] = vals[j][k];
    }
    rows = nw;
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) red[((j * 8 + k) * rows + r) * C8 + v] = vals[j][k];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < NV * C; idx += blockDim.x) {
    const int j = idx / C, rem = idx - j * C, k = rem / C8, vv = rem - k * C8;
    const float* src = red + (size_t)((j * 8 + k) * rows) * C8 + vv;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int rr = 0;
    for (; rr + 4 <= rows; rr += 4) {
      a0 += src[(rr + 0) * C8];
      a1 += src[(rr + 1) * C8];
      a2 += src[(rr + 2) * C8];
      a3 += src[(rr + 3) * C8];
    }
    for (; rr < rows; ++rr) a0 += src[rr * C8];
    chs[j * C + vv * 8 + k] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
}

template <bool ADD2>
__global__ void __launch_bounds__(256, 3) gn_bwd_smem_kernel(
    const __half* __restrict__ x, int64_t ldx, const __half* __restrict__ dy, int64_t lddy,
    const __half* __restrict__ add, int64_t ldadd, const __half* __restrict__ add2, int64_t ldadd2,
    __half* __restrict__ dx, int64_t lddx,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dgb_parts, float* __restrict__ gsum,
    int64_t ld_gsum, int HW, int C, int G, int Cs, int npx_max, int red_floats, int fold, int apply_silu) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  cg::cluster_group cl = cg::this_cluster();
  const int CS = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  extern __shared__ __align__(128) unsigned char gsm[];
  __half* sx = reinterpret_cast<__half*>(gsm);
  __half* sd = sx + (size_t)npx_max * Cs;
  float* red = reinterpret_cast<float*>(sd + (size_t)npx_max * Cs);
  float* chs = red + red_floats;
  float* chs2 = chs + 2 * Cs;
  uint64_t* bar = reinterpret_cast<uint64_t*>(chs2 + Cs);
  const int cpg = C / G, Gs = Cs / cpg;
  float* tot = red;            // both dead before red is used again (gsum)
  float* gab = red + 2 * Cs;
  const int C8 = Cs / 8, rows = blockDim.x / C8;
  const int tid = threadIdx.x, v = tid % C8, r = tid / C8;
  const int b = blockIdx.y, c0 = blockIdx.z * Cs, g0 = c0 / cpg;
  const int p0 = (int)((int64_t)HW * rank / CS), p1 = (int)((int64_t)HW * (rank + 1) / CS), npx = p1 - p0;
  const int64_t rb = (int64_t)b * HW + p0;   // first row of this CTA's slice
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gn_smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched with programmatic stream serialization
  if (tid < 32) {
    const uint32_t row_bytes = (uint32_t)Cs * 2;
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gn_smem_u32(bar)), "r"(2u * npx * row_bytes) : "memory");
    __syncwarp();
    if (ldx == Cs && lddy == Cs) {   // the slice is one contiguous run
      if (tid == 0) gn_bulk_g2s(sx, x + rb * ldx, (uint32_t)npx * row_bytes, bar);
      if (tid == 1) gn_bulk_g2s(sd, dy + rb * lddy, (uint32_t)npx * row_bytes, bar);
    } else {
      for (int p = tid; p < npx; p += 32) {
        gn_bulk_g2s(sx + (size_t)p * Cs, x + (rb + p) * ldx + c0, row_bytes, bar);
        gn_bulk_g2s(sd + (size_t)p * Cs, dy + (rb + p) * lddy + c0, row_bytes, bar);
      }
    }
  }
  // the fan-in operands are read in phase 2: pull their lines towards L2 now
  if (add && (v & 7) == 0)
    for (int p = r; p < npx; p += rows) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(add + (rb + p) * ldadd + c0 + v * 8));
      if (ADD2) asm volatile("prefetch.global.L2 [%0];" ::"l"(add2 + (rb + p) * ldadd2 + c0 + v * 8));
    }
  float k1[8], cz[8];   // k1 = rstd*gamma ; z/2 = x*(k1/2) + cz
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = c0 + v * 8 + k, g = ch / cpg;
    k1[k] = stats[((int64_t)b * G + g) * 2 + 1] * gamma[ch];
    cz[k] = 0.5f * (beta[ch] - stats[((int64_t)b * G + g) * 2 + 0] * k1[k]);
  }
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(gn_smem_u32(bar)) : "memory");
  }
  {
    float ss[2][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ss[0][k] = ss[1][k] = 0.f;
#pragma unroll 2
    for (int p = r; p < npx; p += rows) {
      float fx[8], fd[8];
      unpack8(*reinterpret_cast<const half8*>(sx + (size_t)p * Cs + v * 8), fx);
      half8* dzp = reinterpret_cast<half8*>(sd + (size_t)p * Cs + v * 8);
      unpack8(*dzp, fd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float dz = fd[k];
        if (apply_silu) dz *= dsilu_h(fmaf(fx[k], 0.5f * k1[k], cz[k]));
        fd[k] = dz;
        ss[0][k] += dz;
        ss[1][k] = fmaf(dz, fx[k], ss[1][k]);
      }
      if (apply_silu) *dzp = pack8(fd);   // keep dz, not dy: phase 2 does not evaluate silu' again (own element: no hazard)
    }
    gn_block_channel_sums_w<2>(red, chs, ss, rows, C8, Cs, r, v, fold != 0);
  }
  cl.sync();
  for (int ch = tid; ch < Cs; ch += blockDim.x) {
    float ra[GN_MAX_CS], rx[GN_MAX_CS];   // remote loads first, adds after (one DSMEM round trip)
#pragma unroll
    for (int rk = 0; rk < GN_MAX_CS; ++rk)
      if (rk < CS) {
        const float* rc = cl.map_shared_rank(chs, rk);
        ra[rk] = rc[ch];
        rx[rk] = rc[Cs + ch];
      }
    float sa = 0.f, sxx = 0.f;
#pragma unroll
    for (int rk = 0; rk < GN_MAX_CS; ++rk)
      if (rk < CS) { sa += ra[rk]; sxx += rx[rk]; }
    const int g = (c0 + ch) / cpg;
    const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
    const float sq = rstd * (sxx - mean * sa);   // sum dz*xhat = rstd * (sum dz*x - mean * sum dz)
    tot[ch] = sa;
    tot[Cs + ch] = sq;
    if (rank == 0) {
      if (dgb_parts) {   // per-sample partials, summed over the batch by bd_bias_from_gsum (no atomics, fixed order)
        dgb_parts[(int64_t)b * 2 * C + c0 + ch] = sa;
        dgb_parts[(int64_t)b * 2 * C + C + c0 + ch] = sq;
      } else {
        atomicAdd(dbeta + c0 + ch, sa);
        atomicAdd(dgamma + c0 + ch, sq);
      }
    }
  }
  cl.barrier_arrive();
  __syncthreads();
  for (int g = tid; g < Gs; g += blockDim.x) {
    double ga = 0.0, gq = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      ga += (double)gamma[c0 + c] * (double)tot[c];
      gq += (double)gamma[c0 + c] * (double)tot[Cs + c];
    }
    const double n = (double)HW * cpg;
    gab[2 * g] = (float)(ga / n);
    gab[2 * g + 1] = (float)(gq / n);
  }
  __syncthreads();
  float c1[8], cc0[8], so[1][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int chl = v * 8 + k, g = (c0 + chl) / cpg;
    const float mean = stats[((int64_t)b * G + g) * 2 + 0], rstd = stats[((int64_t)b * G + g) * 2 + 1];
    const float gA = gab[2 * (g - g0)], gB = gab[2 * (g - g0) + 1];
    c1[k] = -rstd * rstd * gB;
    cc0[k] = rstd * (mean * rstd * gB - gA);
    so[0][k] = 0.f;
  }
  constexpr int TU = 2;
  for (int p = r; p < npx; p += rows * TU) {
    half8 ta[TU], tb[ADD2 ? TU : 1];
#pragma unroll
    for (int u = 0; u < TU; ++u)
      if (p + u * rows < npx) {
        const int64_t row = rb + p + u * rows;
        if (add) ta[u] = *reinterpret_cast<const half8*>(add + row * ldadd + c0 + v * 8);
        if (ADD2) tb[u] = *reinterpret_cast<const half8*>(add2 + row * ldadd2 + c0 + v * 8);
      }
#pragma unroll
    for (int u = 0; u < TU; ++u)
      if (p + u * rows < npx) {
        const int pp = p + u * rows;
        float fx[8], fd[8], fa[8];
        unpack8(*reinterpret_cast<const half8*>(sx + (size_t)pp * Cs + v * 8), fx);
        unpack8(*reinterpret_cast<const half8*>(sd + (size_t)pp * Cs + v * 8), fd);
        if (add) unpack8(ta[u], fa);
        if (ADD2) {
          float fb[8];
          unpack8(tb[u], fb);
#pragma unroll
          for (int k = 0; k < 8; ++k) fa[k] += fb[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float o = fmaf(k1[k], fd[k], fmaf(c1[k], fx[k], cc0[k]));
          if (add) o += fa[k];
          fx[k] = o;
          so[0][k] += o;
        }
        *reinterpret_cast<half8*>(dx + (rb + pp) * lddx + c0 + v * 8) = pack8(fx);
      }
  }
  cl.barrier_wait();  // every CTA has finished reading chs[]
  if (gsum) {         // uniform across the cluster
    __syncthreads();  // tot / gab (aliases of red) are dead in every thread
    gn_block_channel_sums_w<1>(red, chs2, so, rows, C8, Cs, r, v, fold != 0);
    cl.sync();
    for (int ch = rank * blockDim.x + tid; ch < Cs; ch += CS * blockDim.x) {
      float ra[GN_MAX_CS];
#pragma unroll
      for (int rk = 0; rk < GN_MAX_CS; ++rk)
        if (rk < CS) ra[rk] = cl.map_shared_rank(chs2, rk)[ch];
      float sa = 0.f;
#pragma unroll
      for (int rk = 0; rk < GN_MAX_CS; ++rk)
        if (rk < CS) sa += ra[rk];
      gsum[(int64_t)b * ld_gsum + c0 + ch] = sa;
    }
    cl.sync();
  }
}

}  // namespace bd

using namespace bd;

// Geometry of the cluster kernels: threads = C8 * rows (<= max_threads: 512 forward, 256 backward -- the backward kernel needs 128 registers), cluster size = the smallest of {1,2,4,8} that lets a
// thread keep its share of the sample in GN_VMAX register vectors (grown while the grid would not fill the GPU).
// Returns false when the sample is too large for one cluster to be a sensible unit (CelebA-HQ resolutions at small
// batch): the three-kernel path, which splits a sample over many independent CTAs, serves those.
static int gn_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static bool gn_fused_geometry(int B, int HW, int C, int max_threads, int vmax, int* threads, int* cs) {
  const int max_cs = GN_MAX_CS;
  if (getenv("BD_GN_V1")) return false;
  const int C8 = C / 8;
  if (C8 > max_threads || C8 < 1) return false;
  int rows = max_threads / C8;
  if (rows > HW) rows = HW;
  if (rows < 1) rows = 1;
  *threads = C8 * rows;
  int c = 1;
  while (c < max_cs && ceil_div(ceil_div(HW, c), rows) > vmax) c *= 2;
  while (c < 8 && (int64_t)B * c < num_sms() && HW / (2 * c) >= rows) c *= 2;
  if (ceil_div(ceil_div(HW, c), rows) > 32) return false;
  *cs = c;
  return true;
}


// Geometry of gn_bwd_smem_kernel: channel chunks (a divisor of G, chunk a multiple of 8 channels), 256-ish threads =
// C8 * rows, the smallest cluster whose slice (x + dy) plus scratch fits three CTAs per SM; grown while the grid is
// under two waves.  False when no such geometry exists (large images: the streaming kernels serve those).
struct GnBwdSmemGeo { int threads, cs, nchunk, Cs, npx_max, red_floats, fold; size_t smem; };
static bool gn_bwd_smem_geometry(int B, int HW, int C, int G, GnBwdSmemGeo* o) {
  if (gn_env_int("BD_GN_BWD_SMEM", 0) == 0 || getenv("BD_GN_V1")) return false;   // opt-in: measured slower
  const size_t budget = (size_t)gn_env_int("BD_GN_BWD_SMEM_KB", 74) * 1024;
  const int cpg = C / G;
  for (int nch = 1; nch <= G; nch *= 2) {
    if (G % nch) break;
    const int Cs = C / nch;
    if (Cs % 8 || Cs % cpg || Cs / 8 > 256) continue;
    const int C8 = Cs / 8;
    int rows = gn_env_int("BD_GN_BWD_SMEM_THREADS", 256) / C8;
    if (rows < 1) continue;
    if (rows > HW) rows = HW;
    const int threads = C8 * rows;
    const int fold = (C8 < 32 && 32 % C8 == 0 && threads % 32 == 0) ? 1 : 0;
    int red_floats = fold ? 16 * (threads / 32) * C8 : threads * 16;
    if (red_floats < 2 * Cs + 2 * (Cs / cpg)) red_floats = 2 * Cs + 2 * (Cs / cpg);
    auto bytes = [&](int c) {
      return (size_t)2 * ceil_div(HW, c) * Cs * 2 + (size_t)red_floats * 4 + (size_t)3 * Cs * 4 + 16;
    };
    int c = 1;
    while (c < GN_MAX_CS && (bytes(c) > budget || ceil_div(HW, c) < 1)) c *= 2;
    if (bytes(c) > budget) continue;
    while (c < GN_MAX_CS && (int64_t)B * c * nch < 2 * num_sms() && HW / (2 * c) >= rows) c *= 2;
    if (HW / c < 1) continue;
    o->threads = threads; o->cs = c; o->nchunk = nch; o->Cs = Cs; o->npx_max = ceil_div(HW, c);
    o->red_floats = red_floats; o->fold = fold; o->smem = bytes(c);
    return true;
  }
  return false;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, int threads, size_t smem, int cs, cudaStream_t st,
                                  Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // programmatic dependent launch: the kernels below execute griddepcontrol.wait before their first global access
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = getenv("BD_NO_PDL") ? 1 : 2;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

extern "C" {

size_t bd_gn_workspace_floats(int B, int C) {
  // forward needs B*splits*G*2 (G <= C), backward B*splits*2*C
  return (size_t)(B > 0 ? B : 1) * ((size_t)gn_max_splits(B) * 2 * (size_t)C + 2 * (size_t)C + 1);   // + arrival counters
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
  {
    int fthreads, cs;
    const int vmax = gn_env_int("BD_GN_VMAX", GN_VMAX) <= 4 ? 4 : 8;
    if (gn_fused_geometry(B, HW, C, gn_env_int("BD_GN_FT", 256), vmax, &fthreads, &cs) &&
        (size_t)fthreads * 64 >= (size_t)cs * 2 * G * sizeof(double)) {   // red[] doubles as the peer-partials gather buffer
      const size_t smem = (size_t)fthreads * 64 + (size_t)2 * C * 4 + (size_t)2 * G * 8 + (size_t)2 * G * 4;
      const bool stream_fwd = gn_env_int("BD_GN_FWD_STREAM", 0) != 0;
      cudaError_t e = launch_cluster(stream_fwd ? gn_fwd_fused_kernel<0> : vmax == 4 ? gn_fwd_fused_kernel<4> : gn_fwd_fused_kernel<8>, dim3(cs, B), fthreads, smem,
                                     cs, (cudaStream_t)stream, (const __half*)x, ld_x, (__half*)y, ld_y, gamma, beta, stats,
                                     HW, C, G, eps, apply_silu);
      if (e != cudaSuccess) { set_error("bd_groupnorm_fwd: cluster launch failed: %s", cudaGetErrorString(e)); return BD_ERR_CUDA; }
      count_launch(1);
      BD_CHECK_LAUNCH();
      return BD_OK;
    }
  }
  int threads, rows, splits, asplits;
  gn_geometry(B, HW, C, 4, &threads, &rows, &splits, &asplits);
  // stats may be omitted by inference callers: park them behind the partials
  float* st = stats ? stats : work + (size_t)B * splits * 2 * C;
  int* counters = reinterpret_cast<int*>(work + (size_t)B * ((size_t)gn_max_splits(B) * 2 * C + 2 * (size_t)C));
  cudaMemsetAsync(counters, 0, (size_t)B * sizeof(int), (cudaStream_t)stream);
  gn_stats_kernel<<<dim3(splits, B), threads, (size_t)rows * C * 2 * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, work, HW, C, G, splits, st, counters, eps);
  gn_apply_kernel<<<dim3(asplits, B), threads, 0, (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, (__half*)y, ld_y, gamma, beta, st, HW, C, G, asplits, apply_silu);
  count_launch(2);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_groupnorm_apply_sums(const void* x, int64_t ld_x, void* y, int64_t ld_y, const float* gamma, const float* beta,
                            const float* sums, int64_t ld_sums, float* stats, int B, int HW, int C, int G, float eps,
                            int apply_silu, void* stream) {
  BD_CHECK_ARG(x && y && gamma && beta && sums, "bd_groupnorm_apply_sums: null pointer");
  BD_CHECK_ARG(C % 8 == 0 && C % G == 0 && ld_x % 8 == 0 && ld_y % 8 == 0 && C <= 2048 && ld_sums >= 2 * (int64_t)C,
               "bd_groupnorm_apply_sums: need C %% 8 == 0, C %% G == 0, ld %% 8 == 0, C <= 2048, ld_sums >= 2C (C=%d G=%d)", C, G);
  BD_CHECK_ARG(G <= GN_APPLY_MAXG, "bd_groupnorm_apply_sums: at most %d groups", GN_APPLY_MAXG);
  if (B == 0) return BD_OK;
  int threads, rows, splits, asplits;
  gn_geometry(B, HW, C, 8, &threads, &rows, &splits, &asplits);
  {
    // programmatic dependent launch like the cluster kernels: the CTAs are scheduled while the producing conv drains
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(asplits, B);
    cfg.blockDim = dim3(threads);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = getenv("BD_NO_PDL") ? 0 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gn_apply_sums_kernel, (const __half*)x, ld_x, (__half*)y, ld_y, gamma, beta, sums,
                                       ld_sums, stats, HW, C, G, eps, asplits, apply_silu);
    if (e != cudaSuccess) { set_error("bd_groupnorm_apply_sums: launch failed: %s", cudaGetErrorString(e)); return BD_ERR_CUDA; }
  }
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_groupnorm_bwd(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, const void* add_dx, int64_t ld_add,
                     const void* add_dx2, int64_t ld_add2, void* dx, int64_t ld_dx, const float* gamma, const float* beta, const float* stats, float* dgamma,
                     float* dbeta, float* work, float* gsum, int64_t ld_gsum, float* dgb_parts, int B, int HW, int C, int G,
                     int apply_silu, void* stream) {
  BD_CHECK_ARG(x && dy && dx && gamma && beta && stats && work && (dgb_parts || (dgamma && dbeta)), "bd_groupnorm_bwd: null pointer");
  BD_CHECK_ARG(C % 8 == 0 && C % G == 0 && ld_x % 8 == 0 && ld_dy % 8 == 0 && ld_dx % 8 == 0 && C <= 2048 &&
                   (!add_dx || ld_add % 8 == 0) && (!add_dx2 || (add_dx && ld_add2 % 8 == 0)),
               "bd_groupnorm_bwd: bad shape (C=%d G=%d; add_dx2 needs add_dx)", C, G);
  if (B == 0) return BD_OK;
  {
    GnBwdSmemGeo geo;
    const bool aligned = !(((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)add_dx | (uintptr_t)add_dx2) & 15);
    if (aligned && HW >= gn_env_int("BD_GN_BWD_SMEM_MINHW", 64) && gn_bwd_smem_geometry(B, HW, C, G, &geo)) {
      auto kern = add_dx2 ? gn_bwd_smem_kernel<true> : gn_bwd_smem_kernel<false>;
      static bool attr_set[2] = {false, false};
      if (!attr_set[add_dx2 ? 1 : 0]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_set[add_dx2 ? 1 : 0] = true;
      }
      cudaError_t e = launch_cluster(kern, dim3(geo.cs, B, geo.nchunk), geo.threads, geo.smem, geo.cs, (cudaStream_t)stream,
                                     (const __half*)x, ld_x, (const __half*)dy, ld_dy, (const __half*)add_dx, ld_add,
                                     (const __half*)add_dx2, ld_add2, (__half*)dx, ld_dx, gamma, beta, stats, dgamma, dbeta,
                                     dgb_parts, gsum, ld_gsum, HW, C, G, geo.Cs, geo.npx_max, geo.red_floats, geo.fold, apply_silu);
      if (e != cudaSuccess) { set_error("bd_groupnorm_bwd: cluster launch failed: %s", cudaGetErrorString(e)); return BD_ERR_CUDA; }
      count_launch(1);
      BD_CHECK_LAUNCH();
      return BD_OK;
    }
  }
  {
    int fthreads, cs;
    const int vmax = gn_env_int("BD_GN_VMAX", GN_VMAX) <= 4 ? 4 : 8;
    // 128 threads (4 CTAs / SM at 128 registers) is the faster shape wherever it can hold a sample; wide tensors at
    // 32 x 32 (384 channels: 2 rows of 48 vectors) only fit the 32-rows-per-thread limit with 256
    if (gn_fused_geometry(B, HW, C, gn_env_int("BD_GN_BT", 128), vmax, &fthreads, &cs) ||
        (!getenv("BD_GN_BT") && gn_fused_geometry(B, HW, C, 256, vmax, &fthreads, &cs))) {
      size_t smem = (size_t)fthreads * 64 + (size_t)5 * C * 4 + (size_t)2 * G * 4;
      const bool stream_bwd = gn_env_int("BD_GN_BWD_STREAM", 0) != 0;
      // dz of the rows beyond the register cache parked in shared memory, while that leaves four CTAs per SM
      const int C8f = C / 8, frows = fthreads / C8f;
      const int tail_rows = ceil_div(ceil_div(HW, cs), frows) - (stream_bwd ? 0 : vmax);
      const size_t dz_bytes = tail_rows > 0 ? (size_t)(tail_rows + 3) * fthreads * 16 : 0;
      const int cache_dz = (tail_rows > 0 && gn_env_int("BD_GN_BWD_DZC", 1) && ((smem + 15) & ~(size_t)15) + dz_bytes <= 46 * 1024) ? 1 : 0;
      if (cache_dz) smem = ((smem + 15) & ~(size_t)15) + dz_bytes;
      // register budget: 128 / thread (2 blocks of 256); 80 or 64 registers spill 400-570 B per thread
      auto kern = add_dx2 ? (stream_bwd ? gn_bwd_fused_kernel<0, 3, true> : vmax == 8 ? gn_bwd_fused_kernel<8, 2, true> : gn_bwd_fused_kernel<4, 2, true>)
                          : (stream_bwd ? gn_bwd_fused_kernel<0, 3> : vmax == 8 ? gn_bwd_fused_kernel<8, 2> : gn_bwd_fused_kernel<4, 2>);
      cudaError_t e = launch_cluster(kern, dim3(cs, B), fthreads, smem,
                                     cs, (cudaStream_t)stream, (const __half*)x, ld_x, (const __half*)dy, ld_dy,
                                     (const __half*)add_dx, ld_add, (const __half*)add_dx2, ld_add2, (__half*)dx, ld_dx, gamma, beta,
                                     stats, dgamma, dbeta, dgb_parts, gsum, ld_gsum, HW, C, G, apply_silu, cache_dz);
      if (e != cudaSuccess) { set_error("bd_groupnorm_bwd: cluster launch failed: %s", cudaGetErrorString(e)); return BD_ERR_CUDA; }
      count_launch(1);
      BD_CHECK_LAUNCH();
      return BD_OK;
    }
  }
  int threads, rows, splits, asplits;
  gn_geometry(B, HW, C, 3, &threads, &rows, &splits, &asplits);
  // group sums live right behind the per-split partials in the workspace, the per-sample arrival counters at its end
  float* gab = work + (size_t)B * splits * 2 * C;
  int* counters = reinterpret_cast<int*>(work + (size_t)B * ((size_t)gn_max_splits(B) * 2 * C + 2 * (size_t)C));
  cudaMemsetAsync(counters, 0, (size_t)B * sizeof(int), (cudaStream_t)stream);
  gn_bwd_reduce_kernel<<<dim3(splits, B), threads, (size_t)rows * C * 2 * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)x, ld_x, (const __half*)dy, ld_dy, gamma, beta, stats, work, HW, C, G, splits, apply_silu, gab, dgamma, dbeta,
      dgb_parts, counters);
  if (gsum) cudaMemset2DAsync(gsum, (size_t)ld_gsum * sizeof(float), 0, (size_t)C * sizeof(float), B, (cudaStream_t)stream);
  if (gsum)
    gn_bwd_apply_kernel<true><<<dim3(asplits, B), threads, (size_t)rows * C * sizeof(float), (cudaStream_t)stream>>>(
        (const __half*)x, ld_x, (const __half*)dy, ld_dy, (const __half*)add_dx, ld_add, (const __half*)add_dx2, ld_add2,
        (__half*)dx, ld_dx, gamma, beta, stats, gab, HW, C, G, asplits, apply_silu, gsum, ld_gsum);
  else
    gn_bwd_apply_kernel<false><<<dim3(asplits, B), threads, 0, (cudaStream_t)stream>>>(
        (const __half*)x, ld_x, (const __half*)dy, ld_dy, (const __half*)add_dx, ld_add, (const __half*)add_dx2, ld_add2,
        (__half*)dx, ld_dx, gamma, beta, stats, gab, HW, C, G, asplits, apply_silu, nullptr, 0);
  count_launch(2);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

}  // extern "C"
