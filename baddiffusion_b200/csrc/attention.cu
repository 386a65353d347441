// Self-attention of diffusers' AttentionBlock (D/models/attention.py:135-162) on the fused-QKV buffer.
//  * generic CUDA-core path (any heads / head_dim / seq): used for the small mid-block maps (seq 16 / 64) and
//    multi-head configs whose head_dim is below the tensor-core K granularity;
//  * row softmax forward / backward used by the tcgen05 GEMM path (scores and probabilities live in HBM there).
// Softmax is computed in fp32 exactly like the reference (attention.py:161).
#include "common.cuh"

namespace bd {
void count_launch(int n);

constexpr int TQ = 16;  // rows of the "stationary" tile per block

// out[i][j] = sum_c A[i][c] * Bg[j][c]   (A: smem fp32 [TQ][d]; Bg: global fp16 rows, stride ldb)
__device__ __forceinline__ void tile_dots(const float* __restrict__ sA, int d, const __half* __restrict__ Bg,
                                          int64_t ldb, int S, float* __restrict__ sOut /*[TQ][S]*/, int rows) {
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    float acc[TQ];
#pragma unroll
    for (int i = 0; i < TQ; ++i) acc[i] = 0.f;
    const __half* br = Bg + (int64_t)j * ldb;
    for (int c = 0; c < d; c += 8) {
      float f[8];
      unpack8(*reinterpret_cast<const half8*>(br + c), f);
#pragma unroll
      for (int i = 0; i < TQ; ++i) {
        const float* a = sA + i * d + c;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[i] = fmaf(a[k], f[k], acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < TQ; ++i)
      if (i < rows) sOut[i * S + j] = acc[i];
  }
}

// out[i][c] = sum_j Wt[i][j] * Vg[j][c]; writes fp16 rows to Og (row stride ldo)
__device__ __forceinline__ void tile_weighted_sum(const float* __restrict__ sW /*[TQ][S]*/, int S,
                                                  const __half* __restrict__ Vg, int64_t ldv, int d,
                                                  __half* __restrict__ Og, int64_t ldo, int rows) {
  const int D8 = d / 8;
  const int groups = blockDim.x / D8 > 0 ? blockDim.x / D8 : 1;
  const int cv = threadIdx.x % D8, rg = threadIdx.x / D8;
  if (rg >= groups) return;
  for (int i0 = rg; i0 < rows; i0 += groups) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < S; ++j) {
      float f[8];
      unpack8(*reinterpret_cast<const half8*>(Vg + (int64_t)j * ldv + cv * 8), f);
      const float wv = sW[i0 * S + j];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(wv, f[k], acc[k]);
    }
    *reinterpret_cast<half8*>(Og + (int64_t)i0 * ldo + cv * 8) = pack8(acc);
  }
}

__device__ __forceinline__ void load_tile_f32(float* sA, const __half* g, int64_t ld, int rows, int d) {
  const int D8 = d / 8;
  for (int i = threadIdx.x; i < TQ * D8; i += blockDim.x) {
    int r = i / D8, v = i % D8;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < rows) unpack8(*reinterpret_cast<const half8*>(g + (int64_t)r * ld + v * 8), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) sA[r * d + v * 8 + k] = f[k];
  }
}

// grid (ceil(S/TQ), B*heads); smem: sQ[TQ][d] + sS[TQ][S]
__global__ void __launch_bounds__(256) attn_fwd_simt_kernel(const __half* __restrict__ qkv, int64_t ld,
                                                            __half* __restrict__ probs, __half* __restrict__ out,
                                                            int64_t ldo, int S, int C, int heads, float scale) {
  extern __shared__ float sm[];
  const int d = C / heads;
  float* sQ = sm;
  float* sS = sm + TQ * d;
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int i0 = blockIdx.x * TQ, rows = min(TQ, S - i0);
  const __half* base = qkv + (int64_t)b * S * ld + h * d;
  load_tile_f32(sQ, base + (int64_t)i0 * ld, ld, rows, d);
  __syncthreads();
  tile_dots(sQ, d, base + C, ld, S, sS, rows);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < rows; i += 8) {
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, sS[i * S + j] * scale);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      float e = expf(sS[i * S + j] * scale - mx);
      sS[i * S + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < S; j += 32) {
      // the reference casts probs back to the activation dtype (fp16 here) before P@V (attention.py:161)
      __half ph = __float2half_rn(sS[i * S + j] * inv);
      sS[i * S + j] = __half2float(ph);
      if (probs) probs[((int64_t)bh * S + i0 + i) * S + j] = ph;
    }
  }
  __syncthreads();
  tile_weighted_sum(sS, S, base + 2 * C, ld, d, out + ((int64_t)b * S + i0) * ldo + h * d, ldo, rows);
}

// backward A: per query tile: D_i = sum_j dP_ij P_ij ; dQ = (P*(dP-D)*scale) K
__global__ void __launch_bounds__(256) attn_bwd_q_simt_kernel(const __half* __restrict__ qkv, int64_t ld,
                                                              const __half* __restrict__ probs,
                                                              const __half* __restrict__ dout, int64_t lddo,
                                                              __half* __restrict__ dqkv, int64_t lddq,
                                                              float* __restrict__ Dvec, int S, int C, int heads,
                                                              float scale) {
  extern __shared__ float sm[];
  const int d = C / heads;
  float* sA = sm;            // dO tile [TQ][d]
  float* sS = sm + TQ * d;   // [TQ][S]
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int i0 = blockIdx.x * TQ, rows = min(TQ, S - i0);
  const __half* base = qkv + (int64_t)b * S * ld + h * d;
  load_tile_f32(sA, dout + ((int64_t)b * S + i0) * lddo + h * d, lddo, rows, d);
  __syncthreads();
  tile_dots(sA, d, base + 2 * C, ld, S, sS, rows);  // dP = dO V^T
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < rows; i += 8) {
    const __half* pr = probs + ((int64_t)bh * S + i0 + i) * S;
    float dsum = 0.f;
    for (int j = lane; j < S; j += 32) dsum += sS[i * S + j] * __half2float(pr[j]);
    dsum = warp_sum(dsum);
    if (lane == 0) Dvec[(int64_t)bh * S + i0 + i] = dsum;
    for (int j = lane; j < S; j += 32) sS[i * S + j] = __half2float(pr[j]) * (sS[i * S + j] - dsum) * scale;
  }
  __syncthreads();
  tile_weighted_sum(sS, S, base + C, ld, d, dqkv + ((int64_t)b * S + i0) * lddq + h * d, lddq, rows);  // dQ = dS K
}

// backward B: per key tile: dK = dS^T Q ; dV = P^T dO
__global__ void __launch_bounds__(256) attn_bwd_kv_simt_kernel(const __half* __restrict__ qkv, int64_t ld,
                                                               const __half* __restrict__ probs,
                                                               const __half* __restrict__ dout, int64_t lddo,
                                                               __half* __restrict__ dqkv, int64_t lddq,
                                                               const float* __restrict__ Dvec, int S, int C, int heads,
                                                               float scale) {
  extern __shared__ float sm[];
  const int d = C / heads;
  float* sA = sm;                    // V tile [TQ][d]
  float* sS = sm + TQ * d;           // dS^T [TQ][S]
  float* sP = sS + TQ * S;           // P^T  [TQ][S]
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int j0 = blockIdx.x * TQ, rows = min(TQ, S - j0);
  const __half* base = qkv + (int64_t)b * S * ld + h * d;
  const __half* dob = dout + (int64_t)b * S * lddo + h * d;
  load_tile_f32(sA, base + 2 * C + (int64_t)j0 * ld, ld, rows, d);
  __syncthreads();
  tile_dots(sA, d, dob, lddo, S, sS, rows);  // dP^T[j][i] = V_j . dO_i
  __syncthreads();
  for (int idx = threadIdx.x; idx < rows * S; idx += blockDim.x) {
    const int j = idx / S, i = idx % S;
    const float p = __half2float(probs[((int64_t)bh * S + i) * S + j0 + j]);
    sP[j * S + i] = p;
    sS[j * S + i] = p * (sS[j * S + i] - Dvec[(int64_t)bh * S + i]) * scale;
  }
  __syncthreads();
  tile_weighted_sum(sS, S, base, ld, d, dqkv + ((int64_t)b * S + j0) * lddq + C + h * d, lddq, rows);        // dK
  tile_weighted_sum(sP, S, dob, lddo, d, dqkv + ((int64_t)b * S + j0) * lddq + 2 * C + h * d, lddq, rows);   // dV
}

// ---------------------------------------------------------------------------------------------
// row softmax for the GEMM path: scores f32 (rows, S) -> probs f16; one warp per row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ scores, __half* __restrict__ probs,
                                                          int64_t rows, int S, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += nw) {
    const float* sr = scores + r * S;
    float v[32];  // S <= 1024
    float mx = -INFINITY;
    int n = 0;
    for (int j = lane; j < S; j += 32, ++n) { v[n] = sr[j] * scale; mx = fmaxf(mx, v[n]); }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = 0; k < n; ++k) { v[k] = expf(v[k] - mx); sum += v[k]; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    n = 0;
    for (int j = lane; j < S; j += 32, ++n) probs[r * S + j] = __float2half_rn(v[n] * inv);
  }
}
// dS = P * (dP - sum_j dP_j P_j) * scale  (dP f32, P f16 -> dS f16)
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ dP, const __half* __restrict__ P,
                                                          __half* __restrict__ dS, int64_t rows, int S, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += nw) {
    float p[32], g[32];
    float dsum = 0.f;
    int n = 0;
    for (int j = lane; j < S; j += 32, ++n) {
      p[n] = __half2float(P[r * S + j]);
      g[n] = dP[r * S + j];
      dsum += p[n] * g[n];
    }
    dsum = warp_sum(dsum);
    n = 0;
    for (int j = lane; j < S; j += 32, ++n) dS[r * S + j] = __float2half_rn(p[n] * (g[n] - dsum) * scale);
  }
}

int attn_fwd_simt(const void* qkv, int64_t ld_qkv, void* probs, void* out, int64_t ld_out, int B, int S, int C,
                  int heads, float scale, cudaStream_t st) {
  const int d = C / heads;
  size_t smem = ((size_t)TQ * d + (size_t)TQ * S) * sizeof(float);
  if (smem > 200 * 1024) return -1;
  if (smem > 48 * 1024) cudaFuncSetAttribute(attn_fwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  attn_fwd_simt_kernel<<<dim3(ceil_div(S, TQ), B * heads), 256, smem, st>>>((const __half*)qkv, ld_qkv, (__half*)probs,
                                                                           (__half*)out, ld_out, S, C, heads, scale);
  count_launch(1);
  return 0;
}

int attn_bwd_simt(const void* qkv, int64_t ld_qkv, const void* probs, const void* d_out, int64_t ld_dout, void* d_qkv,
                  int64_t ld_dqkv, void* work, int B, int S, int C, int heads, float scale, cudaStream_t st) {
  const int d = C / heads;
  size_t smem1 = ((size_t)TQ * d + (size_t)TQ * S) * sizeof(float);
  size_t smem2 = ((size_t)TQ * d + 2 * (size_t)TQ * S) * sizeof(float);
  if (smem2 > 200 * 1024) return -1;
  if (smem1 > 48 * 1024) cudaFuncSetAttribute(attn_bwd_q_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  if (smem2 > 48 * 1024) cudaFuncSetAttribute(attn_bwd_kv_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  dim3 grid(ceil_div(S, TQ), B * heads);
  attn_bwd_q_simt_kernel<<<grid, 256, smem1, st>>>((const __half*)qkv, ld_qkv, (const __half*)probs, (const __half*)d_out,
                                                   ld_dout, (__half*)d_qkv, ld_dqkv, (float*)work, S, C, heads, scale);
  attn_bwd_kv_simt_kernel<<<grid, 256, smem2, st>>>((const __half*)qkv, ld_qkv, (const __half*)probs, (const __half*)d_out,
                                                    ld_dout, (__half*)d_qkv, ld_dqkv, (const float*)work, S, C, heads, scale);
  count_launch(2);
  return 0;
}

int softmax_fwd_launch(const float* scores, void* probs, int64_t rows, int S, float scale, cudaStream_t st) {
  int grid = (int)((rows + 7) / 8);
  if (grid > 8 * num_sms()) grid = 8 * num_sms();
  softmax_fwd_kernel<<<grid, 256, 0, st>>>(scores, (__half*)probs, rows, S, scale);
  count_launch(1);
  return 0;
}
int softmax_bwd_launch(const float* dP, const void* P, void* dS, int64_t rows, int S, float scale, cudaStream_t st) {
  int grid = (int)((rows + 7) / 8);
  if (grid > 8 * num_sms()) grid = 8 * num_sms();
  softmax_bwd_kernel<<<grid, 256, 0, st>>>(dP, (const __half*)P, (__half*)dS, rows, S, scale);
  count_launch(1);
  return 0;
}

}  // namespace bd
