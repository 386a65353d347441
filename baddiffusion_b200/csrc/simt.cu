// Generic CUDA-core kernels: implicit-GEMM convolution (fwd / dgrad / wgrad) for any geometry, the small fp32
// GEMM of the timestep path, the first/last convolutions (3-channel side in fp32 NCHW), column sums.
// These serve (a) the layers whose shapes do not fit the tcgen05 tiles (Cin=3, Cout=3, stride 2, channel
// counts that are not multiples of 64) and (b) the on-device cross-check of the tcgen05 kernels.
#include "common.cuh"

namespace bd {
void count_launch(int n);

struct ConvGeom {
  int B, OH, OW;      // output pixel grid of THIS launch (fwd: Ho,Wo; dgrad: H,W)
  int SH, SW;         // source pixel grid
  int Ck, Nout;       // reduction channels per tap, output channels
  int ks, stride, pad, transposed;
};

// source pixel for output (oh,ow) and tap (r,s); returns false when the tap reads padding / nothing
__device__ __forceinline__ bool src_pixel(const ConvGeom& g, int oh, int ow, int r, int s, int& sh, int& sw) {
  if (!g.transposed) {
    sh = oh * g.stride + r - g.pad;
    sw = ow * g.stride + s - g.pad;
  } else {
    int nh = oh + g.pad - r, nw = ow + g.pad - s;
    if (nh < 0 || nw < 0) return false;
    if (g.stride == 2) {
      if ((nh | nw) & 1) return false;
      sh = nh >> 1;
      sw = nw >> 1;
    } else {
      sh = nh;
      sw = nw;
    }
  }
  return sh >= 0 && sh < g.SH && sw >= 0 && sw < g.SW;
}

constexpr int TM = 64, TN = 64, TK = 16, LDS = 68;

struct ConvEpilogue {
  const float* bias;
  const float* bias2;
  const float* rowbias; int64_t ld_rowbias;
  const __half* residual; int64_t ld_res;
  float scale;
  void* y; int64_t ld_y; int out_f32;
};

__global__ void __launch_bounds__(256) simt_conv_kernel(ConvGeom g, const __half* __restrict__ x, int64_t ldx,
                                                        const __half* __restrict__ w, const __half* __restrict__ x2,
                                                        int64_t ldx2, int C2, const __half* __restrict__ w2,
                                                        ConvEpilogue ep) {
  __shared__ __align__(16) float As[TK][LDS];
  __shared__ __align__(16) float Bs[TK][LDS];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.B * g.OH * g.OW;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader roles
  const bool loadA = tid < 128;
  const int lrow = (tid & 127) >> 1, lvec = tid & 1;
  int ob = 0, oh = 0, ow = 0;
  bool rowvalid = false;
  if (loadA) {
    int64_t m = m0 + lrow;
    rowvalid = m < M;
    if (rowvalid) {
      ow = (int)(m % g.OW);
      int64_t q = m / g.OW;
      oh = (int)(q % g.OH);
      ob = (int)(q / g.OH);
    }
  }
  const int taps = g.ks * g.ks;
  const int nseg = taps + (x2 ? 1 : 0);
  for (int seg = 0; seg < nseg; ++seg) {
    const bool second = seg >= taps;
    const int Ck = second ? C2 : g.Ck;
    const __half* srcrow = nullptr;
    if (loadA && rowvalid) {
      if (second) {
        srcrow = x2 + (((int64_t)ob * g.OH + oh) * g.OW + ow) * ldx2;
      } else {
        int sh, sw;
        if (src_pixel(g, oh, ow, seg / g.ks, seg % g.ks, sh, sw))
          srcrow = x + (((int64_t)ob * g.SH + sh) * g.SW + sw) * ldx;
      }
    }
    const __half* wrow = nullptr;
    if (!loadA && !g.transposed) {
      int n = n0 + lrow;
      if (n < g.Nout) wrow = second ? (w2 + (int64_t)n * C2) : (w + ((int64_t)seg * g.Nout + n) * g.Ck);
    }
    for (int c0 = 0; c0 < Ck; c0 += TK) {
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (loadA || !g.transposed) {
        const int c = c0 + lvec * 8;
        const __half* p = loadA ? srcrow : wrow;
        if (p && c < Ck) unpack8(*reinterpret_cast<const half8*>(p + c), f);
        float(*dst)[LDS] = loadA ? As : Bs;
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[lvec * 8 + k][lrow] = f[k];
      } else {
        // dgrad: forward weights [tap][Ck = Cout_fwd][Nout = Cin_fwd]; row = reduction channel, n contiguous
        const int kk = (tid - 128) >> 3, nv = (tid - 128) & 7;
        const int c = c0 + kk, n = n0 + nv * 8;
        if (c < Ck && n < g.Nout) unpack8(*reinterpret_cast<const half8*>(w + ((int64_t)seg * g.Ck + c) * g.Nout + n), f);
        *reinterpret_cast<float4*>(&Bs[kk][nv * 8]) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(&Bs[kk][nv * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // epilogue
  const int HWo = g.OH * g.OW;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int b = (int)(m / HWo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.Nout) continue;
      float v = acc[i][j];
      if (ep.bias) v += ep.bias[n];
      if (ep.bias2) v += ep.bias2[n];
      if (ep.rowbias) v += ep.rowbias[(int64_t)b * ep.ld_rowbias + n];
      if (ep.residual) v += __half2float(ep.residual[m * ep.ld_res + n]);
      v *= ep.scale;
      if (ep.out_f32) reinterpret_cast<float*>(ep.y)[m * ep.ld_y + n] = v;
      else reinterpret_cast<__half*>(ep.y)[m * ep.ld_y + n] = __float2half_rn(v);
    }
  }
}

// wgrad: dW[tap][co][ci] += sum_p dY[p][co] * X[src(p,tap)][ci]; grid (mt*nt, taps, splits)
__global__ void __launch_bounds__(256) simt_wgrad_kernel(ConvGeom g, const __half* __restrict__ x, int64_t ldx,
                                                         const __half* __restrict__ dy, int64_t lddy,
                                                         float* __restrict__ dw, int Cin, int Cout, int splits) {
  __shared__ __align__(16) float As[TK][LDS];
  __shared__ __align__(16) float Bs[TK][LDS];
  const int tid = threadIdx.x;
  const int nt = ceil_div(Cin, TN);
  const int m0 = (blockIdx.x / nt) * TM, n0 = (blockIdx.x % nt) * TN;
  const int tap = blockIdx.y, r = tap / g.ks, s = tap % g.ks;
  const int64_t P = (int64_t)g.B * g.OH * g.OW;
  const int64_t chunk = ((P + splits - 1) / splits + TK - 1) / TK * TK;
  const int64_t pbeg = (int64_t)blockIdx.z * chunk, pend = pbeg + chunk < P ? pbeg + chunk : P;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool loadA = tid < 128;
  const int lk = (tid & 127) >> 3, lvec = tid & 7;  // pixel within chunk, 8-channel vector
  for (int64_t p0 = pbeg; p0 < pend; p0 += TK) {
    const int64_t p = p0 + lk;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (p < pend) {
      if (loadA) {
        const int c = m0 + lvec * 8;
        if (c < Cout) unpack8(*reinterpret_cast<const half8*>(dy + p * lddy + c), f);
      } else {
        const int c = n0 + lvec * 8;
        int ow = (int)(p % g.OW);
        int64_t q = p / g.OW;
        int oh = (int)(q % g.OH), ob = (int)(q / g.OH), sh, sw;
        if (c < Cin && src_pixel(g, oh, ow, r, s, sh, sw))
          unpack8(*reinterpret_cast<const half8*>(x + (((int64_t)ob * g.SH + sh) * g.SW + sw) * ldx + c), f);
      }
    }
    float(*dst)[LDS] = loadA ? As : Bs;
    *reinterpret_cast<float4*>(&dst[lk][lvec * 8]) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(&dst[lk][lvec * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = m0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = n0 + tx * 4 + j;
      if (ci >= Cin) continue;
      atomicAdd(dw + ((int64_t)tap * Cout + co) * Cin + ci, acc[i][j]);
    }
  }
}

// Batched bias gradients: job j = {gsum (B, ld) f32, ld, dst (C) f32, C}; dst[c] += sum_b gsum[b][c] in sample order.
// gsum rows are what the GroupNorm backward leaves behind (norm.cu), so every conv/linear bias gradient of the UNet is
// produced by ONE launch instead of one column-sum pass over its activation gradient each.
__global__ void __launch_bounds__(256) bias_from_gsum_kernel(const long long* __restrict__ jobs, int B) {
  const long long* j = jobs + 4 * (size_t)blockIdx.y;
  const float* g = reinterpret_cast<const float*>(j[0]);
  const int64_t ld = j[1];
  float* dst = reinterpret_cast<float*>(j[2]);
  const int C = (int)j[3];
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += g[(int64_t)b * ld + c];
  dst[c] += s;
}

// out[b][c] (+)= sum_{p in sample b} x[b,p,c]   (rows_per_b = HW; rows_per_b = all rows with B=1 for a bias grad)
// grid (C8-blocks, B, splits) -> atomics over splits
__global__ void __launch_bounds__(256) colsum_kernel(const __half* __restrict__ x, int64_t ldx, float* __restrict__ out,
                                                     int64_t ld_out, int64_t rows_per_b, int C, int splits) {
  __shared__ float red[8][256];
  const int C8 = C / 8;
  const int vpb = 32;  // 32 vectors (256 channels) per block in x
  const int v = blockIdx.x * vpb + (threadIdx.x & 31);
  const int r = threadIdx.x >> 5;  // 8 row groups
  const int b = blockIdx.y;
  const int64_t p0 = rows_per_b * blockIdx.z / splits, p1 = rows_per_b * (blockIdx.z + 1) / splits;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, f[8];
  if (v < C8)
    for (int64_t p = p0 + r; p < p1; p += 8) {
      unpack8(*reinterpret_cast<const half8*>(x + ((int64_t)b * rows_per_b + p) * ldx + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] += f[k];
    }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[r][(threadIdx.x & 31) * 8 + k] = s[k];
  __syncthreads();
  const int c = blockIdx.x * vpb * 8 + threadIdx.x;
  if (c < C) {
    float t = 0.f;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) t += red[rr][threadIdx.x];
    atomicAdd(out + (int64_t)b * ld_out + c, t);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void zero_f32_kernel(float* p, size_t n);

// small generic fp32 GEMM (strided): C[m,n] (+)= sum_k act(A[m,k]) * B[k,n] + bias[n]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                    const float* __restrict__ Bm, int64_t sbk, int64_t sbn,
                                                    float* __restrict__ C, int64_t scm, int64_t scn,
                                                    const float* __restrict__ bias, int M, int N, int K,
                                                    int accumulate, int act_silu_a, int ksplit) {
  __shared__ float As[TK][LDS];
  __shared__ float Bs[TK][LDS];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  // split-K: slice blockIdx.z of the reduction; partial results meet through atomics (C zeroed by the host)
  const int kper = ((K + ksplit - 1) / ksplit + TK - 1) / TK * TK;
  const int kbeg = blockIdx.z * kper, kend = min(K, kbeg + kper);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    for (int i = tid; i < TK * TM; i += 256) {
      int kk = i % TK, mm = i / TK;  // consecutive threads -> consecutive k (fast when sak == 1)
      int m = m0 + mm, k = k0 + kk;
      float v = (m < M && k < kend) ? A[(int64_t)m * sam + (int64_t)k * sak] : 0.f;
      if (act_silu_a == 1) v = silu_f(v);
      As[kk][mm] = v;
      int nn = i / TK, n = n0 + nn;
      float bv = (n < N && k < kend) ? Bm[(int64_t)k * sbk + (int64_t)n * sbn] : 0.f;
      if (act_silu_a == 2) bv = silu_f(bv);
      Bs[kk][nn] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[k][ty * 4 + i]; bv[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + ((bias && blockIdx.z == 0) ? bias[n] : 0.f);
      float* c = C + (int64_t)m * scm + (int64_t)n * scn;
      if (ksplit > 1) atomicAdd(c, v);
      else *c = accumulate ? (*c + v) : v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// timestep embedding + MLP (one block per sample), fp32
// ---------------------------------------------------------------------------------------------
constexpr int TEMB_SPB = 2;   // samples per block
constexpr int TEMB_THREADS = 1024;
__global__ void __launch_bounds__(TEMB_THREADS) temb_mlp_kernel(const int64_t* __restrict__ t, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ sin_out,
                                                       float* __restrict__ h1, float* __restrict__ emb,
                                                       __half* __restrict__ silu_emb, int B, int dim, int temb, int flip,
                                                       const float* __restrict__ freqs) {
  extern __shared__ float sm[];
  float* se = sm;                      // [SPB][dim]
  float* sa = sm + TEMB_SPB * dim;     // [SPB][temb]
  // L2 prefetch of W1 / W2 first, so it overlaps the sin/cos prologue (see the note at the layer loop)
  {
    const size_t n1 = (size_t)temb * dim * sizeof(float), n2 = (size_t)temb * temb * sizeof(float);
    const size_t first = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128, stride = (size_t)gridDim.x * blockDim.x * 128;
    for (size_t off = first; off < n1; off += stride)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(w1) + off));
    for (size_t off = first; off < n2; off += stride)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(w2) + off));
  }
  const int b0 = blockIdx.x * TEMB_SPB, half_dim = dim / 2;
  for (int i = threadIdx.x; i < TEMB_SPB * half_dim; i += blockDim.x) {
    // embeddings.py:41-52: emb = t * exp(-ln(10000) * i / (half - shift)); the frequency table is evaluated on the
    // host with the reference's own torch expression so the sin/cos arguments are bit-identical
    const int s = i / half_dim, f = i - s * half_dim, b = b0 + s < B ? b0 + s : B - 1;
    float arg = __fmul_rn((float)t[b], freqs[f]);
    float sv = sinf(arg), cv = cosf(arg);
    int is = flip ? half_dim + f : f, ic = flip ? f : half_dim + f;
    se[s * dim + is] = sv;
    se[s * dim + ic] = cv;
  }
  __syncthreads();
  if (sin_out)
    for (int i = threadIdx.x; i < TEMB_SPB * dim; i += blockDim.x)
      if (b0 + i / dim < B) sin_out[(int64_t)b0 * dim + i] = se[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // The kernel is a chain of dependent load round trips, not arithmetic (84 MFLOP): W1 / W2 are the fp32 master weights, which
  // Adam streamed out of L2 at the end of the previous step, so the first touch of every line is an HBM miss and a warp
  // that walks its rows group by group pays ~40 HBM latencies in a row (ncu: every stall sample sits on the first FFMA
  // after a load; 65 kcycles for one block's 1.25 MB).  Therefore: (a) the grid first PREFETCHES both matrices into L2
  // (each line requested once, by whichever block owns it -- one bulk of independent requests), (b) the block is
  // 32 warps so a warp walks only temb / (32 R) row groups per layer (20 dependent round trips in all instead of 40 with 8 warps), R = 4 rows at a time,
  // every lane fetching 16 bytes of each, (c) two samples share one pass over the weights, (d) block i starts its row
  // walk at a different rotation so the blocks do not queue on the same L2 slice, (e) after the butterfly every lane
  // holds every sum, so lane s*R+j finishes (bias, SiLU, store) output (s, j) instead of lane 0 doing all of them in a
  // row.  Per-(sample, row) accumulation order is fixed.
  constexpr int R = 4;
  const int groups = (temb + R - 1) / R, rot = (int)((blockIdx.x * 37u) % (unsigned)groups);
  auto layer = [&](const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ in, int K, auto&& store) {
    for (int gi = warp; gi < groups; gi += nw) {
      int g = gi + rot;
      if (g >= groups) g -= groups;
      const int n0 = g * R;
      float acc[TEMB_SPB][R];
#pragma unroll
      for (int s = 0; s < TEMB_SPB; ++s)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[s][j] = 0.f;
      for (int k = lane * 4; k < K; k += 128) {
        float4 wv[R];
#pragma unroll
        for (int j = 0; j < R; ++j)
          wv[j] = (n0 + j < temb) ? *reinterpret_cast<const float4*>(w + (int64_t)(n0 + j) * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < TEMB_SPB; ++s) {
          const float4 sv = *reinterpret_cast<const float4*>(in + s * K + k);
#pragma unroll
          for (int j = 0; j < R; ++j)
            acc[s][j] = fmaf(wv[j].w, sv.w, fmaf(wv[j].z, sv.z, fmaf(wv[j].y, sv.y, fmaf(wv[j].x, sv.x, acc[s][j]))));
        }
      }
#pragma unroll
      for (int s = 0; s < TEMB_SPB; ++s)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[s][j] = warp_sum(acc[s][j]);
      float mine = 0.f;
#pragma unroll
      for (int s = 0; s < TEMB_SPB; ++s)
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (lane == s * R + j) mine = acc[s][j];
      if (lane < TEMB_SPB * R) {
        const int s = lane / R, n = n0 + lane % R;
        if (n < temb && b0 + s < B) store(s, n, mine + bias[n]);
      }
    }
  };
  layer(w1, b1, se, dim, [&](int s, int n, float h) {
    if (h1) h1[(int64_t)(b0 + s) * temb + n] = h;
    sa[s * temb + n] = silu_f(h);
  });
  __syncthreads();
  layer(w2, b2, sa, temb, [&](int s, int n, float e) {
    emb[(int64_t)(b0 + s) * temb + n] = e;
    if (silu_emb) silu_emb[(int64_t)(b0 + s) * temb + n] = __float2half_rn(silu_f(e));
  });
}

// ---------------------------------------------------------------------------------------------
// conv_in: f32 NCHW (Cin small) -> f16 NHWC; weights OIHW f32 cached in smem as [ci*9+tap][Cout]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_in_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, __half* __restrict__ y,
                                                          int64_t ldy, int B, int Cin, int H, int W, int Cout) {
  extern __shared__ float sw[];  // [Cin*9][Cout] + bias[Cout]
  const int K = Cin * 9;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {  // w packed [tap][Cout][Cin]
    int ci = i % Cin, co = (i / Cin) % Cout, tap = i / (Cin * Cout);
    sw[(ci * 9 + tap) * Cout + co] = w[i];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[K * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int C8 = Cout / 8;
  const int64_t total = (int64_t)B * H * W * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % C8);
    const int64_t p = i / C8;
    const int ww = (int)(p % W);
    const int64_t q = p / W;
    const int hh = (int)(q % H), b = (int)(q / H);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = sw[K * Cout + v * 8 + k];
    for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          int sh = hh + r - 1, sx = ww + s - 1;
          if (sh < 0 || sh >= H || sx < 0 || sx >= W) continue;
          float xv = x[(((int64_t)b * Cin + ci) * H + sh) * W + sx];
          const float* wr = sw + (ci * 9 + r * 3 + s) * Cout + v * 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(xv, wr[k], acc[k]);
        }
    *reinterpret_cast<half8*>(y + p * ldy + v * 8) = pack8(acc);
  }
}

// conv_in wgrad: dW[tap][co][ci] += sum_p dY[p][co] x[ci][src(p,tap)].  Pixels are processed in chunks of 64 whose
// Cin*9 (<= 36) input taps are staged in shared memory once; thread = (co, pixel lane) reads dY coalesced and the taps
// as shared-memory broadcasts.  Block partials meet in shared memory, then one global atomic per output and block.
template <int KMAX>
__global__ void __launch_bounds__(256) conv_in_wgrad_kernel(const float* __restrict__ x, const __half* __restrict__ dy,
                                                            int64_t lddy, float* __restrict__ dw,
                                                            float* __restrict__ dbias, int B, int Cin, int H, int W,
                                                            int Cout) {
  extern __shared__ float sacc[];  // [Cout][K+1] | patches [64][KMAX]
  const int K = Cin * 9;
  float* sx = sacc + Cout * (K + 1);
  for (int i = threadIdx.x; i < Cout * (K + 1); i += blockDim.x) sacc[i] = 0.f;
  const int64_t P = (int64_t)B * H * W;
  const int64_t chunk = (P + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * chunk, p1 = p0 + chunk < P ? p0 + chunk : P;
  const int lanes_p = blockDim.x / Cout > 0 ? blockDim.x / Cout : 1;
  const int co = threadIdx.x % Cout, pl = threadIdx.x / Cout;
  const bool active = pl < lanes_p;
  float acc[KMAX], bacc = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
  for (int64_t pc = p0; pc < p1; pc += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * K; i += blockDim.x) {
      const int pix = i / K, k = i % K;
      const int64_t p = pc + pix;
      float xv = 0.f;
      if (p < p1) {
        const int ww = (int)(p % W);
        const int64_t q = p / W;
        const int hh = (int)(q % H), b = (int)(q / H);
        const int ci = k / 9, tap = k % 9;
        const int sh = hh + tap / 3 - 1, sxx = ww + tap % 3 - 1;
        if (sh >= 0 && sh < H && sxx >= 0 && sxx < W) xv = x[(((int64_t)b * Cin + ci) * H + sh) * W + sxx];
      }
      sx[pix * KMAX + k] = xv;
    }
    __syncthreads();
    if (active) {
      for (int pix = pl; pix < 64 && pc + pix < p1; pix += lanes_p) {
        const float d = __half2float(dy[(pc + pix) * lddy + co]);
        bacc += d;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) acc[k] = fmaf(d, sx[pix * KMAX + k], acc[k]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) atomicAdd(&sacc[co * (K + 1) + k], acc[k]);
    atomicAdd(&sacc[co * (K + 1) + K], bacc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * (K + 1); i += blockDim.x) {
    const int c = i / (K + 1), k = i % (K + 1);
    if (k < K) atomicAdd(dw + ((int64_t)(k % 9) * Cout + c) * Cin + k / 9, sacc[i]);
    else if (dbias) atomicAdd(dbias + c, sacc[i]);
  }
}

// conv_out fwd: f16 NHWC (Cin) -> f32 NCHW (Cout <= 4); warp per pixel, lanes over channels
__global__ void __launch_bounds__(256) conv_out_fwd_kernel(const __half* __restrict__ x, int64_t ldx,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ y, int B, int Cin, int H, int W,
                                                           int Cout) {
  extern __shared__ float sw[];  // packed [tap][co][Cin]
  for (int i = threadIdx.x; i < Cout * Cin * 9; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t P = (int64_t)B * H * W;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < P; p += nwarps) {
    const int ww = (int)(p % W);
    const int64_t q = p / W;
    const int hh = (int)(q % H), b = (int)(q / H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tap = 0; tap < 9; ++tap) {
      int sh = hh + tap / 3 - 1, sx = ww + tap % 3 - 1;
      if (sh < 0 || sh >= H || sx < 0 || sx >= W) continue;
      const __half* xr = x + (((int64_t)b * H + sh) * W + sx) * ldx;
      for (int c = lane * 4; c < Cin; c += 128) {
        const __half2* hp = reinterpret_cast<const __half2*>(xr + c);
        float2 f0 = __half22float2(hp[0]), f1 = __half22float2(hp[1]);
        for (int co = 0; co < Cout; ++co) {
          const float* wr = sw + (tap * Cout + co) * Cin + c;
          acc[co] = fmaf(f0.x, wr[0], fmaf(f0.y, wr[1], fmaf(f1.x, wr[2], fmaf(f1.y, wr[3], acc[co]))));
        }
      }
    }
    for (int co = 0; co < Cout; ++co) {
      float v = warp_sum(acc[co]);
      if (lane == 0) y[(((int64_t)b * Cout + co) * H + hh) * W + ww] = v + (bias ? bias[co] : 0.f);
    }
  }
}

// conv_out dgrad: dX[q][ci] = sum_{tap,co} dY[co][q + 1 - tap] W[co][ci][tap]; thread per (q, 8 channels)
__global__ void __launch_bounds__(256) conv_out_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ dy,
                                                             __half* __restrict__ dx, int64_t lddx, int B, int Cin,
                                                             int H, int W, int Cout) {
  extern __shared__ float sw[];  // packed [tap*Cout+co][Cin]
  for (int i = threadIdx.x; i < Cout * Cin * 9; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int C8 = Cin / 8;
  const int64_t total = (int64_t)B * H * W * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % C8);
    const int64_t p = i / C8;
    const int ww = (int)(p % W);
    const int64_t q = p / W;
    const int hh = (int)(q % H), b = (int)(q / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int tap = 0; tap < 9; ++tap) {
      int oh = hh + 1 - tap / 3, ow = ww + 1 - tap % 3;
      if (oh < 0 || oh >= H || ow < 0 || ow >= W) continue;
      for (int co = 0; co < Cout; ++co) {
        float d = dy[(((int64_t)b * Cout + co) * H + oh) * W + ow];
        const float* wr = sw + (tap * Cout + co) * Cin + v * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(d, wr[k], acc[k]);
      }
    }
    *reinterpret_cast<half8*>(dx + p * lddx + v * 8) = pack8(acc);
  }
}

// conv_out wgrad: dW[tap][co][ci] += sum_p dY[co][p] X[src(p,tap)][ci].  thread = (4-channel quad, pixel lane): a
// warp reads whole pixel rows (coalesced); Cout*9 (<= 36) x 4 accumulators per thread; shared-memory then global atomics.
template <int KMAX>
__global__ void __launch_bounds__(256) conv_out_wgrad_kernel(const __half* __restrict__ x, int64_t ldx,
                                                             const float* __restrict__ dy, float* __restrict__ dw,
                                                             float* __restrict__ dbias, int B, int Cin, int H, int W,
                                                             int Cout) {
  extern __shared__ float sacc[];  // [9*Cout][Cin] + bias[4]
  const int nout = 9 * Cout * Cin;
  for (int i = threadIdx.x; i < nout + 4; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int64_t P = (int64_t)B * H * W;
  const int64_t chunk = (P + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * chunk, p1 = p0 + chunk < P ? p0 + chunk : P;
  const int nq = Cin / 4, lanes_p = blockDim.x / nq;
  const int qd = threadIdx.x % nq, pl = threadIdx.x / nq;
  if (pl < lanes_p) {
    float acc[KMAX][4];
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < KMAX; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
    for (int64_t p = p0 + pl; p < p1; p += lanes_p) {
      const int ww = (int)(p % W);
      const int64_t q = p / W;
      const int hh = (int)(q % H), b = (int)(q / H);
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int co = 0; co < 4; ++co)
        if (co < Cout) d[co] = __ldg(dy + (((int64_t)b * Cout + co) * H + hh) * W + ww);
      if (qd == 0) {
#pragma unroll
        for (int co = 0; co < 4; ++co) bacc[co] += d[co];
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int sh = hh + tap / 3 - 1, sx = ww + tap % 3 - 1;
        if (sh < 0 || sh >= H || sx < 0 || sx >= W) continue;
        const __half2* hp = reinterpret_cast<const __half2*>(x + (((int64_t)b * H + sh) * W + sx) * ldx + qd * 4);
        const float2 f0 = __half22float2(hp[0]), f1 = __half22float2(hp[1]);
#pragma unroll
        for (int co = 0; co < 4; ++co)
          if (co * 9 + tap < KMAX) {
            acc[co * 9 + tap][0] = fmaf(d[co], f0.x, acc[co * 9 + tap][0]);
            acc[co * 9 + tap][1] = fmaf(d[co], f0.y, acc[co * 9 + tap][1]);
            acc[co * 9 + tap][2] = fmaf(d[co], f1.x, acc[co * 9 + tap][2]);
            acc[co * 9 + tap][3] = fmaf(d[co], f1.y, acc[co * 9 + tap][3]);
          }
      }
    }
#pragma unroll
    for (int co = 0; co < 4; ++co)
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (co < Cout && co * 9 + tap < KMAX) atomicAdd(&sacc[(tap * Cout + co) * Cin + qd * 4 + e], acc[co * 9 + tap][e]);
    if (qd == 0) {
#pragma unroll
      for (int co = 0; co < 4; ++co)
        if (co < Cout) atomicAdd(&sacc[nout + co], bacc[co]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nout; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
  if (dbias && threadIdx.x < Cout) atomicAdd(dbias + threadIdx.x, sacc[nout + threadIdx.x]);
}

__global__ void zero_f32_kernel(float* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

// fast paths for the UNet's own shapes (convio.cu)
bool conv_in_fwd_fast(const float* x, const float* w, const float* bias, __half* y, int64_t ldy, int B, int Cin, int H,
                      int W, int Cout, cudaStream_t st, float* gn_sums, int64_t ld_sums);
bool conv_in_fwd_sums_ok(int Cin, int H, int W, int Cout);
bool conv_out_dgrad_fast(const float* w, const float* dy, __half* dx, int64_t lddx, int B, int Cin, int H, int W, int Cout,
                         cudaStream_t st);
bool conv_out_fwd_fast(const __half* x, int64_t ldx, const float* w, const float* bias, float* y, int B, int Cin, int H,
                       int W, int Cout, cudaStream_t st);
bool conv_out_wgrad_fast(const __half* x, int64_t ldx, const float* dy, float* dw, float* dbias, int B, int Cin, int H,
                         int W, int Cout, cudaStream_t st);
bool conv_in_wgrad_fast(const float* x, const __half* dy, int64_t lddy, float* dw, float* dbias, int B, int Cin, int H,
                        int W, int Cout, cudaStream_t st);

// exported to api.cu ---------------------------------------------------------------------------
int simt_conv_launch(const bd_conv_args* a, bool dgrad, cudaStream_t st) {
  ConvGeom g;
  const int stride = a->mode == BD_CONV_S2_PAD01 ? 2 : 1;
  const int pad = a->mode == BD_CONV_S2_PAD01 ? a->pad : a->ksize / 2;
  const int Ho = stride == 2 ? (a->H + (a->pad ? 2 : 1) - 3) / 2 + 1 : a->H;
  const int Wo = stride == 2 ? (a->W + (a->pad ? 2 : 1) - 3) / 2 + 1 : a->W;
  g.B = a->B; g.ks = a->ksize; g.stride = stride; g.pad = pad; g.transposed = dgrad ? 1 : 0;
  if (!dgrad) { g.OH = Ho; g.OW = Wo; g.SH = a->H; g.SW = a->W; g.Ck = a->Cin; g.Nout = a->Cout; }
  else        { g.OH = a->H; g.OW = a->W; g.SH = Ho; g.SW = Wo; g.Ck = a->Cout; g.Nout = a->Cin; }
  ConvEpilogue ep{a->bias, dgrad ? nullptr : a->bias2, a->rowbias, a->ld_rowbias, (const __half*)a->residual, a->ld_res,
                  a->out_scale, a->y, a->ld_y, a->out_dtype == BD_OUT_F32};
  const int64_t M = (int64_t)g.B * g.OH * g.OW;
  dim3 grid(ceil_div(M, TM), ceil_div(g.Nout, TN));
  simt_conv_kernel<<<grid, 256, 0, st>>>(g, (const __half*)a->x, a->ld_x, (const __half*)a->w,
                                         dgrad ? nullptr : (const __half*)a->x2, a->ld_x2, a->Cin2,
                                         (const __half*)a->w2, ep);
  count_launch(1);
  return 0;
}

int simt_wgrad_launch(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, float* dw, float* dbias, int B, int H,
                      int W, int Cin, int Cout, int ksize, int mode, int pad_in, int accumulate, cudaStream_t st) {
  ConvGeom g;
  const int stride = mode == BD_CONV_S2_PAD01 ? 2 : 1;
  const int pad = mode == BD_CONV_S2_PAD01 ? pad_in : ksize / 2;
  const int Ho = stride == 2 ? (H + (pad_in ? 2 : 1) - 3) / 2 + 1 : H;
  const int Wo = stride == 2 ? (W + (pad_in ? 2 : 1) - 3) / 2 + 1 : W;
  g.B = B; g.ks = ksize; g.stride = stride; g.pad = pad; g.transposed = 0;
  g.OH = Ho; g.OW = Wo; g.SH = H; g.SW = W; g.Ck = Cin; g.Nout = Cout;
  const int taps = ksize * ksize;
  if (!accumulate) {
    zero_f32_kernel<<<ceil_div((size_t)Cout * Cin * taps, 256 * 8), 256, 0, st>>>(dw, (size_t)Cout * Cin * taps);
    if (dbias) zero_f32_kernel<<<1, 256, 0, st>>>(dbias, Cout);
    count_launch(dbias ? 2 : 1);
  }
  const int tiles = ceil_div(Cout, TM) * ceil_div(Cin, TN);
  const int64_t P = (int64_t)B * Ho * Wo;
  int splits = ceil_div(2 * num_sms(), tiles * taps);
  int maxs = (int)((P + 255) / 256);
  if (splits > maxs) splits = maxs;
  if (splits < 1) splits = 1;
  simt_wgrad_kernel<<<dim3(tiles, taps, splits), 256, 0, st>>>(g, (const __half*)x, ld_x, (const __half*)dy, ld_dy, dw,
                                                               Cin, Cout, splits);
  count_launch(1);
  if (dbias) {
    int cs = ceil_div(num_sms(), ceil_div(Cout, 256));
    if (cs > (int)((P + 63) / 64)) cs = (int)((P + 63) / 64);
    colsum_kernel<<<dim3(ceil_div(Cout, 256), 1, cs), 256, 0, st>>>((const __half*)dy, ld_dy, dbias, 0, P, Cout, cs);
    count_launch(1);
  }
  return 0;
}

}  // namespace bd

using namespace bd;

extern "C" {

int bd_sgemm(const float* A, int64_t sam, int64_t sak, const float* Bm, int64_t sbk, int64_t sbn, float* C,
             int64_t scm, int64_t scn, const float* bias, int M, int N, int K, int accumulate, int act_silu_a,
             void* stream) {
  BD_CHECK_ARG(A && Bm && C && M > 0 && N > 0 && K > 0, "bd_sgemm: bad argument");
  const int gx = ceil_div(M, TM), gy = ceil_div(N, TN);
  int ksplit = 1;
  if (gx * gy < num_sms() && K >= 512 && scn == 1 && scm == N) {  // few tiles, long reduction, dense C
    ksplit = ceil_div(2 * num_sms(), gx * gy);
    if (ksplit > K / 128) ksplit = K / 128;
    if (ksplit < 1) ksplit = 1;
  }
  if (ksplit > 1 && !accumulate) {
    zero_f32_kernel<<<ceil_div((size_t)M * N, 2048), 256, 0, (cudaStream_t)stream>>>(C, (size_t)M * N);
    count_launch(1);
  }
  sgemm_kernel<<<dim3(gx, gy, ksplit), 256, 0, (cudaStream_t)stream>>>(
      A, sam, sak, Bm, sbk, sbn, C, scm, scn, bias, M, N, K, accumulate, act_silu_a, ksplit);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_temb_mlp(const int64_t* t, const float* w1, const float* b1, const float* w2, const float* b2, float* sin_out,
                float* h1, float* emb, void* silu_emb_f16, int B, int dim, int temb, int flip_sin_to_cos,
                const float* freqs, void* stream) {
  BD_CHECK_ARG(t && w1 && b1 && w2 && b2 && emb && freqs && B > 0 && dim > 0 && dim % 4 == 0 && temb > 0 && temb % 4 == 0 &&
                   !((uintptr_t)w1 & 15) && !((uintptr_t)w2 & 15),
               "bd_temb_mlp: bad argument (dim, temb multiples of 4; 16-byte aligned weights)");
  temb_mlp_kernel<<<ceil_div(B, TEMB_SPB), TEMB_THREADS, (size_t)TEMB_SPB * (dim + temb) * sizeof(float), (cudaStream_t)stream>>>(
      t, w1, b1, w2, b2, sin_out, h1, emb, (__half*)silu_emb_f16, B, dim, temb, flip_sin_to_cos, freqs);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_colsum_f16(const void* x, int64_t ld_x, float* out, int64_t ld_out, int B, int64_t rows_per_b, int C,
                  int accumulate, void* stream) {
  BD_CHECK_ARG(x && out && C % 8 == 0 && ld_x % 8 == 0 && B > 0, "bd_colsum_f16: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  BD_CHECK_ARG(accumulate || B == 1 || ld_out == C, "bd_colsum_f16: non-accumulating call needs contiguous out rows");
  if (!accumulate) {
    zero_f32_kernel<<<ceil_div((size_t)B * C, 2048), 256, 0, st>>>(out, (size_t)B * C);
    count_launch(1);
  }
  int cs = ceil_div(2 * num_sms(), ceil_div(C, 256) * B);
  if (cs > (int)((rows_per_b + 63) / 64)) cs = (int)((rows_per_b + 63) / 64);
  if (cs < 1) cs = 1;
  colsum_kernel<<<dim3(ceil_div(C, 256), B, cs), 256, 0, st>>>((const __half*)x, ld_x, out, ld_out, rows_per_b, C, cs);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_bias_from_gsum(const void* jobs, int njobs, int max_c, int B, void* stream) {
  BD_CHECK_ARG(jobs && njobs >= 0 && max_c > 0 && B > 0, "bd_bias_from_gsum: bad argument");
  if (njobs == 0) return BD_OK;
  bias_from_gsum_kernel<<<dim3(ceil_div(max_c, 256), njobs), 256, 0, (cudaStream_t)stream>>>((const long long*)jobs, B);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_conv_in_fwd_gn_sums_supported(int Cin, int H, int W, int Cout) {
  // off by default with the generic kernels' statistics (BD_GN_SUMS_GENERIC=1, see umma.cu fprop_gn_sums_supported)
  return getenv("BD_NO_CONVIO") == nullptr && getenv("BD_NO_GN_SUMS") == nullptr && getenv("BD_GN_SUMS_GENERIC") != nullptr &&
         conv_in_fwd_sums_ok(Cin, H, W, Cout);
}

int bd_conv_in_fwd_sums(const float* x_nchw, const float* w_packed, const float* bias, void* y, int64_t ld_y, float* gn_sums,
                        int64_t ld_sums, int B, int Cin, int H, int W, int Cout, void* stream) {
  BD_CHECK_ARG(x_nchw && w_packed && y && gn_sums && Cin > 0 && Cin <= 4 && Cout % 8 == 0 && ld_y % 8 == 0, "bd_conv_in_fwd_sums: bad argument");
  if (!bd_conv_in_fwd_gn_sums_supported(Cin, H, W, Cout) ||
      !conv_in_fwd_fast(x_nchw, w_packed, bias, (__half*)y, ld_y, B, Cin, H, W, Cout, (cudaStream_t)stream, gn_sums, ld_sums)) {
    set_error("bd_conv_in_fwd_sums: shape outside the register-tiled kernel (query bd_conv_in_fwd_gn_sums_supported)");
    return BD_ERR_UNSUPPORTED;
  }
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_conv_in_fwd(const float* x_nchw, const float* w_packed, const float* bias, void* y, int64_t ld_y, int B, int Cin,
                   int H, int W, int Cout, void* stream) {
  BD_CHECK_ARG(x_nchw && w_packed && y && Cin > 0 && Cin <= 4 && Cout % 8 == 0 && ld_y % 8 == 0, "bd_conv_in_fwd: need Cin <= 4, Cout %% 8 == 0");
  if (getenv("BD_NO_CONVIO") == nullptr &&
      conv_in_fwd_fast(x_nchw, w_packed, bias, (__half*)y, ld_y, B, Cin, H, W, Cout, (cudaStream_t)stream, nullptr, 0)) {
    BD_CHECK_LAUNCH();
    return BD_OK;
  }
  size_t smem = ((size_t)Cin * 9 * Cout + Cout) * sizeof(float);
  BD_CHECK_ARG(smem <= 200 * 1024, "bd_conv_in_fwd: Cout too large");
  if (smem > 48 * 1024) cudaFuncSetAttribute(conv_in_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  size_t total = (size_t)B * H * W * (Cout / 8);
  int grid = (int)((total + 255) / 256);
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  conv_in_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x_nchw, w_packed, bias, (__half*)y, ld_y, B, Cin, H, W, Cout);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_conv_in_wgrad(const float* x_nchw, const void* dy, int64_t ld_dy, float* dw, float* dbias, int B, int Cin,
                     int H, int W, int Cout, int accumulate, void* stream) {
  BD_CHECK_ARG(x_nchw && dy && dw && Cin > 0 && Cin <= 4, "bd_conv_in_wgrad: need Cin <= 4");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    zero_f32_kernel<<<ceil_div((size_t)Cout * Cin * 9, 2048), 256, 0, st>>>(dw, (size_t)Cout * Cin * 9);
    if (dbias) zero_f32_kernel<<<1, 256, 0, st>>>(dbias, Cout);
    count_launch(2);
  }
  if (getenv("BD_NO_CONVIO") == nullptr &&
      conv_in_wgrad_fast(x_nchw, (const __half*)dy, ld_dy, dw, dbias, B, Cin, H, W, Cout, st)) {
    BD_CHECK_LAUNCH();
    return BD_OK;
  }
  BD_CHECK_ARG(Cout <= 256, "bd_conv_in_wgrad: Cout <= 256");
  conv_in_wgrad_kernel<36><<<2 * num_sms(), 256, ((size_t)Cout * (Cin * 9 + 1) + 64 * 36) * sizeof(float), st>>>(x_nchw, (const __half*)dy, ld_dy, dw, dbias, B, Cin, H, W, Cout);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_conv_out_fwd(const void* x, int64_t ld_x, const float* w_packed, const float* bias, float* y_nchw, int B, int Cin,
                    int H, int W, int Cout, void* stream) {
  BD_CHECK_ARG(x && w_packed && y_nchw && Cout > 0 && Cout <= 4 && Cin % 4 == 0 && ld_x % 4 == 0, "bd_conv_out_fwd: need Cout <= 4, Cin %% 4 == 0");
  if (getenv("BD_NO_CONVIO") == nullptr &&
      conv_out_fwd_fast((const __half*)x, ld_x, w_packed, bias, y_nchw, B, Cin, H, W, Cout, (cudaStream_t)stream)) {
    BD_CHECK_LAUNCH();
    return BD_OK;
  }
  size_t smem = (size_t)Cout * Cin * 9 * sizeof(float);
  BD_CHECK_ARG(smem <= 200 * 1024, "bd_conv_out_fwd: Cin too large");
  if (smem > 48 * 1024) cudaFuncSetAttribute(conv_out_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  conv_out_fwd_kernel<<<4 * num_sms(), 256, smem, (cudaStream_t)stream>>>((const __half*)x, ld_x, w_packed, bias, y_nchw, B, Cin, H, W, Cout);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

int bd_conv_out_bwd(const void* x, int64_t ld_x, const float* w_packed, const float* dy_nchw, void* dx, int64_t ld_dx,
                    float* dw, float* dbias, int B, int Cin, int H, int W, int Cout, int accumulate, void* stream) {
  // dx == nullptr: parameter gradients only; dw == nullptr: data gradient only (the engine runs the two halves on
  // different streams: only the data gradient is on the critical path of backward)
  BD_CHECK_ARG(x && w_packed && dy_nchw && (dx || dw) && Cout > 0 && Cout <= 4 && Cin % 8 == 0 && ld_dx % 8 == 0, "bd_conv_out_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = (size_t)Cout * Cin * 9 * sizeof(float);
  BD_CHECK_ARG(smem <= 200 * 1024, "bd_conv_out_bwd: Cin too large");
  if (smem > 48 * 1024) cudaFuncSetAttribute(conv_out_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  size_t total = (size_t)B * H * W * (Cin / 8);
  int grid = (int)((total + 255) / 256);
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  const bool fast = getenv("BD_NO_CONVIO") == nullptr;
  if (dx && !(fast && conv_out_dgrad_fast(w_packed, dy_nchw, (__half*)dx, ld_dx, B, Cin, H, W, Cout, st))) {
    conv_out_dgrad_kernel<<<grid, 256, smem, st>>>(w_packed, dy_nchw, (__half*)dx, ld_dx, B, Cin, H, W, Cout);
    count_launch(1);
  }
  if (!dw) {
    BD_CHECK_LAUNCH();
    return BD_OK;
  }
  if (!accumulate) {
    zero_f32_kernel<<<ceil_div((size_t)Cout * Cin * 9, 2048), 256, 0, st>>>(dw, (size_t)Cout * Cin * 9);
    if (dbias) zero_f32_kernel<<<1, 256, 0, st>>>(dbias, Cout);
    count_launch(2);
  }
  if (fast && conv_out_wgrad_fast((const __half*)x, ld_x, dy_nchw, dw, dbias, B, Cin, H, W, Cout, st)) {
    BD_CHECK_LAUNCH();
    return BD_OK;
  }
  BD_CHECK_ARG(Cin % 4 == 0 && Cin / 4 <= 256 && 256 % (Cin / 4) == 0, "bd_conv_out_bwd: Cin/4 must divide 256");
  conv_out_wgrad_kernel<36><<<2 * num_sms(), 256, ((size_t)9 * Cout * Cin + 4) * sizeof(float), st>>>((const __half*)x, ld_x, dy_nchw, dw, dbias, B, Cin, H, W, Cout);
  count_launch(1);
  BD_CHECK_LAUNCH();
  return BD_OK;
}

}  // extern "C"
