// First / last 3x3 convolution of the UNet (D/models/unet_2d.py:124 conv_in, :217 conv_out): one side has <= 4
// channels (the image), the other the UNet width.  K = 27 is far too thin for the tensor cores, and the layers move
// 33.5 MB each at the CIFAR-10 size, so they are written as register-tiled fp32 SIMT kernels whose floor is the FMA
// pipe (453 MFMA = 12.6 us at B=128) next to the HBM time of the wide tensor (5 us):
//   narrow_to_wide_kernel : conv_in forward and conv_out dgrad   (thread = 8 pixels x 8 channels, weights in smem)
//   wide_to_narrow_kernel : conv_out forward   (lane = 4 channels with the 27 x COUT x 4 weights in REGISTERS, a warp
//                           walks a row with a sliding 3-column window; 32 partial sums per butterfly reduction)
//   wide_wgrad_kernel     : conv_out wgrad     (same walk; 9 x COUT x 4 accumulators per lane)
//   narrow_wgrad_kernel   : conv_in wgrad      (lane = 4 output channels, 27 x 4 accumulators; image patch in smem)
// The general (any shape) kernels stay in simt.cu; api entry points try these first.
#include "common.cuh"

namespace bd {

// ---------------------------------------------------------------------------------------------------------------
// narrow -> wide:  y[p][o] = bias[o] + sum_{c,r,s} src[b][c][h+r-1][w+s-1] * wk[c*9+r*3+s][o]
// src fp32 NCHW with Cs <= 4 channels; y fp16 NHWC view.  Weight element (tap, c, o) sits at w[tap*st_tap + c*st_c +
// o*st_o]; flip = 1 mirrors the taps (dgrad of a stride-1 conv).
// ---------------------------------------------------------------------------------------------------------------
constexpr int NW_PX = 8;        // pixels per thread (along W)
constexpr int NW_PITCH = 12;    // patch row: 10 values + 2 pad -> three aligned float4

__global__ void __launch_bounds__(256, 2) narrow_to_wide_kernel(const float* __restrict__ src, const float* __restrict__ w,
                                                                int st_tap, int st_c, int st_o, int flip,
                                                                const float* __restrict__ bias, __half* __restrict__ y,
                                                                int64_t ldy, int B, int Cs, int H, int W, int Cw,
                                                                float* __restrict__ gn_sums, int64_t ld_sums) {
  extern __shared__ __align__(16) float nw_smem[];
  const int K = Cs * 9, C8 = Cw / 8, PG = blockDim.x / C8;
  float* sw = nw_smem;                   // [K][Cw]
  float* sb = sw + K * Cw;               // [Cw]
  float* patch = sb + Cw;                // [PG][Cs*3][NW_PITCH]
  float* gred = patch + PG * Cs * 3 * NW_PITCH;   // [PG][Cw][2], only when gn_sums (GroupNorm statistics of y, see bd_conv_args)
  for (int i = threadIdx.x; i < K * Cw; i += blockDim.x) {
    const int o = i % Cw, k = i / Cw, c = k / 9, t = k % 9;
    const int tap = flip ? 8 - t : t;
    sw[i] = w[(int64_t)tap * st_tap + (int64_t)c * st_c + (int64_t)o * st_o];
  }
  for (int i = threadIdx.x; i < Cw; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  const int v = threadIdx.x % C8, g = threadIdx.x / C8;
  const int64_t ngroups = (int64_t)B * H * W / NW_PX;
  const int64_t ntiles = (ngroups + PG - 1) / PG;
  const int prow = Cs * 3;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();  // weights staged / previous tile's patch consumed
    // the tile's first pixel group: one set of divisions per thread; later groups are reached by carrying
    const int64_t gid0 = tile * PG;
    const int64_t pix0 = gid0 * NW_PX;
    const int tw0 = (int)(pix0 % W);
    const int64_t tq = pix0 / W;
    const int th0 = (int)(tq % H), tb0 = (int)(tq / H);
    for (int i = threadIdx.x; i < PG * prow * 10; i += blockDim.x) {
      const int j = i % 10, rest = i / 10, cr = rest % prow, gg = rest / prow;
      float val = 0.f;
      if (gid0 + gg < ngroups) {
        int w0 = tw0 + gg * NW_PX, h = th0, b = tb0;
        while (w0 >= W) { w0 -= W; if (++h == H) { h = 0; ++b; } }
        const int c = cr / 3, r = cr - c * 3;
        const int sh = h + r - 1, sx = w0 - 1 + j;
        if (sh >= 0 && sh < H && sx >= 0 && sx < W) val = src[(((int64_t)b * Cs + c) * H + sh) * W + sx];
      }
      patch[(gg * prow + cr) * NW_PITCH + j] = val;
    }
    __syncthreads();
    const int64_t gid = tile * PG + g;
    const bool active = gid < ngroups && g < PG;
    float gs1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, gs2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active) {
    float acc[NW_PX][8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(sb + v * 8), b1 = *reinterpret_cast<const float4*>(sb + v * 8 + 4);
#pragma unroll
      for (int j = 0; j < NW_PX; ++j) {
        acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
        acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
      }
    }
    const float* pg = patch + g * prow * NW_PITCH;
    for (int cr = 0; cr < prow; ++cr) {
      float xr[12];
      {
        const float4 a = *reinterpret_cast<const float4*>(pg + cr * NW_PITCH);
        const float4 b = *reinterpret_cast<const float4*>(pg + cr * NW_PITCH + 4);
        const float4 c = *reinterpret_cast<const float4*>(pg + cr * NW_PITCH + 8);
        xr[0] = a.x; xr[1] = a.y; xr[2] = a.z; xr[3] = a.w; xr[4] = b.x; xr[5] = b.y; xr[6] = b.z; xr[7] = b.w;
        xr[8] = c.x; xr[9] = c.y; xr[10] = c.z; xr[11] = c.w;
      }
      const int kbase = (cr / 3) * 9 + (cr % 3) * 3;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float* wr = sw + (kbase + s) * Cw + v * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < NW_PX; ++j)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[j][k] = fmaf(xr[j + s], wv[k], acc[j][k]);
      }
    }
    __half* yp = y + gid * NW_PX * ldy + v * 8;
#pragma unroll
    for (int j = 0; j < NW_PX; ++j) {
      const half8 hv = pack8(acc[j]);
      *reinterpret_cast<half8*>(yp + j * ldy) = hv;
      if (gn_sums) {
        float f[8];
        unpack8(hv, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { gs1[k] += f[k]; gs2[k] = fmaf(f[k], f[k], gs2[k]); }
      }
    }
    }  // active
    if (gn_sums) {
      // the tile's PG x 8 pixels belong to ONE image (launcher: H*W % (PG*8) == 0): fold the pixel groups through shared
      // memory, one atomic per (tile, channel, quantity)
      if (g < PG) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          gred[(g * Cw + v * 8 + k) * 2] = gs1[k];
          gred[(g * Cw + v * 8 + k) * 2 + 1] = gs2[k];
        }
      }
      __syncthreads();
      const int64_t smp = (tile * PG * NW_PX) / ((int64_t)H * W);
      for (int i = threadIdx.x; i < Cw * 2; i += blockDim.x) {
        float t = 0.f;
        for (int gg = 0; gg < PG; ++gg) t += gred[gg * Cw * 2 + i];
        atomicAdd(gn_sums + smp * ld_sums + i, t);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// shared pieces of the warp-walk kernels: lane owns channels 4*lane .. 4*lane+3 of a 128-channel fp16 NHWC tensor
// ---------------------------------------------------------------------------------------------------------------
constexpr int WALK_SEG = 16;  // pixels per work item (a row segment)

struct Col3 {
  float v[3][4];
};


struct RowPtrs {
  const __half* p[3];  // rows h-1, h, h+1 at (column 0, this lane's channels); nullptr outside the image
};

__device__ __forceinline__ void row_ptrs(RowPtrs& rp, const __half* __restrict__ x, int64_t ldx, int b, int h, int H,
                                         int W, int lane) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int sh = h + r - 1;
    rp.p[r] = (sh >= 0 && sh < H) ? x + ((int64_t)b * H + sh) * W * ldx + lane * 4 : nullptr;
  }
}

__device__ __forceinline__ void load_raw(uint2 (&raw)[3], const RowPtrs& rp, int64_t ldx, int wcol, int W) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
    raw[r] = (rp.p[r] != nullptr && wcol >= 0 && wcol < W) ? *reinterpret_cast<const uint2*>(rp.p[r] + (int64_t)wcol * ldx)
                                                          : make_uint2(0u, 0u);
}

__device__ __forceinline__ void cvt_col(Col3& c, const uint2 (&raw)[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw[r].x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw[r].y));
    c.v[r][0] = f0.x; c.v[r][1] = f0.y; c.v[r][2] = f1.x; c.v[r][3] = f1.y;
  }
}

// window set-up of a row segment: columns w0-1 and w0 converted, w0+1 .. w0+PD in flight (PD = prefetch distance)
template <int PD>
__device__ __forceinline__ void walk_begin(Col3& ca, Col3& cb, uint2 (&ring)[PD][3], const RowPtrs& rp, int64_t ldx,
                                           int w0, int W) {
  uint2 t0[3], t1[3];
  load_raw(t0, rp, ldx, w0 - 1, W);
  load_raw(t1, rp, ldx, w0, W);
#pragma unroll
  for (int d = 0; d < PD; ++d) load_raw(ring[d], rp, ldx, w0 + 1 + d, W);
  cvt_col(ca, t0);
  cvt_col(cb, t1);
}

// 32 per-lane values -> lane l ends with the warp-wide sum of value l  (31 shuffles instead of 32 x 5)
__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = lane & off;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// conv_out forward: y[b][co][h][w] = bias[co] + sum_{tap,ci} x[p+tap][ci] * w[tap][co][ci];  Cin == 128
template <int COUT>
__global__ void __launch_bounds__(256, 1) wide_to_narrow_kernel(const __half* __restrict__ x, int64_t ldx,
                                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                                float* __restrict__ y, int B, int H, int W) {
  constexpr int PD = 4;  // prefetch distance; divides the 8-pixel reduction batch, so ring slots are compile-time
  const int lane = threadIdx.x & 31;
  float wr[9][COUT][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      const float4 f = *reinterpret_cast<const float4*>(w + (t * COUT + co) * 128 + lane * 4);
      wr[t][co][0] = f.x; wr[t][co][1] = f.y; wr[t][co][2] = f.z; wr[t][co][3] = f.w;
    }
  float bv = 0.f;
  if (bias && (lane & 3) < COUT) bv = bias[lane & 3];
  const int nseg = (W + WALK_SEG - 1) / WALK_SEG;
  const int64_t total = (int64_t)B * H * nseg;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t it = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < total; it += nwarps) {
    const int w0 = (int)(it % nseg) * WALK_SEG;
    const int64_t q = it / nseg;
    const int h = (int)(q % H), b = (int)(q / H);
    Col3 ca, cb, cc;
    RowPtrs rp;
    uint2 ring[PD][3];
    row_ptrs(rp, x, ldx, b, h, H, W, lane);
    walk_begin(ca, cb, ring, rp, ldx, w0, W);
#pragma unroll 1
    for (int j0 = 0; j0 < WALK_SEG; j0 += 8) {
      float red[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cvt_col(cc, ring[j % PD]);
        if (j0 + j + PD < WALK_SEG) load_raw(ring[j % PD], rp, ldx, w0 + j0 + j + 1 + PD, W);
#pragma unroll
        for (int co = 0; co < 4; ++co) {
          float a = 0.f;
          if (co < COUT) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                a = fmaf(ca.v[r][k], wr[r * 3 + 0][co < COUT ? co : 0][k], a);
                a = fmaf(cb.v[r][k], wr[r * 3 + 1][co < COUT ? co : 0][k], a);
                a = fmaf(cc.v[r][k], wr[r * 3 + 2][co < COUT ? co : 0][k], a);
              }
          }
          red[j * 4 + co] = a;
        }
        ca = cb;
        cb = cc;
      }
      const float sum = butterfly32(red, lane);
      const int j = lane >> 2, co = lane & 3, wx = w0 + j0 + j;
      if (co < COUT && wx < W) y[(((int64_t)b * COUT + co) * H + h) * W + wx] = sum + bv;
    }
  }
}

// conv_out wgrad: dW[tap][co][ci] += sum_p dy[b][co][p] * x[p+tap][ci];  dbias[co] += sum_p dy.   Cin == 128
template <int COUT>
__global__ void __launch_bounds__(256, 1) wide_wgrad_kernel(const __half* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ dy, float* __restrict__ dw,
                                                            float* __restrict__ dbias, int B, int H, int W) {
  extern __shared__ __align__(16) float ww_smem[];  // [9*COUT*128] + [4]
  constexpr int NOUT = 9 * COUT * 128;
  for (int i = threadIdx.x; i < NOUT + 4; i += blockDim.x) ww_smem[i] = 0.f;
  __syncthreads();
  constexpr int PD = 2;  // prefetch distance (register budget: 108 accumulators)
  const int lane = threadIdx.x & 31;
  float acc[9][COUT][4];
  float bacc[COUT];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[t][co][0] = acc[t][co][1] = acc[t][co][2] = acc[t][co][3] = 0.f;
#pragma unroll
  for (int co = 0; co < COUT; ++co) bacc[co] = 0.f;
  const int nseg = (W + WALK_SEG - 1) / WALK_SEG;
  const int64_t total = (int64_t)B * H * nseg;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t it = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < total; it += nwarps) {
    const int w0 = (int)(it % nseg) * WALK_SEG;
    const int64_t q = it / nseg;
    const int h = (int)(q % H), b = (int)(q / H);
    float dseg[COUT];  // lane j holds dy of pixel w0 + j
#pragma unroll
    for (int co = 0; co < COUT; ++co)
      dseg[co] = (lane < WALK_SEG && w0 + lane < W) ? dy[(((int64_t)b * COUT + co) * H + h) * W + w0 + lane] : 0.f;
#pragma unroll
    for (int co = 0; co < COUT; ++co) bacc[co] += dseg[co];
    Col3 ca, cb, cc;
    RowPtrs rp;
    uint2 ring[PD][3];
    row_ptrs(rp, x, ldx, b, h, H, W, lane);
    walk_begin(ca, cb, ring, rp, ldx, w0, W);
#pragma unroll 2
    for (int j = 0; j < WALK_SEG; ++j) {
      cvt_col(cc, ring[j % PD]);
      if (j + PD < WALK_SEG) load_raw(ring[j % PD], rp, ldx, w0 + j + 1 + PD, W);
      float d[COUT];
#pragma unroll
      for (int co = 0; co < COUT; ++co) d[co] = __shfl_sync(0xffffffffu, dseg[co], j);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int co = 0; co < COUT; ++co)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            acc[r * 3 + 0][co][k] = fmaf(d[co], ca.v[r][k], acc[r * 3 + 0][co][k]);
            acc[r * 3 + 1][co][k] = fmaf(d[co], cb.v[r][k], acc[r * 3 + 1][co][k]);
            acc[r * 3 + 2][co][k] = fmaf(d[co], cc.v[r][k], acc[r * 3 + 2][co][k]);
          }
      ca = cb;
      cb = cc;
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int co = 0; co < COUT; ++co)
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&ww_smem[(t * COUT + co) * 128 + lane * 4 + k], acc[t][co][k]);
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    const float s = warp_sum(bacc[co]);
    if (lane == 0) atomicAdd(&ww_smem[NOUT + co], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NOUT; i += blockDim.x) atomicAdd(dw + i, ww_smem[i]);
  if (dbias && threadIdx.x < COUT) atomicAdd(dbias + threadIdx.x, ww_smem[NOUT + threadIdx.x]);
}

// conv_in wgrad: dW[tap][co][ci] += sum_p dY[p][co] * x[b][ci][p+tap];  dbias[co] += sum_p dY[p][co].   Cout == 128
template <int CIN>
__global__ void __launch_bounds__(256, 1) narrow_wgrad_kernel(const float* __restrict__ x, const __half* __restrict__ dy,
                                                              int64_t lddy, float* __restrict__ dw,
                                                              float* __restrict__ dbias, int B, int H, int W) {
  constexpr int K = CIN * 9, PW = WALK_SEG + 2, PATCH = CIN * 3 * PW;
  extern __shared__ __align__(16) float nwg_smem[];  // sacc [K+1][128] | per-warp patches [8][PATCH]
  float* sacc = nwg_smem;
  float* patches = sacc + 128 * (K + 1);
  for (int i = threadIdx.x; i < 128 * (K + 1); i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* patch = patches + warp * PATCH;
  float acc[CIN][9][4];
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < CIN; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[c][t][0] = acc[c][t][1] = acc[c][t][2] = acc[c][t][3] = 0.f;
  const int nseg = (W + WALK_SEG - 1) / WALK_SEG;
  const int64_t total = (int64_t)B * H * nseg;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t it = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; it < total; it += nwarps) {
    const int w0 = (int)(it % nseg) * WALK_SEG;
    const int64_t q = it / nseg;
    const int h = (int)(q % H), b = (int)(q / H);
    __syncwarp();
    for (int i = lane; i < PATCH; i += 32) {
      const int j = i % PW, cr = i / PW, c = cr / 3, r = cr % 3;
      const int sh = h + r - 1, sx = w0 - 1 + j;
      patch[i] = (sh >= 0 && sh < H && sx >= 0 && sx < W) ? x[(((int64_t)b * CIN + c) * H + sh) * W + sx] : 0.f;
    }
    __syncwarp();
    const __half* dp = dy + (((int64_t)b * H + h) * W + w0) * lddy + lane * 4;
    uint2 raws[WALK_SEG];  // the whole segment's dY in flight together with the patch
#pragma unroll
    for (int j = 0; j < WALK_SEG; ++j)
      raws[j] = (w0 + j < W) ? *reinterpret_cast<const uint2*>(dp + j * lddy) : make_uint2(0u, 0u);
#pragma unroll
    for (int j = 0; j < WALK_SEG; ++j) {
      const uint2 raw = raws[j];
      const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
      const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
      const float d[4] = {f0.x, f0.y, f1.x, f1.y};
#pragma unroll
      for (int k = 0; k < 4; ++k) bacc[k] += d[k];
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const float xv = patch[(c * 3 + r) * PW + j + s];
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[c][r * 3 + s][k] = fmaf(d[k], xv, acc[c][r * 3 + s][k]);
          }
    }
  }
  // block partials: sacc[k][co] (conflict-free: a warp touches 128 consecutive floats), bias sums in row K
#pragma unroll
  for (int c = 0; c < CIN; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&sacc[(c * 9 + t) * 128 + lane * 4 + k], acc[c][t][k]);
#pragma unroll
  for (int k = 0; k < 4; ++k) atomicAdd(&sacc[K * 128 + lane * 4 + k], bacc[k]);
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * (K + 1); i += blockDim.x) {
    const int co = i % 128, k = i / 128;
    if (k < K) atomicAdd(dw + ((int64_t)(k % 9) * 128 + co) * CIN + k / 9, sacc[i]);   // packed [tap][Cout][Cin]
    else if (dbias) atomicAdd(dbias + co, sacc[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// launchers (return false when the shape is outside the fast path)
// ---------------------------------------------------------------------------------------------------------------
static bool narrow_to_wide_ok(int Cs, int W, int Cw) {
  const int C8 = Cw / 8;
  return Cs >= 1 && Cs <= 4 && W % NW_PX == 0 && Cw % 8 == 0 && C8 <= 256 && 256 % C8 == 0 && Cw <= 512;
}

static bool narrow_to_wide_launch(const float* src, const float* w, int st_tap, int st_c, int st_o, int flip,
                                  const float* bias, __half* y, int64_t ldy, int B, int Cs, int H, int W, int Cw,
                                  cudaStream_t st, float* gn_sums = nullptr, int64_t ld_sums = 0) {
  if (!narrow_to_wide_ok(Cs, W, Cw) || ldy % 8) return false;
  const int C8 = Cw / 8, PG = 256 / C8;
  if (gn_sums && ((int64_t)H * W) % (PG * NW_PX)) return false;
  const size_t smem = ((size_t)Cs * 9 * Cw + Cw + (size_t)PG * Cs * 3 * NW_PITCH + (gn_sums ? (size_t)PG * Cw * 2 : 0)) * sizeof(float);
  if (smem > 100 * 1024) return false;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(narrow_to_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  const int64_t ntiles = ((int64_t)B * H * W / NW_PX + PG - 1) / PG;
  int grid = 2 * num_sms();
  if (grid > ntiles) grid = (int)ntiles;
  if (grid < 1) return true;
  narrow_to_wide_kernel<<<grid, 256, smem, st>>>(src, w, st_tap, st_c, st_o, flip, bias, y, ldy, B, Cs, H, W, Cw, gn_sums, ld_sums);
  count_launch(1);
  return true;
}

bool conv_in_fwd_fast(const float* x, const float* w, const float* bias, __half* y, int64_t ldy, int B, int Cin, int H,
                      int W, int Cout, cudaStream_t st, float* gn_sums, int64_t ld_sums) {
  // packed [tap][Cout][Cin]: c = ci, o = co
  return narrow_to_wide_launch(x, w, Cout * Cin, 1, Cin, 0, bias, y, ldy, B, Cin, H, W, Cout, st, gn_sums, ld_sums);
}
bool conv_in_fwd_sums_ok(int Cin, int H, int W, int Cout) {
  if (!narrow_to_wide_ok(Cin, W, Cout)) return false;
  const int PG = 256 / (Cout / 8);
  return ((int64_t)H * W) % (PG * NW_PX) == 0 &&
         ((size_t)Cin * 9 * Cout + Cout + (size_t)PG * Cin * 3 * NW_PITCH + (size_t)PG * Cout * 2) * sizeof(float) <= 100 * 1024;
}

bool conv_out_dgrad_fast(const float* w, const float* dy, __half* dx, int64_t lddx, int B, int Cin, int H, int W, int Cout,
                         cudaStream_t st) {
  // packed [tap][Cout][Cin]: the narrow side is co, the wide side ci; taps mirrored
  return narrow_to_wide_launch(dy, w, Cout * Cin, Cin, 1, 1, nullptr, dx, lddx, B, Cout, H, W, Cin, st);
}

bool conv_out_fwd_fast(const __half* x, int64_t ldx, const float* w, const float* bias, float* y, int B, int Cin, int H,
                       int W, int Cout, cudaStream_t st) {
  if (Cin != 128 || Cout != 3 || ldx % 4 || ((uintptr_t)w & 15) || ((uintptr_t)x & 7)) return false;
  wide_to_narrow_kernel<3><<<num_sms(), 256, 0, st>>>(x, ldx, w, bias, y, B, H, W);
  count_launch(1);
  return true;
}

bool conv_out_wgrad_fast(const __half* x, int64_t ldx, const float* dy, float* dw, float* dbias, int B, int Cin, int H,
                         int W, int Cout, cudaStream_t st) {
  if (Cin != 128 || Cout != 3 || ldx % 4 || ((uintptr_t)x & 7)) return false;
  wide_wgrad_kernel<3><<<num_sms(), 256, (9 * 3 * 128 + 4) * sizeof(float), st>>>(x, ldx, dy, dw, dbias, B, H, W);
  count_launch(1);
  return true;
}

bool conv_in_wgrad_fast(const float* x, const __half* dy, int64_t lddy, float* dw, float* dbias, int B, int Cin, int H,
                        int W, int Cout, cudaStream_t st) {
  if (Cin != 3 || Cout != 128 || lddy % 4 || ((uintptr_t)dy & 7)) return false;
  constexpr int K = 27, PATCH = 3 * 3 * (WALK_SEG + 2);
  narrow_wgrad_kernel<3><<<num_sms(), 256, (128 * (K + 1) + 8 * PATCH) * sizeof(float), st>>>(x, dy, lddy, dw, dbias, B, H, W);
  count_launch(1);
  return true;
}

}  // namespace bd
