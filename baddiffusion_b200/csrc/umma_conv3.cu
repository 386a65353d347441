// 3x3 stride-1 convolution on tcgen05 with input-halo reuse ("conv3 halo kernel").
//
// One CTA computes a REGION of up to 4 MMA blocks (each 8 px wide x 16 rows = 128 pixels) x 128 output channels:
// 4 accumulators of 128 columns fill the 512 TMEM columns of the SM.  Per 64-channel k-block
//   * the input region plus its 1-pixel halo is fetched ONCE by a single TMA box {64 ch, rw+2, rh+2, rn} (out-of-
//     bounds zero fill = conv padding) and stays in shared memory for all 9 taps: a tap is nothing but a different
//     START ADDRESS of the UMMA A descriptor (row pitch = halo row, SBO = pitch * 128 B; the SWIZZLE_128B pattern is
//     a function of the absolute smem address, so row-shifted views of one TMA-written tile are valid operands);
//   * each tap's 128x64 weight tile streams through a small ring and is shared by the 4 blocks.
// Shared-memory traffic from L2 per MAC drops ~5x against the one-tile-per-CTA kernel (umma.cu), which measured
// L2->SM bound (361 TFLOP/s at 5.5 TB/s of operand traffic).  Used for conv fwd (K-major weights) and dgrad
// (MN-major view of the same weights, taps flipped), with the fused 1x1-shortcut K segment and the same epilogue.
#include "umma_common.cuh"

namespace bd {
namespace umma {

constexpr int C3_BN = 128;
constexpr int C3_BK = 64;
constexpr int C3_MAXBLK = 4;
constexpr int C3_ASTAGES = 2;
constexpr int C3_BSTAGES = 3;
constexpr int C3_A_STAGE_BYTES = 82 * 1024;           // >= 18*18*2*128 = 82,944 and 18*34*128 = 78,336
constexpr int C3_B_STAGE_BYTES = C3_BN * C3_BK * 2;   // 16 KB
constexpr int C3_BAR_OFFSET = C3_ASTAGES * C3_A_STAGE_BYTES + C3_BSTAGES * C3_B_STAGE_BYTES;
// persistent kernel (H % 32 == 0 only): tighter halo stages make room for the epilogue's transpose tiles
constexpr int C3P_A_STAGE_BYTES = 77 * 1024;          // >= 18*34*128 = 78,336
constexpr int C3P_EPI_PITCH = 80;                     // bytes per staged pixel row (32 channels = 64 B + 16 B: conflict-free stmatrix)
constexpr int C3P_EPI_BYTES = 16 * 16 * C3P_EPI_PITCH;  // 16 warps x 16 pixels
constexpr int C3P_B_OFFSET = C3_ASTAGES * C3P_A_STAGE_BYTES;
constexpr int C3P_EPI_OFFSET = C3P_B_OFFSET + C3_BSTAGES * C3_B_STAGE_BYTES;
constexpr int C3P_BAR_OFFSET = C3P_EPI_OFFSET + C3P_EPI_BYTES;
constexpr int C3P_SMEM = C3P_BAR_OFFSET + (2 * C3_ASTAGES + 2 * C3_BSTAGES + 2) * 8 + 16 + 1024;
static_assert(C3P_SMEM <= 232448, "conv3p shared memory");
constexpr int C3_SMEM = C3_BAR_OFFSET + (2 * C3_ASTAGES + 2 * C3_BSTAGES + 2) * 8 + 16 + 1024;

struct Conv3Params {
  int MB;                                   // MMA blocks in the region
  int blk_row[C3_MAXBLK];                   // halo-row index of each block's pixel (0,0)
  int blk_n[C3_MAXBLK], blk_h[C3_MAXBLK], blk_w[C3_MAXBLK];  // block origin inside the region
  int pitch;                                // halo row pitch in pixels
  int reg_w, reg_h, reg_n;                  // region extent
  int tiles_w, tiles_h;                     // regions per image
  int W, H, NB, HW, N;
  int nkb, nkb2;                            // k-blocks of the 3x3 source / of the fused 1x1 source
  int flip;                                 // dgrad: tap offsets negated
  uint32_t a_bytes;                         // bytes of one halo box
  uint32_t a_sbo, b_lbo, b_sbo, idesc;
  const float* bias;
  const float* bias2;
  const float* rowbias;
  int64_t ld_rowbias;
  const __half* residual;
  int64_t ld_res;
  float scale;
  void* y;
  int64_t ld_y;
  int out_f32;
  int* error_flag;
  float* gn_sums;   // optional (B, ld_sums) = [channel][2]: += per-(sample, channel) sum / sum of squares of the fp16 outputs
  int64_t ld_sums;  //   (statistics of the GroupNorm that consumes y, accumulated here so that pass never reduces)
  long long* dbg;   // optional per-CTA timeline (64 slots per CTA), bring-up only
  int epi_mode;     // bring-up: 1 = skip global stores, 2 = skip phase 2, 3 = skip the whole epilogue
};

template <bool B_MN>
__global__ void __launch_bounds__(576, 1) umma_conv3_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                            const __grid_constant__ CUtensorMap tmA1,
                                                            const __grid_constant__ CUtensorMap tmB0,
                                                            const __grid_constant__ CUtensorMap tmB1,
                                                            const Conv3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + C3_ASTAGES * C3_A_STAGE_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C3_BAR_OFFSET);
  uint64_t* a_empty = a_full + C3_ASTAGES;
  uint64_t* b_full = a_empty + C3_ASTAGES;
  uint64_t* b_empty = b_full + C3_BSTAGES;
  uint64_t* tmem_full = b_empty + C3_BSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, n_tile = blockIdx.y;
  const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, tn = tile / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.reg_w, h0 = th * p.reg_h, n0 = tn * p.reg_n;
  const int ncols = p.MB * C3_BN;  // 256 or 512 TMEM columns

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB0);
    prefetch_tmap(&tmB1);
    for (int s = 0; s < C3_ASTAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C3_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // prologue above overlaps the previous kernel; global memory is touched only below

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true;
      for (int seg = 0; seg < 2 && ok; ++seg) {
        const int nkb = seg ? p.nkb2 : p.nkb;
        const int ntap = seg ? 1 : 9;
        const CUtensorMap* mapA = seg ? &tmA1 : &tmA0;
        const CUtensorMap* mapB = seg ? &tmB1 : &tmB0;
        for (int kb = 0; kb < nkb && ok; ++kb) {
          ok = mbar_wait(&a_empty[as], aph ^ 1, p.error_flag, 1);
          if (!ok) break;
          mbar_expect_tx(&a_full[as], p.a_bytes);
          tma_load_4d(mapA, &a_full[as], smem + as * C3_A_STAGE_BYTES, kb * C3_BK, w0 - 1, h0 - 1, n0);
          for (int t = 0; t < ntap; ++t) {
            ok = mbar_wait(&b_empty[bs], bph ^ 1, p.error_flag, 1);
            if (!ok) break;
            uint8_t* sb = smem_b + bs * C3_B_STAGE_BYTES;
            mbar_expect_tx(&b_full[bs], C3_B_STAGE_BYTES);
            if (!B_MN) {
              tma_load_3d(mapB, &b_full[bs], sb, kb * C3_BK, n_tile * C3_BN, t);
            } else {
              tma_load_3d(mapB, &b_full[bs], sb, n_tile * C3_BN, kb * C3_BK, t);
              tma_load_3d(mapB, &b_full[bs], sb + 64 * C3_BK * 2, n_tile * C3_BN + 64, kb * C3_BK, t);
            }
            if (++bs == C3_BSTAGES) { bs = 0; bph ^= 1; }
          }
          if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true, first = true;
      long long* dbg = p.dbg ? p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 64 : nullptr;
      int di = 0;
      if (dbg) dbg[di++] = clock64();
      for (int seg = 0; seg < 2 && ok; ++seg) {
        const int nkb = seg ? p.nkb2 : p.nkb;
        const int ntap = seg ? 1 : 9;
        for (int kb = 0; kb < nkb && ok; ++kb) {
          ok = mbar_wait(&a_full[as], aph, p.error_flag, 2);
          if (!ok) break;
          if (dbg && di < 40) dbg[di++] = clock64();
          const uint32_t sa = smem_u32(smem + as * C3_A_STAGE_BYTES);
          for (int t = 0; t < ntap; ++t) {
            ok = mbar_wait(&b_full[bs], bph, p.error_flag, 2);
            if (!ok) break;
            if (dbg && di < 40) dbg[di++] = clock64();
            tc_fence_after();
            int dy = seg ? 0 : t / 3 - 1, dx = seg ? 0 : t % 3 - 1;
            if (p.flip) { dy = -dy; dx = -dx; }
            const uint32_t sb = smem_u32(smem_b + bs * C3_B_STAGE_BYTES);
            const int tap_row = (dy + 1) * p.pitch + (dx + 1);
            for (int mb = 0; mb < p.MB; ++mb) {
              const uint32_t a0 = sa + (uint32_t)(p.blk_row[mb] + tap_row) * 128u;
#pragma unroll
              for (int k = 0; k < C3_BK / 16; ++k) {
                const uint64_t ad = make_desc(a0 + k * 32, 1, p.a_sbo);
                const uint64_t bd = make_desc(sb + (B_MN ? k * 2048 : k * 32), p.b_lbo, p.b_sbo);
                umma_f16(tmem_base + mb * C3_BN, ad, bd, p.idesc, (first && k == 0) ? 0u : 1u);
              }
            }
            first = false;
            umma_commit(&b_empty[bs]);
            if (++bs == C3_BSTAGES) { bs = 0; bph ^= 1; }
          }
          if (ok) umma_commit(&a_empty[as]);
          if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
        }
      }
      if (ok) umma_commit(tmem_full);
      if (dbg) dbg[di++] = clock64();
    }
  } else {
    // ===== epilogue: 16 warps; warp w drains TMEM lanes 32*(w%4).. and column quarter (w-2)/4 of every block =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;   // 0..3: 32-column window
    const int r = q * 32 + lane;          // row inside an MMA block: 8 px wide x 16 rows
    const int dh = r >> 3, dw = r & 7;
    const bool ok = mbar_wait(tmem_full, 0, p.error_flag, 3);
    tc_fence_after();
    long long* dbg = (p.dbg && warp == 2 && lane == 0) ? p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 64 : nullptr;
    if (dbg) dbg[48] = clock64();
    if (ok) {
      // all MMAs have retired: the operand stages are free and serve as the fp32 staging tiles (one per warp)
      float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * 32 * (C3_BN / 4 + 4);
      EpiArgs e{p.bias, p.bias2, p.residual, p.ld_res, p.scale, p.y, p.ld_y, p.out_f32};
      for (int mb = 0; mb < p.MB; ++mb) {
        const int n = n0 + p.blk_n[mb], h = h0 + p.blk_h[mb] + dh, w = w0 + p.blk_w[mb] + dw;
        const bool valid = n < p.NB && h < p.H && w < p.W;
        const int64_t m = ((int64_t)n * p.H + h) * p.W + w;
        epilogue_warp<C3_BN / 4>(tmem_base + ((uint32_t)(q * 32) << 16) + mb * C3_BN + half * (C3_BN / 4), stage, lane, m, m,
                                 valid, n_tile * C3_BN + half * (C3_BN / 4), e, p.rowbias, p.ld_rowbias, p.HW);
      }
    }
  }
  if (p.dbg && warp == 2 && lane == 0) p.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 64 + 49] = clock64();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Conv3Call {
  const void* a; int64_t ld_a; int Ca;
  const void* a2; int64_t ld_a2; int Ca2;     // fused 1x1 segment source (nullable)
  int NB, H, W;
  const void* b; int64_t ld_b; int b_rows;    // weights [9][b_rows][ld_b]
  const void* b2; int64_t ld_b2;              // [N][Ca2]
  int N; bool b_mn; bool flip;
  const float* bias; const float* bias2; const float* rowbias; int64_t ld_rowbias;
  const void* residual; int64_t ld_res; float scale; void* y; int64_t ld_y; int out_f32;
  float* gn_sums; int64_t ld_sums;   // see Conv3Params::gn_sums (persistent / cluster kernels, fp16 stmatrix epilogue only)
};

static bool conv3_geometry(int H, int W, Conv3Params* p) {
  // region = 16 px wide; 32 rows of one image (H >= 32) or 16 rows of two images (H == 16)
  if (W % 16 || H % 16) return false;
  p->reg_w = 16;
  p->pitch = 18;
  if (H % 32 == 0) {
    p->reg_h = 32; p->reg_n = 1; p->MB = 4;
    for (int i = 0; i < 4; ++i) {
      p->blk_n[i] = 0; p->blk_h[i] = (i >> 1) * 16; p->blk_w[i] = (i & 1) * 8;
      p->blk_row[i] = p->blk_h[i] * p->pitch + p->blk_w[i];
    }
    p->a_bytes = 18u * 34u * 128u;
  } else {
    if (H != 16) return false;
    p->reg_h = 16; p->reg_n = 2; p->MB = 4;
    for (int i = 0; i < 4; ++i) {
      p->blk_n[i] = i >> 1; p->blk_h[i] = 0; p->blk_w[i] = (i & 1) * 8;
      p->blk_row[i] = p->blk_n[i] * (18 * p->pitch) + p->blk_w[i];
    }
    p->a_bytes = 18u * 18u * 2u * 128u;
  }
  p->tiles_w = W / p->reg_w;
  p->tiles_h = H / p->reg_h;
  return true;
}

int conv3_supported(const Conv3Call& c) {
  Conv3Params p;
  if (getenv("BD_NO_CONV3")) return 0;
  if (c.Ca % 64 || (c.a2 && c.Ca2 % 64) || c.N % C3_BN) return 0;
  if (c.ld_a % 8 || (c.a2 && c.ld_a2 % 8) || c.ld_b % 8 || c.ld_y % 8 || (c.residual && c.ld_res % 8)) return 0;
  if (((uintptr_t)c.a & 15) || ((uintptr_t)c.b & 15) || (c.a2 && ((uintptr_t)c.a2 & 15))) return 0;
  return conv3_geometry(c.H, c.W, &p) ? 1 : 0;
}

// in-epilogue GroupNorm statistics: the stmatrix (fp16, residual folded into the MMA) epilogue of conv3p / conv3c only
int conv3_gn_sums_supported(const Conv3Call& c) {
  if (!conv3_supported(c) || c.out_f32 || c.b_mn || getenv("BD_NO_GN_SUMS")) return 0;
  if (c.residual && !(!c.a2 && c.N <= 512 && c.N % 64 == 0 && !getenv("BD_NO_RES_ID"))) return 0;   // explicit-residual epilogue
  if (c.H == 16 && c.W == 16 && c.N % 256 == 0 && !getenv("BD_NO_CONV3W")) return 1;                 // conv3w: one image per CTA
  if (c.H % 32 || getenv("BD_NO_CONV3P") || getenv("BD_NO_CONV3T")) return 0;
  return 1;
}

int conv3t_launch(const CUtensorMap& mx0, const CUtensorMap& mx1, const CUtensorMap& mw0, const CUtensorMap& mw1,
                  Conv3Params p, bool a_mn, dim3 grid, cudaStream_t st);
int conv3w_launch_fwd(const Conv3Call& c, Conv3Params p, cudaStream_t st);
int conv3c_cluster_size(int H, int W, int NB, int N);
int conv3c_launch(const CUtensorMap& mx0, const CUtensorMap& mx1, const CUtensorMap& mw0, const CUtensorMap& mw1,
                  Conv3Params p, bool a_mn, int cs, cudaStream_t st);

// 512 x 512 fp16 identity: the weight of the "residual" K-segment of the persistent kernel.  y = conv(x) + r is run as
// conv(x) + I * r on the tensor cores (exact: products by 1.0 accumulate in fp32), so the epilogue never has to gather
// the residual in the accumulator's channel-per-lane layout (measured: 31 kcycles per tile as strided 2-byte loads).
constexpr int C3_ID_N = 512;
static __half* g_identity = nullptr;
__global__ void fill_identity_kernel(__half* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * n) p[i] = __float2half((i / n == i % n) ? 1.0f : 0.0f);
}
const __half* conv3_identity() {   // called from bd_init(): never inside a graph capture
  if (!g_identity) {
    if (cudaMalloc(&g_identity, (size_t)C3_ID_N * C3_ID_N * sizeof(__half)) != cudaSuccess) { g_identity = nullptr; return nullptr; }
    fill_identity_kernel<<<ceil_div(C3_ID_N * C3_ID_N, 256), 256>>>(g_identity, C3_ID_N);
    cudaDeviceSynchronize();
  }
  return g_identity;
}

int conv3_launch(const Conv3Call& c_in, cudaStream_t st) {
  Conv3Call c = c_in;
  const bool wide16 = c.H == 16 && c.W == 16 && c.N % 256 == 0 && !getenv("BD_NO_CONV3W");
  if (c.residual && !c.a2 && !c.b_mn && !c.out_f32 && c.N <= C3_ID_N && c.N % 64 == 0 && g_identity && !getenv("BD_NO_RES_ID") &&
      (wide16 || (c.H % 32 == 0 && !getenv("BD_NO_CONV3P") && !getenv("BD_NO_CONV3T")))) {
    c.a2 = c.residual; c.ld_a2 = c.ld_res; c.Ca2 = c.N;
    c.b2 = g_identity; c.ld_b2 = C3_ID_N;
    c.residual = nullptr;
  }
  Conv3Params p;
  memset(&p, 0, sizeof(p));
  if (!conv3_geometry(c.H, c.W, &p)) { set_error("conv3 halo kernel: unsupported geometry %dx%d", c.H, c.W); return BD_ERR_UNSUPPORTED; }
  p.W = c.W; p.H = c.H; p.NB = c.NB; p.HW = c.H * c.W; p.N = c.N;
  p.nkb = c.Ca / 64;
  p.nkb2 = c.a2 ? c.Ca2 / 64 : 0;
  p.flip = c.flip ? 1 : 0;
  p.a_sbo = (uint32_t)(p.pitch * 128) >> 4;
  if (c.b_mn) { p.b_lbo = 512; p.b_sbo = 64; } else { p.b_lbo = 1; p.b_sbo = 64; }
  p.idesc = (1u << 4) | ((c.b_mn ? 1u : 0u) << 16) | ((uint32_t)(C3_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.bias = c.bias; p.bias2 = c.bias2; p.rowbias = c.rowbias; p.ld_rowbias = c.ld_rowbias;
  p.residual = (const __half*)c.residual; p.ld_res = c.ld_res; p.scale = c.scale;
  p.y = c.y; p.ld_y = c.ld_y; p.out_f32 = c.out_f32;
  p.gn_sums = c.gn_sums; p.ld_sums = c.ld_sums;
  p.error_flag = error_flag();
  if (c.gn_sums && !conv3_gn_sums_supported(c_in)) {
    set_error("conv3: gn_sums needs the persistent H %% 32 == 0 kernels with fp16 output (query bd_conv_fwd_gn_sums_supported)");
    return BD_ERR_UNSUPPORTED;
  }
  if (wide16 && !c.residual) return conv3w_launch_fwd(c, p, st);
  {
    const char* e = getenv("BD_CONV3_DBG_PTR");  // device pointer of a (ctas x 64) int64 buffer, bring-up only
    p.dbg = e ? (long long*)strtoull(e, nullptr, 0) : nullptr;
    const char* m = getenv("BD_CONV3_EPI");
    p.epi_mode = m ? atoi(m) : 0;
  }
  // cluster kernel (conv3c): 8 px x 32 row strips, weight tiles fetched in 128/cs-row slices and multicast
  const int cs = (p.reg_n == 1 && !getenv("BD_NO_CONV3T") && !getenv("BD_NO_CONV3P")) ? conv3c_cluster_size(c.H, c.W, c.NB, c.N) : 0;
  if (cs) {
    p.reg_w = 8; p.pitch = 10; p.tiles_w = c.W / 8;
    p.a_bytes = 10u * 34u * 128u;
    p.a_sbo = (uint32_t)(p.pitch * 128) >> 4;
  }
  const uint32_t wrows = cs ? (uint32_t)(128 / cs) : 0;   // rows of one multicast slice
  CUtensorMap ma0, ma1, mb0, mb1;
  const uint32_t box[4] = {64, (uint32_t)p.reg_w + 2, (uint32_t)p.reg_h + 2, (uint32_t)p.reg_n};
  {
    uint64_t dims[4] = {(uint64_t)c.Ca, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_a, (uint64_t)c.W * c.ld_a, (uint64_t)c.H * c.W * c.ld_a};
    if (!make_map(&ma0, c.a, 4, dims, str, box)) return BD_ERR_CUDA;
  }
  if (c.a2) {
    uint64_t dims[4] = {(uint64_t)c.Ca2, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_a2, (uint64_t)c.W * c.ld_a2, (uint64_t)c.H * c.W * c.ld_a2};
    if (!make_map(&ma1, c.a2, 4, dims, str, box)) return BD_ERR_CUDA;
  } else {
    ma1 = ma0;
  }
  {
    // K-major: [tap][N rows][K cols]; MN-major: [tap][K rows][N cols]
    const int cols = c.b_mn ? c.N : c.Ca;
    uint64_t dims[3] = {(uint64_t)cols, (uint64_t)c.b_rows, 9};
    uint64_t str[2] = {(uint64_t)c.ld_b, (uint64_t)c.b_rows * c.ld_b};
    uint32_t bbox[3] = {64, cs ? wrows : (c.b_mn ? 64u : (uint32_t)C3_BN), 1};
    if (!make_map(&mb0, c.b, 3, dims, str, bbox)) return BD_ERR_CUDA;
  }
  if (c.a2) {
    uint64_t dims[3] = {(uint64_t)c.Ca2, (uint64_t)c.N, 1};
    uint64_t str[2] = {(uint64_t)c.ld_b2, (uint64_t)c.N * c.ld_b2};
    uint32_t bbox[3] = {64, cs ? wrows : (uint32_t)C3_BN, 1};
    if (!make_map(&mb1, c.b2, 3, dims, str, bbox)) return BD_ERR_CUDA;
  } else {
    mb1 = mb0;
  }
  const int tiles = p.tiles_w * p.tiles_h * ceil_div(c.NB, p.reg_n);
  dim3 grid(tiles, c.N / C3_BN);
  if (cs) return conv3c_launch(ma0, ma1, mb0, mb1, p, c.b_mn, cs, st);
  if (p.reg_n == 1 && !getenv("BD_NO_CONV3T"))  // H % 32 == 0: weights-as-A / 256-pixel-B orientation
    return conv3t_launch(ma0, ma1, mb0, mb1, p, c.b_mn, grid, st);
  static bool attr_set[2] = {false, false};
  if (c.b_mn) {
    if (!attr_set[1]) { cudaFuncSetAttribute(umma_conv3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM); attr_set[1] = true; }
    launch_pdl(umma_conv3_kernel<true>, grid, dim3(576), C3_SMEM, st, ma0, ma1, mb0, mb1, p);
  } else {
    if (!attr_set[0]) { cudaFuncSetAttribute(umma_conv3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM); attr_set[0] = true; }
    launch_pdl(umma_conv3_kernel<false>, grid, dim3(576), C3_SMEM, st, ma0, ma1, mb0, mb1, p);
  }
  count_launch(1);
  return BD_OK;
}

}  // namespace umma
}  // namespace bd

// =============================================================================================================
// Transposed-orientation variant for images with H % 32 == 0:   D[cout, pixel] = W[cout, k] * X[pixel, k]^T
//   A operand = the 128 x 64 weight tile of a tap (K-major for fwd, MN-major view of the same weights for dgrad),
//   B operand = 256 pixels (8 px x 32 rows) of the resident halo tile, again addressed by a shifted descriptor.
// A 128 x 256 x 16 MMA reads 4 KB (A) + 8 KB (B) of shared memory per 128 cycles = 96 B/cycle, which leaves headroom
// under the 128 B/cycle shared-memory port; the 128 x 128 orientation needs all 128 B/cycle (measured 77 cycles per
// MMA stand-alone, ~110 with the TMA writes of the pipeline in flight).  Two 256-column accumulators = 512 TMEM columns.
// The epilogue transposes through shared memory: TMEM lanes are output channels, so each thread owns ONE channel of 32
// pixels; it scatters them into a [pixel][channel] fp32 tile and the 4 warps of a group then write whole pixel rows.
// =============================================================================================================
namespace bd {
namespace umma {

constexpr int C3T_STAGE_FLOATS = 32 * (C3_BN + 4);  // one [32 pixels][128 channels] fp32 staging tile

template <bool A_MN>
__global__ void __launch_bounds__(576, 1) umma_conv3t_kernel(const __grid_constant__ CUtensorMap tmX0,
                                                             const __grid_constant__ CUtensorMap tmX1,
                                                             const __grid_constant__ CUtensorMap tmW0,
                                                             const __grid_constant__ CUtensorMap tmW1,
                                                             const Conv3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem + C3_ASTAGES * C3_A_STAGE_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C3_BAR_OFFSET);
  uint64_t* a_empty = a_full + C3_ASTAGES;
  uint64_t* b_full = a_empty + C3_ASTAGES;
  uint64_t* b_empty = b_full + C3_BSTAGES;
  uint64_t* tmem_full = b_empty + C3_BSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, n_tile = blockIdx.y;
  const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, tn = tile / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.reg_w, h0 = th * p.reg_h, n0 = tn;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX0);
    prefetch_tmap(&tmX1);
    prefetch_tmap(&tmW0);
    prefetch_tmap(&tmW1);
    for (int s = 0; s < C3_ASTAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C3_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // prologue above overlaps the previous kernel; global memory is touched only below

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true;
      for (int seg = 0; seg < 2 && ok; ++seg) {
        const int nkb = seg ? p.nkb2 : p.nkb;
        const int ntap = seg ? 1 : 9;
        const CUtensorMap* mapX = seg ? &tmX1 : &tmX0;
        const CUtensorMap* mapW = seg ? &tmW1 : &tmW0;
        for (int kb = 0; kb < nkb && ok; ++kb) {
          ok = mbar_wait(&a_empty[as], aph ^ 1, p.error_flag, 1);
          if (!ok) break;
          mbar_expect_tx(&a_full[as], p.a_bytes);
          tma_load_4d(mapX, &a_full[as], smem + as * C3_A_STAGE_BYTES, kb * C3_BK, w0 - 1, h0 - 1, n0);
          for (int t = 0; t < ntap; ++t) {
            ok = mbar_wait(&b_empty[bs], bph ^ 1, p.error_flag, 1);
            if (!ok) break;
            uint8_t* sw = smem_w + bs * C3_B_STAGE_BYTES;
            mbar_expect_tx(&b_full[bs], C3_B_STAGE_BYTES);
            if (!A_MN) {
              tma_load_3d(mapW, &b_full[bs], sw, kb * C3_BK, n_tile * C3_BN, t);
            } else {
              tma_load_3d(mapW, &b_full[bs], sw, n_tile * C3_BN, kb * C3_BK, t);
              tma_load_3d(mapW, &b_full[bs], sw + 64 * C3_BK * 2, n_tile * C3_BN + 64, kb * C3_BK, t);
            }
            if (++bs == C3_BSTAGES) { bs = 0; bph ^= 1; }
          }
          if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true, first = true;
      long long* dbg = p.dbg ? p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 64 : nullptr;
      int di = 0;
      if (dbg) dbg[di++] = clock64();
      for (int seg = 0; seg < 2 && ok; ++seg) {
        const int nkb = seg ? p.nkb2 : p.nkb;
        const int ntap = seg ? 1 : 9;
        for (int kb = 0; kb < nkb && ok; ++kb) {
          ok = mbar_wait(&a_full[as], aph, p.error_flag, 2);
          if (!ok) break;
          if (dbg && di < 40) dbg[di++] = clock64();
          const uint32_t sx = smem_u32(smem + as * C3_A_STAGE_BYTES);
          for (int t = 0; t < ntap; ++t) {
            ok = mbar_wait(&b_full[bs], bph, p.error_flag, 2);
            if (!ok) break;
            if (dbg && di < 40) dbg[di++] = clock64();
            tc_fence_after();
            int dy = seg ? 0 : t / 3 - 1, dx = seg ? 0 : t % 3 - 1;
            if (p.flip) { dy = -dy; dx = -dx; }
            const uint32_t sw = smem_u32(smem_w + bs * C3_B_STAGE_BYTES);
            const int tap_row = (dy + 1) * p.pitch + (dx + 1);
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
              const uint32_t x0 = sx + (uint32_t)(blk * 8 + tap_row) * 128u;
#pragma unroll
              for (int k = 0; k < C3_BK / 16; ++k) {
                const uint64_t wd = A_MN ? make_desc(sw + k * 2048, 512, 64) : make_desc(sw + k * 32, 1, 64);
                const uint64_t xd = make_desc(x0 + k * 32, 1, p.a_sbo);
                umma_f16(tmem_base + blk * 256, wd, xd, p.idesc, (first && k == 0) ? 0u : 1u);
              }
            }
            first = false;
            umma_commit(&b_empty[bs]);
            if (++bs == C3_BSTAGES) { bs = 0; bph ^= 1; }
          }
          if (ok) umma_commit(&a_empty[as]);
          if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
        }
      }
      if (ok) umma_commit(tmem_full);
      if (dbg) dbg[di++] = clock64();
    }
  } else {
    // ===== transposing epilogue: 4 groups of 4 warps; group g drains pixel half (g&1) of block (g>>1); warp quadrant
    // q owns channels 32q.. =====
    const int q = warp & 3, grp4 = (warp - 2) >> 2, grp = grp4 >> 1, phalf = grp4 & 1;
    const int tid_g = ((warp - 2) & 3) * 32 + lane;  // 0..127 inside the group
    const int ch = q * 32 + lane;                    // output channel inside the 128-wide n tile
    const bool ok = mbar_wait(tmem_full, 0, p.error_flag, 3);
    tc_fence_after();
    long long* dbg = (p.dbg && warp == 2 && lane == 0) ? p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 64 : nullptr;
    if (dbg) dbg[48] = clock64();
    if (ok) {
      float* stage0 = reinterpret_cast<float*>(smem) + grp4 * 2 * C3T_STAGE_FLOATS;
      const int gcol = n_tile * C3_BN + ch;
      float badd = 0.f;
      if (p.bias) badd += p.bias[gcol];
      if (p.bias2) badd += p.bias2[gcol];
      if (p.rowbias) badd += p.rowbias[(int64_t)n0 * p.ld_rowbias + gcol];
      const int piece = tid_g & 15, rbase = tid_g >> 4;  // phase 2: 8 channels of rows rbase, rbase+8, ...
      const int col = n_tile * C3_BN + piece * 8;
#pragma unroll 1
      for (int j0 = phalf * 128; j0 < phalf * 128 + 128 && p.epi_mode < 3; j0 += 32) {
        float* stage = stage0 + ((j0 >> 5) & 1) * C3T_STAGE_FLOATS;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + grp * 256 + j0, v);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) stage[jj * (C3_BN + 4) + ch] = __uint_as_float(v[jj]) + badd;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp4) : "memory");
        if (p.epi_mode == 2) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = rbase + 8 * i;           // pixel inside the chunk
          const int j = j0 + row;                  // pixel inside the block: 8 px x 32 rows
          const int h = h0 + (j >> 3), w = w0 + grp * 8 + (j & 7);
          const int64_t m = ((int64_t)n0 * p.H + h) * p.W + w;
          const float* sp = stage + row * (C3_BN + 4) + piece * 8;
          const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 4);
          float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          if (p.residual) {
            float g[8];
            unpack8(*reinterpret_cast<const half8*>(p.residual + m * p.ld_res + col), g);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += g[k];
          }
          if (p.scale != 1.0f) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] *= p.scale;
          }
          if (p.epi_mode == 1 && f[0] != 12345.678f) continue;
          if (p.out_f32) {
            float* yr = reinterpret_cast<float*>(p.y) + m * p.ld_y + col;
            *reinterpret_cast<float4*>(yr) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(yr + 4) = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            *reinterpret_cast<half8*>(reinterpret_cast<__half*>(p.y) + m * p.ld_y + col) = pack8(f);
          }
        }
      }
    }
  }
  if (p.dbg && warp == 2 && lane == 0) p.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 64 + 49] = clock64();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =============================================================================================================
// Persistent variant of the transposed kernel (the default for H % 32 == 0).  One CTA per SM walks a strided list of
// 128-channel x 512-pixel tiles:
//   * the producer warp runs ahead across tile boundaries (both halo stages and the weight ring are refilled while
//     the previous tile is still in its epilogue), so the first MMA of a tile never waits for a cold pipeline, and
//     the TMEM allocation / barrier set-up / tensor-map prefetch are paid once per SM instead of once per tile;
//   * the epilogue goes straight from TMEM to global memory: a TMEM lane is an output CHANNEL, so the 32 lanes of a
//     warp hold 32 consecutive channels of one pixel and every st.global.b16 of the warp is one full 64-byte run of
//     a pixel row -- no shared-memory transpose, no named barriers (the staged epilogue of umma_conv3t_kernel cost
//     ~14 kcycles per tile against 18.4 kcycles of MMA issue);
//   * accumulators are handed back to the MMA warp through a tmem_empty barrier (16 epilogue warps arrive).
// TMEM is full with one tile (2 x 256 fp32 columns), so epilogue and MMA of consecutive tiles do not overlap; a
// 256-pixel tile would double-buffer but doubles the weight traffic per MAC to ~41 B/clk/SM, above what L2 sustains.
// =============================================================================================================
template <bool A_MN, bool GNS = false>   // GNS: accumulate GroupNorm statistics of the output in the stmatrix epilogue (no timeline stamps)
__global__ void __launch_bounds__(576, 1) umma_conv3p_kernel(const __grid_constant__ CUtensorMap tmX0,
                                                             const __grid_constant__ CUtensorMap tmX1,
                                                             const __grid_constant__ CUtensorMap tmW0,
                                                             const __grid_constant__ CUtensorMap tmW1,
                                                             const Conv3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem + C3P_B_OFFSET;
  uint8_t* smem_epi = smem + C3P_EPI_OFFSET;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C3P_BAR_OFFSET);
  uint64_t* a_empty = a_full + C3_ASTAGES;
  uint64_t* b_full = a_empty + C3_ASTAGES;
  uint64_t* b_empty = b_full + C3_BSTAGES;
  uint64_t* tmem_full = b_empty + C3_BSTAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_n = p.tiles_w * p.tiles_h * p.NB;  // pixel tiles per 128-channel slab
  const int total = per_n * (p.N / C3_BN);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX0);
    prefetch_tmap(&tmX1);
    prefetch_tmap(&tmW0);
    prefetch_tmap(&tmW1);
    for (int s = 0; s < C3_ASTAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C3_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 16);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // prologue above overlaps the previous kernel; global memory is touched only below

  if (warp == 0) {
    // ===== TMA producer: free-running over all tiles of this CTA =====
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x) {
        const int n_tile = tile / per_n, rem = tile - n_tile * per_n;
        const int tw = rem % p.tiles_w, th = (rem / p.tiles_w) % p.tiles_h, n0 = rem / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.reg_w, h0 = th * p.reg_h;
        for (int seg = 0; seg < 2 && ok; ++seg) {
          const int nkb = seg ? p.nkb2 : p.nkb;
          const int ntap = seg ? 1 : 9;
          const CUtensorMap* mapX = seg ? &tmX1 : &tmX0;
          const CUtensorMap* mapW = seg ? &tmW1 : &tmW0;
          for (int kb = 0; kb < nkb && ok; ++kb) {
            ok = mbar_wait(&a_empty[as], aph ^ 1, p.error_flag, 1);
            if (!ok) break;
            mbar_expect_tx(&a_full[as], p.a_bytes);
            tma_load_4d(mapX, &a_full[as], smem + as * C3P_A_STAGE_BYTES, kb * C3_BK, w0 - 1, h0 - 1, n0);
            for (int t = 0; t < ntap; ++t) {
              ok = mbar_wait(&b_empty[bs], bph ^ 1, p.error_flag, 1);
              if (!ok) break;
              uint8_t* sw = smem_w + bs * C3_B_STAGE_BYTES;
              mbar_expect_tx(&b_full[bs], C3_B_STAGE_BYTES);
              if (!A_MN) {
                tma_load_3d(mapW, &b_full[bs], sw, kb * C3_BK, n_tile * C3_BN, t);
              } else {
                tma_load_3d(mapW, &b_full[bs], sw, n_tile * C3_BN, kb * C3_BK, t);
                tma_load_3d(mapW, &b_full[bs], sw + 64 * C3_BK * 2, n_tile * C3_BN + 64, kb * C3_BK, t);
              }
              if (++bs == C3_BSTAGES) { bs = 0; bph ^= 1; }
            }
            if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0, tph = 0;
      bool ok = true;
      int di = 0;
      for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x) {
        ok = mbar_wait(tmem_empty, tph ^ 1, p.error_flag, 4);  // accumulators drained by the previous epilogue
        if (!ok) break;
        tc_fence_after();
        if (!GNS && p.dbg && di < 60) p.dbg[(size_t)blockIdx.x * 64 + di] = clock64();   // [4i]: tile i may start
        bool first = true;
        for (int seg = 0; seg < 2 && ok; ++seg) {
          const int nkb = seg ? p.nkb2 : p.nkb;
          const int ntap = seg ? 1 : 9;
          for (int kb = 0; kb < nkb && ok; ++kb) {
            ok = mbar_wait(&a_full[as], aph, p.error_flag, 2);
            if (!ok) break;
            const uint32_t sx = smem_u32(smem + as * C3P_A_STAGE_BYTES);
            for (int t = 0; t < ntap; ++t) {
              ok = mbar_wait(&b_full[bs], bph, p.error_flag, 2);
              if (!ok) break;
              tc_fence_after();
              int dy = seg ? 0 : t / 3 - 1, dx = seg ? 0 : t % 3 - 1;
              if (p.flip) { dy = -dy; dx = -dx; }
              const uint32_t sw = smem_u32(smem_w + bs * C3_B_STAGE_BYTES);
              const int tap_row = (dy + 1) * p.pitch + (dx + 1);
#pragma unroll
              for (int blk = 0; blk < 2; ++blk) {
                const uint32_t x0 = sx + (uint32_t)(blk * 8 + tap_row) * 128u;
#pragma unroll
                for (int k = 0; k < C3_BK / 16; ++k) {
                  const uint64_t wd = A_MN ? make_desc(sw + k * 2048, 512, 64) : make_desc(sw + k * 32, 1, 64);
                  const uint64_t xd = make_desc(x0 + k * 32, 1, p.a_sbo);
                  umma_f16(tmem_base + blk * 256, wd, xd, p.idesc, (first && k == 0) ? 0u : 1u);
                }
              }
              first = false;
              umma_commit(&b_empty[bs]);
              if (++bs == C3_BSTAGES) { bs = 0; bph ^= 1; }
            }
            if (ok) umma_commit(&a_empty[as]);
            if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
          }
        }
        if (ok) umma_commit(tmem_full);
        if (!GNS && p.dbg && di < 60) p.dbg[(size_t)blockIdx.x * 64 + di + 1] = clock64();  // [4i+1]: last MMA of tile i issued
        di += 4;
        tph ^= 1;
      }
    }
  } else {
    // ===== epilogue: 4 groups of 4 warps; group g4 drains pixel half (g4 & 1) of block (g4 >> 1); warp quadrant q owns
    // channels 32q..32q+31 (TMEM lanes), one channel per thread =====
    const int q = warp & 3, g4 = (warp - 2) >> 2, blk = g4 >> 1, phalf = g4 & 1;
    const int ch = q * 32 + lane;
    uint32_t tph = 0;
    bool ok = true;
    int di = 0;
    for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x) {
      const int n_tile = tile / per_n, rem = tile - n_tile * per_n;
      const int tw = rem % p.tiles_w, th = (rem / p.tiles_w) % p.tiles_h, n0 = rem / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.reg_w + blk * 8, h0 = th * p.reg_h;
      const int gcol = n_tile * C3_BN + ch;
      float badd = 0.f;
      if (p.bias) badd += p.bias[gcol];
      if (p.bias2) badd += p.bias2[gcol];
      if (p.rowbias) badd += p.rowbias[(int64_t)n0 * p.ld_rowbias + gcol];
      ok = mbar_wait(tmem_full, tph, p.error_flag, 3);
      tph ^= 1;
      if (!ok) break;
      tc_fence_after();
      const bool stamp = !GNS && p.dbg && warp == 2 && lane == 0 && di < 60;
      if (stamp) p.dbg[(size_t)blockIdx.x * 64 + di + 2] = clock64();   // [4i+2]: accumulators of tile i complete
      if (!p.out_f32 && !p.residual) {
        // ---- fp16 output: TMEM fragments -> f16x2 -> stmatrix.trans into a warp-private [16 px][32 ch] tile -> 16-byte
        // read-back -> st.global.v4 (every warp store = 8 pixels x 64 contiguous bytes)
        const int fr = lane >> 2;                 // fragment row: channels fr, fr+8 (+16, +24 for the upper half-quadrant)
        float bq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = n_tile * C3_BN + q * 32 + fr + 8 * i;
          float b = 0.f;
          if (p.bias) b += p.bias[c];
          if (p.bias2) b += p.bias2[c];
          if (p.rowbias) b += p.rowbias[(int64_t)n0 * p.ld_rowbias + c];
          bq[i] = b;
        }
        const uint32_t stage = smem_u32(smem_epi + (warp - 2) * 16 * C3P_EPI_PITCH);
        // stmatrix row address of this thread: matrix lane>>3 = {ch 0-7 | ch 8-15} x {px 0-7 | px 8-15}, row lane&7 = pixel
        const uint32_t st_addr = stage + (uint32_t)(((lane & 7) + ((lane >> 4) & 1) * 8) * C3P_EPI_PITCH + ((lane >> 3) & 1) * 16);
        const uint32_t rd_addr = stage + (uint32_t)((lane >> 2) * C3P_EPI_PITCH + (lane & 3) * 16);
        const float sc = p.scale;
        __half* ybase = reinterpret_cast<__half*>(p.y) + n_tile * C3_BN + q * 32 + (lane & 3) * 8;
        // software pipeline: the TMEM loads of chunk c+1 are in flight while chunk c goes through shared memory
        uint32_t va[8], vb[8];
        float gs1[4] = {0.f, 0.f, 0.f, 0.f}, gs2[4] = {0.f, 0.f, 0.f, 0.f};
        {
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + blk * 256 + phalf * 128;
          tmem_ld_16x256b_x2_nowait(ta, va);
          tmem_ld_16x256b_x2_nowait(ta + (16u << 16), vb);
        }
#pragma unroll 1
        for (int j0 = phalf * 128; j0 < phalf * 128 + 128; j0 += 16) {
          tmem_wait_ld();
          uint32_t m[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            m[i] = pack_f16x2((__uint_as_float(va[2 * i]) + bq[i & 1]) * sc, (__uint_as_float(va[2 * i + 1]) + bq[i & 1]) * sc);
            m[4 + i] = pack_f16x2((__uint_as_float(vb[2 * i]) + bq[2 + (i & 1)]) * sc, (__uint_as_float(vb[2 * i + 1]) + bq[2 + (i & 1)]) * sc);
          }
          if (GNS) {   // statistics of the ROUNDED outputs: exactly what the consumer will read back
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&m[i]));
              const int a = (i & 1) + ((i >> 2) << 1);   // channel fr + 8 a of this warp's quadrant
              gs1[a] += f.x + f.y;
              gs2[a] = fmaf(f.x, f.x, fmaf(f.y, f.y, gs2[a]));
            }
          }
          if (j0 + 16 < phalf * 128 + 128) {
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + blk * 256 + j0 + 16;
            tmem_ld_16x256b_x2_nowait(ta, va);
            tmem_ld_16x256b_x2_nowait(ta + (16u << 16), vb);
          }
          __syncwarp();   // the previous chunk's read-back is complete
          stmatrix_x4_trans(st_addr, m[0], m[1], m[2], m[3]);
          stmatrix_x4_trans(st_addr + 32, m[4], m[5], m[6], m[7]);
          __syncwarp();
          // chunk = image rows (j0>>3), (j0>>3)+1 of the block, 8 pixels each; this lane: pixel (lane>>2) of each row
          const int64_t m0 = ((int64_t)n0 * p.H + h0 + (j0 >> 3)) * p.W + w0 + (lane >> 2);
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            uint4 val;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                         : "r"(rd_addr + it * 8 * C3P_EPI_PITCH));
            *reinterpret_cast<uint4*>(ybase + (m0 + (int64_t)it * p.W) * p.ld_y) = val;
          }
        }
        if (GNS) {
          // the 4 lanes that share a fragment row hold the same 4 channels: fold them, one lane per row issues the adds
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            gs1[a] += __shfl_xor_sync(0xffffffffu, gs1[a], 1);
            gs1[a] += __shfl_xor_sync(0xffffffffu, gs1[a], 2);
            gs2[a] += __shfl_xor_sync(0xffffffffu, gs2[a], 1);
            gs2[a] += __shfl_xor_sync(0xffffffffu, gs2[a], 2);
          }
          if ((lane & 3) == 0) {
            float* sp = p.gn_sums + (int64_t)n0 * p.ld_sums + 2 * (n_tile * C3_BN + q * 32 + (lane >> 2));
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              atomicAdd(sp + 16 * a, gs1[a]);
              atomicAdd(sp + 16 * a + 1, gs2[a]);
            }
          }
        }
      } else {
      const int ldr = (int)p.ld_res, ldy = (int)p.ld_y;
#pragma unroll 1
      for (int j0 = phalf * 128; j0 < phalf * 128 + 128; j0 += 16) {
        uint32_t v[16];
        tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + blk * 256 + j0, v);
        // pixel j of the block = (row j >> 3, column j & 7): this chunk is 2 image rows of 8 pixels
        const int64_t m0 = ((int64_t)n0 * p.H + h0 + (j0 >> 3)) * p.W + w0;
        unsigned short res[16];
        if (p.residual) {
          const unsigned short* rp = reinterpret_cast<const unsigned short*>(p.residual) + m0 * p.ld_res + gcol;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const unsigned short* rr = rp + (int64_t)r * p.W * p.ld_res;
#pragma unroll
            for (int px = 0; px < 8; ++px) res[r * 8 + px] = rr[px * ldr];
          }
        }
        tmem_wait_ld();
        if (p.out_f32) {
          float* yp = reinterpret_cast<float*>(p.y) + m0 * p.ld_y + gcol;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            float* yr = yp + (int64_t)r * p.W * p.ld_y;
#pragma unroll
            for (int px = 0; px < 8; ++px) {
              float f = __uint_as_float(v[r * 8 + px]) + badd;
              if (p.residual) f += __half2float(__ushort_as_half(res[r * 8 + px]));
              yr[px * ldy] = f * p.scale;
            }
          }
        } else {
          __half* yp = reinterpret_cast<__half*>(p.y) + m0 * p.ld_y + gcol;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            __half* yr = yp + (int64_t)r * p.W * p.ld_y;
#pragma unroll
            for (int px = 0; px < 8; ++px) {
              float f = __uint_as_float(v[r * 8 + px]) + badd;
              if (p.residual) f += __half2float(__ushort_as_half(res[r * 8 + px]));
              yr[px * ldy] = __float2half_rn(f * p.scale);
            }
          }
        }
      }
      }  // direct-store path (fp32 output / explicit residual)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      if (stamp) p.dbg[(size_t)blockIdx.x * 64 + di + 3] = clock64();   // [4i+3]: this warp's share of tile i stored
      di += 4;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =============================================================================================================
// Cluster variant of the persistent kernel ("conv3c", the default for H % 32 == 0): the epilogue of tile i runs UNDER
// the MMAs of tile i+1.
//   * Tile = 128 channels x 256 pixels (one 8 px x 32 row strip): ONE 256-column accumulator, so the 512 TMEM columns
//     hold two tiles and the MMA warp ping-pongs between them while the 16 epilogue warps drain the other one.
//   * A 256-pixel tile re-fetches every weight tile twice as often as the 512-pixel tile of umma_conv3p_kernel (41
//     B/clk/SM of L2->SM traffic, the measured chip limit).  The CS CTAs of a thread-block cluster therefore walk CS
//     neighbouring pixel tiles of the SAME channel slab in lockstep and share the weight ring: CTA r fetches rows
//     [r*128/CS, (r+1)*128/CS) of each tap's 128 x 64 weight tile and TMA-MULTICASTS them into the ring slot of every
//     CTA of the cluster; a slot is released by tcgen05.commit multicast to all CTAs' b_empty barriers (count CS), so
//     no CTA overwrites a slot a peer's tensor core still reads.  L2->SM traffic per CTA: 43.5 KB halo + 147/CS KB of
//     weights per 4.6 kcycles of MMA issue.
//   * Everything else (orientation D[cout, pixel], taps as shifted descriptors of the resident halo tile, K-segments
//     for the fused 1x1 shortcut / identity residual, stmatrix epilogue) is umma_conv3p_kernel's.
// =============================================================================================================
constexpr int C3C_A_STAGE_BYTES = 43 * 1024;          // >= 10*34*128 = 43,520
constexpr int C3C_BSTAGES = 6;
constexpr int C3C_B_OFFSET = C3_ASTAGES * C3C_A_STAGE_BYTES;
constexpr int C3C_EPI_OFFSET = C3C_B_OFFSET + C3C_BSTAGES * C3_B_STAGE_BYTES;
constexpr int C3C_BAR_OFFSET = C3C_EPI_OFFSET + C3P_EPI_BYTES;
constexpr int C3C_SMEM = C3C_BAR_OFFSET + (2 * C3_ASTAGES + 2 * C3C_BSTAGES + 4) * 8 + 16 + 1024;
static_assert(C3C_SMEM <= 232448, "conv3c shared memory");
static_assert(C3C_A_STAGE_BYTES >= 10 * 34 * 128 && C3C_A_STAGE_BYTES % 1024 == 0, "conv3c halo stage");

template <bool A_MN, int CS>
__global__ void __launch_bounds__(576, 1) umma_conv3c_kernel(const __grid_constant__ CUtensorMap tmX0,
                                                             const __grid_constant__ CUtensorMap tmX1,
                                                             const __grid_constant__ CUtensorMap tmW0,
                                                             const __grid_constant__ CUtensorMap tmW1,
                                                             const Conv3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem + C3C_B_OFFSET;
  uint8_t* smem_epi = smem + C3C_EPI_OFFSET;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C3C_BAR_OFFSET);
  uint64_t* a_empty = a_full + C3_ASTAGES;
  uint64_t* b_full = a_empty + C3_ASTAGES;
  uint64_t* b_empty = b_full + C3C_BSTAGES;
  uint64_t* tmem_full = b_empty + C3C_BSTAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x / CS, nclusters = gridDim.x / CS;
  const int per_n = p.tiles_w * p.tiles_h * p.NB;  // pixel tiles per 128-channel slab (a multiple of CS)
  const int ngroups = per_n * (p.N / C3_BN) / CS;  // a group = CS neighbouring tiles of one slab, one per CTA
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
  constexpr int kSliceRows = 128 / CS;             // weight-tile rows (K-major) / k rows of one 64-channel half (MN-major) per CTA

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX0);
    prefetch_tmap(&tmX1);
    prefetch_tmap(&tmW0);
    prefetch_tmap(&tmW1);
    for (int s = 0; s < C3_ASTAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C3C_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], CS); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 16); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // every CTA's barriers exist before a peer multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // prologue above overlaps the previous kernel; global memory is touched only below

  if (warp == 0) {
    // ===== TMA producer: halo tiles of this CTA (unicast) + this CTA's slice of every weight tile (multicast) =====
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true;
      for (int g = cluster_id; g < ngroups && ok; g += nclusters) {
        const int tile = g * CS + rank;
        const int n_tile = tile / per_n, rem = tile - n_tile * per_n;
        const int tw = rem % p.tiles_w, th = (rem / p.tiles_w) % p.tiles_h, n0 = rem / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.reg_w, h0 = th * p.reg_h;
        for (int seg = 0; seg < 2 && ok; ++seg) {
          const int nkb = seg ? p.nkb2 : p.nkb;
          const int ntap = seg ? 1 : 9;
          const CUtensorMap* mapX = seg ? &tmX1 : &tmX0;
          const CUtensorMap* mapW = seg ? &tmW1 : &tmW0;
          for (int kb = 0; kb < nkb && ok; ++kb) {
            ok = mbar_wait(&a_empty[as], aph ^ 1, p.error_flag, 1);
            if (!ok) break;
            mbar_expect_tx(&a_full[as], p.a_bytes);
            tma_load_4d(mapX, &a_full[as], smem + as * C3C_A_STAGE_BYTES, kb * C3_BK, w0 - 1, h0 - 1, n0);
            for (int t = 0; t < ntap; ++t) {
              ok = mbar_wait(&b_empty[bs], bph ^ 1, p.error_flag, 1);   // released by ALL CTAs of the cluster
              if (!ok) break;
              uint8_t* sw = smem_w + bs * C3_B_STAGE_BYTES;
              mbar_expect_tx(&b_full[bs], C3_B_STAGE_BYTES);            // the CS slices that land in THIS CTA's slot
              if (!A_MN) {
                tma_load_3d_mc(mapW, &b_full[bs], sw + rank * kSliceRows * 128, kb * C3_BK, n_tile * C3_BN + rank * kSliceRows, t, kMask);
              } else {
                // MN-major tile = two [64 k][64 n] halves; CTA r owns k rows [part*kSliceRows, +kSliceRows) of half `half`
                const int half = rank / (CS / 2), part = rank % (CS / 2);
                tma_load_3d_mc(mapW, &b_full[bs], sw + half * (64 * C3_BK * 2) + part * kSliceRows * 128,
                               n_tile * C3_BN + half * 64, kb * C3_BK + part * kSliceRows, t, kMask);
              }
              if (++bs == C3C_BSTAGES) { bs = 0; bph ^= 1; }
            }
            if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator (tile & 1) =====
    if (lane == 0) {
      int as = 0, bs = 0, acc = 0, di = 0;
      uint32_t aph = 0, bph = 0, tph = 0;   // tph: bit `acc` = phase of accumulator acc
      bool ok = true;
      for (int g = cluster_id; g < ngroups && ok; g += nclusters) {
        ok = mbar_wait(&tmem_empty[acc], ((tph >> acc) & 1u) ^ 1u, p.error_flag, 4);  // drained by the epilogue two tiles ago
        if (!ok) break;
        tc_fence_after();
        if (p.dbg && di < 60) p.dbg[(size_t)blockIdx.x * 64 + di] = clock64();   // [4i]: tile i may start
        const uint32_t tacc = tmem_base + acc * 256;
        bool first = true;
        for (int seg = 0; seg < 2 && ok; ++seg) {
          const int nkb = seg ? p.nkb2 : p.nkb;
          const int ntap = seg ? 1 : 9;
          for (int kb = 0; kb < nkb && ok; ++kb) {
            ok = mbar_wait(&a_full[as], aph, p.error_flag, 2);
            if (!ok) break;
            const uint32_t sx = smem_u32(smem + as * C3C_A_STAGE_BYTES);
            for (int t = 0; t < ntap; ++t) {
              ok = mbar_wait(&b_full[bs], bph, p.error_flag, 2);
              if (!ok) break;
              tc_fence_after();
              int dy = seg ? 0 : t / 3 - 1, dx = seg ? 0 : t % 3 - 1;
              if (p.flip) { dy = -dy; dx = -dx; }
              const uint32_t sw = smem_u32(smem_w + bs * C3_B_STAGE_BYTES);
              const uint32_t x0 = sx + (uint32_t)((dy + 1) * p.pitch + (dx + 1)) * 128u;
#pragma unroll
              for (int k = 0; k < C3_BK / 16; ++k) {
                const uint64_t wd = A_MN ? make_desc(sw + k * 2048, 512, 64) : make_desc(sw + k * 32, 1, 64);
                const uint64_t xd = make_desc(x0 + k * 32, 1, p.a_sbo);
                umma_f16(tacc, wd, xd, p.idesc, (first && k == 0) ? 0u : 1u);
              }
              first = false;
              umma_commit_mc(&b_empty[bs], kMask);   // this CTA is done with the slot: tell every producer of the cluster
              if (++bs == C3C_BSTAGES) { bs = 0; bph ^= 1; }
            }
            if (ok) umma_commit(&a_empty[as]);
            if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
          }
        }
        if (ok) umma_commit(&tmem_full[acc]);
        if (p.dbg && di < 60) p.dbg[(size_t)blockIdx.x * 64 + di + 1] = clock64();  // [4i+1]: last MMA of tile i issued
        di += 4;
        tph ^= 1u << acc;
        acc ^= 1;
      }
    }
  } else {
    // ===== epilogue: 16 warps; warp quadrant q owns channels 32q..32q+31 (TMEM lanes), group g4 the 64 pixel columns
    // [64 g4, 64 g4 + 64) = image rows 8 g4 .. 8 g4 + 7 of the strip =====
    const int q = warp & 3, g4 = (warp - 2) >> 2;
    const int ch = q * 32 + lane;
    uint32_t tph = 0;
    int acc = 0, di = 0;
    bool ok = true;
    for (int g = cluster_id; g < ngroups && ok; g += nclusters) {
      const bool stamp = p.dbg && warp == 2 && lane == 0 && di < 60;
      const int tile = g * CS + rank;
      const int n_tile = tile / per_n, rem = tile - n_tile * per_n;
      const int tw = rem % p.tiles_w, th = (rem / p.tiles_w) % p.tiles_h, n0 = rem / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.reg_w, h0 = th * p.reg_h;
      const int gcol = n_tile * C3_BN + ch;
      float badd = 0.f;
      if (p.bias) badd += p.bias[gcol];
      if (p.bias2) badd += p.bias2[gcol];
      if (p.rowbias) badd += p.rowbias[(int64_t)n0 * p.ld_rowbias + gcol];
      float bq[4];
      {
        const int fr = lane >> 2;                 // fragment row: channels fr, fr+8 (+16, +24 for the upper half-quadrant)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = n_tile * C3_BN + q * 32 + fr + 8 * i;
          float b = 0.f;
          if (p.bias) b += p.bias[c];
          if (p.bias2) b += p.bias2[c];
          if (p.rowbias) b += p.rowbias[(int64_t)n0 * p.ld_rowbias + c];
          bq[i] = b;
        }
      }
      ok = mbar_wait(&tmem_full[acc], (tph >> acc) & 1u, p.error_flag, 3);
      tph ^= 1u << acc;
      if (!ok) break;
      tc_fence_after();
      if (stamp) p.dbg[(size_t)blockIdx.x * 64 + di + 2] = clock64();   // [4i+2]: accumulator of tile i complete
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      if (!p.out_f32 && !p.residual) {
        // fp16 output: TMEM fragments -> f16x2 -> stmatrix.trans into a warp-private [16 px][32 ch] tile -> 16-byte
        // read-back -> st.global.v4 (every warp store = 8 pixels x 64 contiguous bytes)
        const uint32_t stage = smem_u32(smem_epi + (warp - 2) * 16 * C3P_EPI_PITCH);
        const uint32_t st_addr = stage + (uint32_t)(((lane & 7) + ((lane >> 4) & 1) * 8) * C3P_EPI_PITCH + ((lane >> 3) & 1) * 16);
        const uint32_t rd_addr = stage + (uint32_t)((lane >> 2) * C3P_EPI_PITCH + (lane & 3) * 16);
        const float sc = p.scale;
        __half* ybase = reinterpret_cast<__half*>(p.y) + n_tile * C3_BN + q * 32 + (lane & 3) * 8;
        uint32_t va[8], vb[8];
        float gs1[4] = {0.f, 0.f, 0.f, 0.f}, gs2[4] = {0.f, 0.f, 0.f, 0.f};
        tmem_ld_16x256b_x2_nowait(tacc + g4 * 64, va);
        tmem_ld_16x256b_x2_nowait(tacc + (16u << 16) + g4 * 64, vb);
#pragma unroll 1
        for (int j0 = g4 * 64; j0 < g4 * 64 + 64; j0 += 16) {
          tmem_wait_ld();
          uint32_t m[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            m[i] = pack_f16x2((__uint_as_float(va[2 * i]) + bq[i & 1]) * sc, (__uint_as_float(va[2 * i + 1]) + bq[i & 1]) * sc);
            m[4 + i] = pack_f16x2((__uint_as_float(vb[2 * i]) + bq[2 + (i & 1)]) * sc, (__uint_as_float(vb[2 * i + 1]) + bq[2 + (i & 1)]) * sc);
          }
          if (p.gn_sums) {   // statistics of the ROUNDED outputs: exactly what the consumer will read back
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&m[i]));
              const int a = (i & 1) + ((i >> 2) << 1);   // channel fr + 8 a of this warp's quadrant
              gs1[a] += f.x + f.y;
              gs2[a] = fmaf(f.x, f.x, fmaf(f.y, f.y, gs2[a]));
            }
          }
          if (j0 + 16 < g4 * 64 + 64) {
            tmem_ld_16x256b_x2_nowait(tacc + j0 + 16, va);
            tmem_ld_16x256b_x2_nowait(tacc + (16u << 16) + j0 + 16, vb);
          }
          __syncwarp();   // the previous chunk's read-back is complete
          stmatrix_x4_trans(st_addr, m[0], m[1], m[2], m[3]);
          stmatrix_x4_trans(st_addr + 32, m[4], m[5], m[6], m[7]);
          __syncwarp();
          // chunk = image rows (j0>>3), (j0>>3)+1 of the strip, 8 pixels each; this lane: pixel (lane>>2) of each row
          const int64_t m0 = ((int64_t)n0 * p.H + h0 + (j0 >> 3)) * p.W + w0 + (lane >> 2);
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            uint4 val;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                         : "r"(rd_addr + it * 8 * C3P_EPI_PITCH));
            *reinterpret_cast<uint4*>(ybase + (m0 + (int64_t)it * p.W) * p.ld_y) = val;
          }
        }
        if (p.gn_sums) {
          // the 4 lanes that share a fragment row hold the same 4 channels: fold them, one lane per row issues the adds
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            gs1[a] += __shfl_xor_sync(0xffffffffu, gs1[a], 1);
            gs1[a] += __shfl_xor_sync(0xffffffffu, gs1[a], 2);
            gs2[a] += __shfl_xor_sync(0xffffffffu, gs2[a], 1);
            gs2[a] += __shfl_xor_sync(0xffffffffu, gs2[a], 2);
          }
          if ((lane & 3) == 0) {
            float* sp = p.gn_sums + (int64_t)n0 * p.ld_sums + 2 * (n_tile * C3_BN + q * 32 + (lane >> 2));
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              atomicAdd(sp + 16 * a, gs1[a]);
              atomicAdd(sp + 16 * a + 1, gs2[a]);
            }
          }
        }
      } else {
        const int ldr = (int)p.ld_res, ldy = (int)p.ld_y;
#pragma unroll 1
        for (int j0 = g4 * 64; j0 < g4 * 64 + 64; j0 += 16) {
          uint32_t v[16];
          tmem_ld16_nowait(tacc + j0, v);
          const int64_t m0 = ((int64_t)n0 * p.H + h0 + (j0 >> 3)) * p.W + w0;
          unsigned short res[16];
          if (p.residual) {
            const unsigned short* rp = reinterpret_cast<const unsigned short*>(p.residual) + m0 * p.ld_res + gcol;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const unsigned short* rr = rp + (int64_t)r * p.W * p.ld_res;
#pragma unroll
              for (int px = 0; px < 8; ++px) res[r * 8 + px] = rr[px * ldr];
            }
          }
          tmem_wait_ld();
#pragma unroll
          for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int px = 0; px < 8; ++px) {
              float f = __uint_as_float(v[r * 8 + px]) + badd;
              if (p.residual) f += __half2float(__ushort_as_half(res[r * 8 + px]));
              f *= p.scale;
              const int64_t off = (m0 + (int64_t)r * p.W + px) * ldy + gcol;
              if (p.out_f32) reinterpret_cast<float*>(p.y)[off] = f;
              else reinterpret_cast<__half*>(p.y)[off] = __float2half_rn(f);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (stamp) p.dbg[(size_t)blockIdx.x * 64 + di + 3] = clock64();   // [4i+3]: this warp's share of tile i stored
      di += 4;
      acc ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA retires while a peer may still multicast into its ring or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, int cs,
                                             cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = getenv("BD_NO_PDL") ? 1 : 2;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// cluster size of the conv3c kernel for this launch (0: use the non-cluster persistent kernel).
// Measured on B200 (scripts/ab_conv3c.py): a 256-pixel tile makes shared memory the bound -- 96 B/clk of MMA operand reads
// + 41 B/clk of TMA fills + the overlapped epilogue's stmatrix / read-back against the 128 B/clk port: 13.9 kcycles of MMA
// issue per 256-pixel tile (ideal 9.2 k) against 19.4 k per 512-pixel tile for conv3p -- so hiding the epilogue buys
// nothing at 128 -> 128 @ 32x32, B = 128 (38.0 vs 39.5 us).  What the smaller tile does buy is a finer walk: the cluster
// kernel wins whenever 512-pixel tiles fill the last round over the SMs badly and 256-pixel tiles fill it better
// (CelebA-HQ shapes: 4 x 256x256 128->128 -9 %, 4 x 64x64 256->256 -30 %; 256->256 @ 32x32 B = 128 -3 %).  Cluster size 4
// never beat 2 (weights are not the binding traffic once multicast halves them).
// BD_CONV3C=0 disables, BD_CONV3C=2|4 forces that cluster size, unset = the round-fill heuristic with clusters of 2.
int conv3c_cluster_size(int H, int W, int NB, int N) {
  const char* e = getenv("BD_CONV3C");
  const int forced = e ? atoi(e) : -1;
  if (forced == 0) return 0;
  if (H % 32 || W % 8) return 0;
  const long long per_n = (long long)(W / 8) * (H / 32) * NB;
  if (forced == 2 || forced == 4) return per_n % forced == 0 ? forced : 0;
  if (per_n % 2 || W % 16) return 0;
  const long long sms = num_sms(), slabs = N / C3_BN;
  const long long n256 = per_n * slabs, n512 = n256 / 2;
  const double fill512 = (double)n512 / (double)(((n512 + sms - 1) / sms) * sms);
  const long long cl = sms / 2, g256 = n256 / 2;   // groups of 2 tiles walk over sms/2 clusters
  const double fill256 = (double)g256 / (double)(((g256 + cl - 1) / cl) * cl);
  return fill256 > fill512 + 0.05 ? 2 : 0;
}

int conv3c_launch(const CUtensorMap& mx0, const CUtensorMap& mx1, const CUtensorMap& mw0, const CUtensorMap& mw1,
                  Conv3Params p, bool a_mn, int cs, cudaStream_t st) {
  p.idesc = (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int ngroups = p.tiles_w * p.tiles_h * p.NB * (p.N / C3_BN) / cs;
  int nclusters = num_sms() / cs;
  if (nclusters > ngroups) nclusters = ngroups;
  void (*k)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, Conv3Params) =
      cs == 4 ? (a_mn ? umma_conv3c_kernel<true, 4> : umma_conv3c_kernel<false, 4>)
              : (a_mn ? umma_conv3c_kernel<true, 2> : umma_conv3c_kernel<false, 2>);
  // The tile walk is static (group g -> cluster g % nclusters): every cluster must be resident at once.  A GPC whose SM
  // count is not a multiple of cs cannot host a cluster on its last SMs, so ask the driver how many fit.
  static int max_clusters[4] = {0, 0, 0, 0};
  const int ai = (cs == 4 ? 2 : 0) + (a_mn ? 1 : 0);
  if (!max_clusters[ai]) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C3C_SMEM);
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3((num_sms() / cs) * cs);
    q.blockDim = dim3(576);
    q.dynamicSmemBytes = C3C_SMEM;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = cs; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k, &q) != cudaSuccess || n < 1) { cudaGetLastError(); n = num_sms() / cs; }
    max_clusters[ai] = (int)env_u32("BD_CONV3C_CLUSTERS", (uint32_t)n);
  }
  if (nclusters > max_clusters[ai]) nclusters = max_clusters[ai];
  cudaError_t e = launch_pdl_cluster(k, dim3(nclusters * cs), dim3(576), C3C_SMEM, cs, st, mx0, mx1, mw0, mw1, p);
  if (e != cudaSuccess) { set_error("conv3c: cluster launch failed: %s", cudaGetErrorString(e)); return BD_ERR_CUDA; }
  count_launch(1);
  return BD_OK;
}

int conv3t_launch(const CUtensorMap& mx0, const CUtensorMap& mx1, const CUtensorMap& mw0, const CUtensorMap& mw1,
                  Conv3Params p, bool a_mn, dim3 grid, cudaStream_t st) {
  p.idesc = (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (!getenv("BD_NO_CONV3P")) {
    // persistent kernel: one CTA per SM over all (channel slab, pixel tile) pairs, channel slab slowest so the CTAs
    // that run together share their weight tiles in L2
    const int total = (int)(grid.x * grid.y);
    const int ctas = total < num_sms() ? total : num_sms();
    static bool pattr[2] = {false, false};
    if (a_mn) {
      if (!pattr[1]) { cudaFuncSetAttribute(umma_conv3p_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3P_SMEM); pattr[1] = true; }
      launch_pdl(umma_conv3p_kernel<true>, dim3(ctas), dim3(576), C3P_SMEM, st, mx0, mx1, mw0, mw1, p);
    } else {
      if (p.gn_sums) {
        static bool gattr = false;
        if (!gattr) { cudaFuncSetAttribute(umma_conv3p_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3P_SMEM); gattr = true; }
        launch_pdl(umma_conv3p_kernel<false, true>, dim3(ctas), dim3(576), C3P_SMEM, st, mx0, mx1, mw0, mw1, p);
      } else {
        if (!pattr[0]) { cudaFuncSetAttribute(umma_conv3p_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3P_SMEM); pattr[0] = true; }
        launch_pdl(umma_conv3p_kernel<false>, dim3(ctas), dim3(576), C3P_SMEM, st, mx0, mx1, mw0, mw1, p);
      }
    }
    count_launch(1);
    return BD_OK;
  }
  static bool attr_set[2] = {false, false};
  if (a_mn) {
    if (!attr_set[1]) { cudaFuncSetAttribute(umma_conv3t_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM); attr_set[1] = true; }
    launch_pdl(umma_conv3t_kernel<true>, grid, dim3(576), C3_SMEM, st, mx0, mx1, mw0, mw1, p);
  } else {
    if (!attr_set[0]) { cudaFuncSetAttribute(umma_conv3t_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM); attr_set[0] = true; }
    launch_pdl(umma_conv3t_kernel<false>, grid, dim3(576), C3_SMEM, st, mx0, mx1, mw0, mw1, p);
  }
  count_launch(1);
  return BD_OK;
}

// =============================================================================================================
// 16 x 16 images (the UNet's second resolution, 37 % of its FLOPs): one CTA = ONE image x 256 output channels.
//   D[pixel, cout]: A = 128 pixels (8 px x 16 rows, a shifted view of the image's halo tile), B = a 256 x 64 weight
//   tile, so the MMAs are 128 x 256 x 16 -- 96 B/clk of shared-memory operand reads, which sustains the full issue
//   rate; the 128 x 128 MMAs of umma_conv3_kernel need 128 B/clk and measured 77-110 cycles instead of 64.
//   Two blocks (left / right half of the image) x 256 fp32 columns = all 512 TMEM columns.  The input image is
//   fetched once for all 256 channels (the 128-channel kernel fetched it once per channel slab).
//   Epilogue: TMEM fragments (tcgen05.ld 16x256b) -> +bias -> f16x2 -> stmatrix into a warp-private [32 px][32 ch]
//   tile -> 16-byte read-back -> st.global.v4 (8 pixels x 64 contiguous bytes per warp store).
// Used for forward (K-major weights) and dgrad (MN-major view of the same weights, taps flipped); the 1x1 shortcut /
// identity-residual K segment works as in the other kernels.
// =============================================================================================================
constexpr int C3W_BN = 256;
constexpr int C3W_A_STAGE_BYTES = 41 * 1024;            // >= 18*18*128 = 41,472
constexpr int C3W_B_STAGE_BYTES = C3W_BN * C3_BK * 2;   // 32 KB
constexpr int C3W_EPI_PITCH = 80;
constexpr int C3W_EPI_WARP_BYTES = 32 * C3W_EPI_PITCH;  // [32 px][32 ch] f16, padded rows
constexpr int C3W_B_OFFSET = C3_ASTAGES * C3W_A_STAGE_BYTES;
constexpr int C3W_EPI_OFFSET = C3W_B_OFFSET + C3_BSTAGES * C3W_B_STAGE_BYTES;
constexpr int C3W_BIAS_OFFSET = C3W_EPI_OFFSET + 16 * C3W_EPI_WARP_BYTES;
constexpr int C3W_BAR_OFFSET = C3W_BIAS_OFFSET + C3W_BN * 4;
constexpr int C3W_SMEM = C3W_BAR_OFFSET + (2 * C3_ASTAGES + 2 * C3_BSTAGES + 1) * 8 + 16 + 1024;
static_assert(C3W_SMEM <= 232448, "conv3w shared memory");

__device__ __forceinline__ void tmem_ld_16x256b_x4_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void stmatrix_x4(uint32_t smem_row_addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};"
               ::"r"(smem_row_addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

template <bool B_MN>
__global__ void __launch_bounds__(576, 1) umma_conv3w_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                             const __grid_constant__ CUtensorMap tmA1,
                                                             const __grid_constant__ CUtensorMap tmB0,
                                                             const __grid_constant__ CUtensorMap tmB1,
                                                             const Conv3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + C3W_B_OFFSET;
  uint8_t* smem_epi = smem + C3W_EPI_OFFSET;
  float* sbias = reinterpret_cast<float*>(smem + C3W_BIAS_OFFSET);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C3W_BAR_OFFSET);
  uint64_t* a_empty = a_full + C3_ASTAGES;
  uint64_t* b_full = a_empty + C3_ASTAGES;
  uint64_t* b_empty = b_full + C3_BSTAGES;
  uint64_t* tmem_full = b_empty + C3_BSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x, n_tile = blockIdx.y;   // image, 256-channel slab

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB0);
    prefetch_tmap(&tmB1);
    for (int s = 0; s < C3_ASTAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C3_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + C3W_BN) {   // bias + bias2 + per-image row bias of this slab, once
    pdl_wait();                                           // (reads global memory: not part of the overlappable prologue)
    const int c = n_tile * C3W_BN + (threadIdx.x - 64);
    float b = 0.f;
    if (p.bias) b += p.bias[c];
    if (p.bias2) b += p.bias2[c];
    if (p.rowbias) b += p.rowbias[(int64_t)n0 * p.ld_rowbias + c];
    sbias[threadIdx.x - 64] = b;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // prologue above overlaps the previous kernel; global memory is touched only below

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true;
      for (int seg = 0; seg < 2 && ok; ++seg) {
        const int nkb = seg ? p.nkb2 : p.nkb;
        const int ntap = seg ? 1 : 9;
        const CUtensorMap* mapA = seg ? &tmA1 : &tmA0;
        const CUtensorMap* mapB = seg ? &tmB1 : &tmB0;
        for (int kb = 0; kb < nkb && ok; ++kb) {
          ok = mbar_wait(&a_empty[as], aph ^ 1, p.error_flag, 1);
          if (!ok) break;
          mbar_expect_tx(&a_full[as], p.a_bytes);
          tma_load_4d(mapA, &a_full[as], smem + as * C3W_A_STAGE_BYTES, kb * C3_BK, -1, -1, n0);
          for (int t = 0; t < ntap; ++t) {
            ok = mbar_wait(&b_empty[bs], bph ^ 1, p.error_flag, 1);
            if (!ok) break;
            uint8_t* sb = smem_b + bs * C3W_B_STAGE_BYTES;
            mbar_expect_tx(&b_full[bs], C3W_B_STAGE_BYTES);
            if (!B_MN) {
              tma_load_3d(mapB, &b_full[bs], sb, kb * C3_BK, n_tile * C3W_BN, t);                    // box {64 k, 256 n}
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)                                                            // 4 boxes {64 n, 64 k}
                tma_load_3d(mapB, &b_full[bs], sb + j * 64 * C3_BK * 2, n_tile * C3W_BN + j * 64, kb * C3_BK, t);
            }
            if (++bs == C3_BSTAGES) { bs = 0; bph ^= 1; }
          }
          if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true, first = true;
      for (int seg = 0; seg < 2 && ok; ++seg) {
        const int nkb = seg ? p.nkb2 : p.nkb;
        const int ntap = seg ? 1 : 9;
        for (int kb = 0; kb < nkb && ok; ++kb) {
          ok = mbar_wait(&a_full[as], aph, p.error_flag, 2);
          if (!ok) break;
          const uint32_t sa = smem_u32(smem + as * C3W_A_STAGE_BYTES);
          for (int t = 0; t < ntap; ++t) {
            ok = mbar_wait(&b_full[bs], bph, p.error_flag, 2);
            if (!ok) break;
            tc_fence_after();
            int dy = seg ? 0 : t / 3 - 1, dx = seg ? 0 : t % 3 - 1;
            if (p.flip) { dy = -dy; dx = -dx; }
            const uint32_t sb = smem_u32(smem_b + bs * C3W_B_STAGE_BYTES);
            const int tap_row = (dy + 1) * p.pitch + (dx + 1);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
              const uint32_t a0 = sa + (uint32_t)(mb * 8 + tap_row) * 128u;   // block mb = image columns 8mb .. 8mb+7
#pragma unroll
              for (int k = 0; k < C3_BK / 16; ++k) {
                const uint64_t ad = make_desc(a0 + k * 32, 1, p.a_sbo);
                const uint64_t bd = B_MN ? make_desc(sb + k * 2048, 512, 64) : make_desc(sb + k * 32, 1, 64);
                umma_f16(tmem_base + mb * C3W_BN, ad, bd, p.idesc, (first && k == 0) ? 0u : 1u);
              }
            }
            first = false;
            umma_commit(&b_empty[bs]);
            if (++bs == C3_BSTAGES) { bs = 0; bph ^= 1; }
          }
          if (ok) umma_commit(&a_empty[as]);
          if (++as == C3_ASTAGES) { as = 0; aph ^= 1; }
        }
      }
      if (ok) umma_commit(tmem_full);
    }
  } else {
    // ===== epilogue: warp quadrant q = TMEM lanes (pixels) 32q.., column group cg = (warp-2)>>2: block cg>>1, channel
    // half cg&1 (128 channels = 4 chunks of 32) =====
    const int q = warp & 3, cg = (warp - 2) >> 2, blk = cg >> 1, chalf = cg & 1;
    const bool ok = mbar_wait(tmem_full, 0, p.error_flag, 3);
    tc_fence_after();
    if (ok) {
      const uint32_t stage = smem_u32(smem_epi + (warp - 2) * C3W_EPI_WARP_BYTES);
      // stmatrix (non-transposed): matrix i = lane>>3 holds pixels (i&1)*8.. of the 16-lane half, channels (i>>1)*8..
      const uint32_t st_row = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * C3W_EPI_PITCH + ((lane >> 4) & 1) * 16);
      const uint32_t rd_addr = stage + (uint32_t)((lane >> 2) * C3W_EPI_PITCH + (lane & 3) * 16);
      const float sc = p.scale;
      const int fc = 2 * (lane & 3);   // fragment columns fc, fc+1 (+8j)
      // pixel of TMEM lane r of block blk: row r>>3, column blk*8 + (r&7); read-back step `it` of this warp = image row 4q+it
      __half* ybase = reinterpret_cast<__half*>(p.y) +
                      (((int64_t)n0 * p.H + 4 * q) * p.W + blk * 8 + (lane >> 2)) * p.ld_y + n_tile * C3W_BN + chalf * 128 +
                      (lane & 3) * 8;
      float* ybase32 = reinterpret_cast<float*>(p.y) +
                       (((int64_t)n0 * p.H + 4 * q) * p.W + blk * 8 + (lane >> 2)) * p.ld_y + n_tile * C3W_BN + chalf * 128 +
                       (lane & 3) * 8;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        const int col = chalf * 128 + c0;                 // first channel of this chunk inside the 256-slab
        __syncwarp();                                     // previous chunk's read-back is complete
#pragma unroll
        for (int lh = 0; lh < 2; ++lh) {                  // lanes (pixels) 0-15, 16-31 of the quadrant
          uint32_t v[16];
          tmem_ld_16x256b_x4_nowait(tmem_base + ((uint32_t)(q * 32 + lh * 16) << 16) + blk * C3W_BN + col, v);
          tmem_wait_ld();
          uint32_t m[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {                   // channels col + 8j + fc, +1
            const float2 b = *reinterpret_cast<const float2*>(sbias + col + 8 * j + fc);
            m[2 * j] = pack_f16x2((__uint_as_float(v[4 * j]) + b.x) * sc, (__uint_as_float(v[4 * j + 1]) + b.y) * sc);
            m[2 * j + 1] = pack_f16x2((__uint_as_float(v[4 * j + 2]) + b.x) * sc, (__uint_as_float(v[4 * j + 3]) + b.y) * sc);
          }
          const uint32_t base = stage + lh * 16 * C3W_EPI_PITCH + st_row;
          stmatrix_x4(base, m[0], m[1], m[2], m[3]);        // channels +0..15
          stmatrix_x4(base + 32, m[4], m[5], m[6], m[7]);   // channels +16..31
        }
        __syncwarp();
        float gs1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gs2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int it = 0; it < 4; ++it) {                  // 8 pixels (one image row of the block) x 64 bytes per step
          uint4 val;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                       : "r"(rd_addr + it * 8 * C3W_EPI_PITCH));
          if (p.gn_sums) {   // GroupNorm statistics of the rounded outputs: this lane = 8 channels of one pixel
            const __half2* h = reinterpret_cast<const __half2*>(&val);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __half22float2(h[k]);
              gs1[2 * k] += f.x; gs1[2 * k + 1] += f.y;
              gs2[2 * k] = fmaf(f.x, f.x, gs2[2 * k]); gs2[2 * k + 1] = fmaf(f.y, f.y, gs2[2 * k + 1]);
            }
          }
          if (!p.out_f32) {
            *reinterpret_cast<uint4*>(ybase + (int64_t)it * p.W * p.ld_y + c0) = val;
          } else {
            const __half2* h = reinterpret_cast<const __half2*>(&val);
            float* yr = ybase32 + (int64_t)it * p.W * p.ld_y + c0;
            const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
            *reinterpret_cast<float4*>(yr) = make_float4(f0.x, f0.y, f1.x, f1.y);
            *reinterpret_cast<float4*>(yr + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
          }
        }
        if (p.gn_sums) {   // fold the 8 lanes (pixels) that hold the same 8 channels, lanes 0..3 issue the adds
#pragma unroll
          for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
              gs1[k] += __shfl_xor_sync(0xffffffffu, gs1[k], o);
              gs2[k] += __shfl_xor_sync(0xffffffffu, gs2[k], o);
            }
          }
          if (lane < 4) {
            float* sp = p.gn_sums + (int64_t)n0 * p.ld_sums + 2 * (n_tile * C3W_BN + col + lane * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              atomicAdd(sp + 2 * k, gs1[k]);
              atomicAdd(sp + 2 * k + 1, gs2[k]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int conv3w_launch_fwd(const Conv3Call& c, Conv3Params p, cudaStream_t st) {
  // geometry: one 16 x 16 image per CTA, halo pitch 18
  p.pitch = 18;
  p.a_bytes = 18u * 18u * 128u;
  p.a_sbo = (uint32_t)(p.pitch * 128) >> 4;
  p.idesc = (1u << 4) | ((c.b_mn ? 1u : 0u) << 16) | ((uint32_t)(C3W_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  CUtensorMap ma0, ma1, mb0, mb1;
  const uint32_t box[4] = {64, 18, 18, 1};
  {
    uint64_t dims[4] = {(uint64_t)c.Ca, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_a, (uint64_t)c.W * c.ld_a, (uint64_t)c.H * c.W * c.ld_a};
    if (!make_map(&ma0, c.a, 4, dims, str, box)) return BD_ERR_CUDA;
  }
  if (c.a2) {
    uint64_t dims[4] = {(uint64_t)c.Ca2, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_a2, (uint64_t)c.W * c.ld_a2, (uint64_t)c.H * c.W * c.ld_a2};
    if (!make_map(&ma1, c.a2, 4, dims, str, box)) return BD_ERR_CUDA;
  } else {
    ma1 = ma0;
  }
  {
    const int cols = c.b_mn ? c.N : c.Ca;
    uint64_t dims[3] = {(uint64_t)cols, (uint64_t)c.b_rows, 9};
    uint64_t str[2] = {(uint64_t)c.ld_b, (uint64_t)c.b_rows * c.ld_b};
    uint32_t bbox[3] = {64, c.b_mn ? 64u : (uint32_t)C3W_BN, 1};
    if (!make_map(&mb0, c.b, 3, dims, str, bbox)) return BD_ERR_CUDA;
  }
  if (c.a2) {
    uint64_t dims[3] = {(uint64_t)c.Ca2, (uint64_t)c.N, 1};
    uint64_t str[2] = {(uint64_t)c.ld_b2, (uint64_t)c.N * c.ld_b2};
    uint32_t bbox[3] = {64, (uint32_t)C3W_BN, 1};
    if (!make_map(&mb1, c.b2, 3, dims, str, bbox)) return BD_ERR_CUDA;
  } else {
    mb1 = mb0;
  }
  dim3 grid(c.NB, c.N / C3W_BN);
  static bool attr_set[2] = {false, false};
  if (c.b_mn) {
    if (!attr_set[1]) { cudaFuncSetAttribute(umma_conv3w_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3W_SMEM); attr_set[1] = true; }
    launch_pdl(umma_conv3w_kernel<true>, grid, dim3(576), C3W_SMEM, st, ma0, ma1, mb0, mb1, p);
  } else {
    if (!attr_set[0]) { cudaFuncSetAttribute(umma_conv3w_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3W_SMEM); attr_set[0] = true; }
    launch_pdl(umma_conv3w_kernel<false>, grid, dim3(576), C3W_SMEM, st, ma0, ma1, mb0, mb1, p);
  }
  count_launch(1);
  return BD_OK;
}

}  // namespace umma
}  // namespace bd
