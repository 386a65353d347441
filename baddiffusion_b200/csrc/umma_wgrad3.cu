// 3x3 weight gradient on tcgen05 with operand reuse ("wgrad3"):  dW[tap][co][ci] = sum_p dY[p][co] * X[p + tap][ci]
//
// A CTA owns (128 output channels) x (128 input channels) x (one filter ROW: 3 taps dx = -1, 0, +1) and a slice of the
// pixels (split-K).  Per 128-pixel k-block (8 px x 16 rows, or 8 x 8 x 2 images) ONE dY tile and ONE X tile with a
// 1-pixel column halo are fetched; the three taps of the row are three accumulators (384 TMEM columns) fed by the same
// dY tile and by descriptor-shifted views of the same X tile (both operands are MN-major: a pixel is a 128-byte
// K-row, so a column shift is a +128 B start address; the swizzle is address-based).  Shared-memory fill traffic per
// MMA drops from 8 KB (one tap per CTA, umma.cu) to 3 KB.  Partial sums leave through red.global.add.v4.f32.
#include "umma_common.cuh"

namespace bd {
namespace umma {

constexpr int W3_STAGES = 3;
constexpr int W3_A_BYTES = 2 * 128 * 128;        // dY: two 64-channel chunks x 128 pixels x 128 B
constexpr int W3_B_BYTES = 2 * 160 * 128;        // X : two 64-channel chunks x (16 rows x 10 px) x 128 B
constexpr int W3_STAGE_BYTES = W3_A_BYTES + W3_B_BYTES;   // 72 KB
constexpr int W3_BAR_OFFSET = W3_STAGES * W3_STAGE_BYTES;
constexpr int W3_SMEM = W3_BAR_OFFSET + (2 * W3_STAGES + 1) * 8 + 16 + 1024;

struct Wgrad3Params {
  int R, NI;              // k-block = 8 px x R rows x NI images, R * NI == 16
  int tiles_w, tiles_h;   // k-blocks per image (tiles_h counts groups of R rows)
  int kblocks_total;
  int n_tiles, splits;
  int Mtot, Ntot;
  uint32_t idesc;
  float* y;               // packed [9][Mtot][Ntot] fp32
  int* error_flag;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(192, 1) umma_wgrad3_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const Wgrad3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + W3_BAR_OFFSET);
  uint64_t* empty = full + W3_STAGES;
  uint64_t* tmem_full = empty + W3_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x / p.n_tiles, n_tile = blockIdx.x % p.n_tiles;
  const int dy = (int)blockIdx.y - 1;
  const int per = (p.kblocks_total + p.splits - 1) / p.splits;
  const int kb_begin = blockIdx.z * per;
  const int kb_end = min(p.kblocks_total, kb_begin + per);
  const int nk = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < W3_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
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

  if (nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (!mbar_wait(&empty[stage], phase ^ 1, p.error_flag, 1)) break;
          const int pw = (kb % p.tiles_w) * 8, ph = ((kb / p.tiles_w) % p.tiles_h) * p.R;
          const int pn = (kb / (p.tiles_w * p.tiles_h)) * p.NI;
          uint8_t* sa = smem + stage * W3_STAGE_BYTES;
          uint8_t* sb = sa + W3_A_BYTES;
          mbar_expect_tx(&full[stage], W3_STAGE_BYTES);
          tma_load_4d(&tmA, &full[stage], sa, m_tile * 128, pw, ph, pn);
          tma_load_4d(&tmA, &full[stage], sa + 128 * 128, m_tile * 128 + 64, pw, ph, pn);
          tma_load_4d(&tmB, &full[stage], sb, n_tile * 128, pw - 1, ph + dy, pn);
          tma_load_4d(&tmB, &full[stage], sb + 160 * 128, n_tile * 128 + 64, pw - 1, ph + dy, pn);
          if (++stage == W3_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int it = 0; it < nk; ++it) {
          ok = mbar_wait(&full[stage], phase, p.error_flag, 2);
          if (!ok) break;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * W3_STAGE_BYTES);
          const uint32_t sb = sa + W3_A_BYTES;
#pragma unroll
          for (int dxi = 0; dxi < 3; ++dxi) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              // K step = 16 pixels = two 8-pixel image rows: dY rows are dense (atom stride 1024 B), X rows have the
              // 10-pixel halo pitch (atom stride 1280 B) and start dxi pixels into the halo row.
              const uint64_t ad = make_desc(sa + k * 2048, (128 * 128) >> 4, 1024 >> 4);
              const uint64_t bd = make_desc(sb + (uint32_t)(2 * k * 10 + dxi) * 128u, (160 * 128) >> 4, 1280 >> 4);
              umma_f16(tmem_base + dxi * 128, ad, bd, p.idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == W3_STAGES) { stage = 0; phase ^= 1; }
        }
        if (ok) umma_commit(tmem_full);
      }
    } else {
      const int q = warp & 3;
      const int row = m_tile * 128 + q * 32 + lane;
      const bool ok = mbar_wait(tmem_full, 0, p.error_flag, 3);
      tc_fence_after();
      if (ok) {
#pragma unroll 1
        for (int dxi = 0; dxi < 3; ++dxi) {
          const int tap = (dy + 1) * 3 + dxi;
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + dxi * 128 + c0, v);
            const int col = n_tile * 128 + c0;
            if (row < p.Mtot && col < p.Ntot) {
              float* yr = p.y + ((int64_t)tap * p.Mtot + row) * p.Ntot + col;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                red_add_v4(yr + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                           __uint_as_float(v[j + 3]));
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

struct Wgrad3Call {
  const void* dy; int64_t ld_dy; int Cout;
  const void* x; int64_t ld_x; int Cin;
  int NB, H, W;
  float* dw;
};

static bool wgrad3_geometry(int H, int W, Wgrad3Params* p) {
  if (W % 8 || W < 8) return false;
  if (H % 16 == 0) { p->R = 16; p->NI = 1; }
  else if (H == 8) { p->R = 8; p->NI = 2; }
  else return false;
  p->tiles_w = W / 8;
  p->tiles_h = H / p->R;
  return true;
}

int wgrad3_supported(const Wgrad3Call& c) {
  Wgrad3Params p;
  if (getenv("BD_NO_WGRAD3")) return 0;
  if (c.Cout % 128 || c.Cin % 128) return 0;
  if (c.ld_dy % 8 || c.ld_x % 8 || ((uintptr_t)c.dy & 15) || ((uintptr_t)c.x & 15) || ((uintptr_t)c.dw & 15)) return 0;
  return wgrad3_geometry(c.H, c.W, &p) ? 1 : 0;
}

int wgrad3_launch(const Wgrad3Call& c, cudaStream_t st) {
  Wgrad3Params p;
  memset(&p, 0, sizeof(p));
  if (!wgrad3_geometry(c.H, c.W, &p)) { set_error("wgrad3: unsupported geometry %dx%d", c.H, c.W); return BD_ERR_UNSUPPORTED; }
  p.kblocks_total = p.tiles_w * p.tiles_h * ceil_div(c.NB, p.NI);
  p.Mtot = c.Cout; p.Ntot = c.Cin;
  p.n_tiles = c.Cin / 128;
  const int m_tiles = c.Cout / 128;
  const int ctas = m_tiles * p.n_tiles * 3;
  int splits = num_sms() / ctas;
  if (splits < 1) splits = 1;
  int maxs = p.kblocks_total / 4 > 0 ? p.kblocks_total / 4 : 1;   // >= 4 k-blocks per CTA
  if (splits > maxs) splits = maxs;
  p.splits = splits;
  p.idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.y = c.dw;
  p.error_flag = error_flag();
  CUtensorMap ma, mb;
  {
    uint64_t dims[4] = {(uint64_t)c.Cout, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_dy, (uint64_t)c.W * c.ld_dy, (uint64_t)c.H * c.W * c.ld_dy};
    uint32_t box[4] = {64, 8, (uint32_t)p.R, (uint32_t)p.NI};
    if (!make_map(&ma, c.dy, 4, dims, str, box)) return BD_ERR_CUDA;
  }
  {
    uint64_t dims[4] = {(uint64_t)c.Cin, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_x, (uint64_t)c.W * c.ld_x, (uint64_t)c.H * c.W * c.ld_x};
    uint32_t box[4] = {64, 10, (uint32_t)p.R, (uint32_t)p.NI};
    if (!make_map(&mb, c.x, 4, dims, str, box)) return BD_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(umma_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W3_SMEM); attr_set = true; }
  launch_pdl(umma_wgrad3_kernel, dim3(m_tiles * p.n_tiles, 3, splits), dim3(192), (size_t)W3_SMEM, st, ma, mb, p);
  count_launch(1);
  return BD_OK;
}

}  // namespace umma
}  // namespace bd
