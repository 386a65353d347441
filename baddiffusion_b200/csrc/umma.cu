// tcgen05 (UMMA) implicit-GEMM kernels for sm_100a: TMA-staged, im2col-free, fp16 operands, fp32 TMEM accumulators.
//
//   umma_fprop_kernel : D[pixel, n] = sum_{segments} A[pixel + shift, k] * B[n, k]
//       A = fp16 NHWC activation view, loaded as 4-D TMA boxes {64 ch, bw, bh, bn} (bw*bh*bn = 128 pixels) whose
//       origin is shifted per filter tap; TMA out-of-bounds zero fill *is* the convolution padding.
//       B = weights [z][N][K] (K-major) or [z][K][N] (MN-major: dgrad re-uses the forward weights untransposed,
//       attention re-uses V / K in place).  Serves conv 3x3/1x1 fwd + dgrad, every Linear, QK^T, PV, dP, dQ.
//   umma_wgrad_kernel : D[m, n] = sum_{pixels} A[pixel, m] * B[pixel + shift, n]   (both operands MN-major)
//       serves conv/linear weight gradients (split-K, fp32 atomics) and attention dK / dV (batched, fp16 out).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias / temb / residual -> global).  mbarrier ring of kStages stages.
#include "umma_common.cuh"

namespace bd {
namespace umma {

constexpr int BM = 128;
constexpr int BK = 64;            // fp16 elements per k-block = 128 B = one SWIZZLE_128B row
constexpr int kMaxSeg = 10;       // 9 taps + fused 1x1 shortcut

struct KSeg {
  int a_src;   // which A tensor map (0/1)
  int a_c0;    // first channel in that map
  int nblk;    // number of 64-wide k-blocks in this run
  int dx, dy;  // spatial shift of the tile origin
  int b_src;   // which B tensor map (0/1)
  int b_k0;    // K coordinate (fprop) in the B map
  int b_z;     // 3rd coordinate of the B map (tap)
};

struct FpropParams {
  int nseg;
  KSeg seg[kMaxSeg];
  int total_kblocks;
  int splits;            // split-K factor = cluster size along z (1: none)
  int mshare;            // CTAs of a cluster along x (neighbouring M tiles of one N tile) that SHARE every B (weight) tile:
                         // each fetches 1/mshare of it and TMA-multicasts the slice into all of them (1: none)
  int tiles_w, tiles_h;  // tiles per image
  int bw, bh, bn;        // pixel box (product 128)
  int W, H, NB;          // image geometry
  int HW;                // rows per sample for rowbias
  int N;                 // output columns
  int batched;           // B map z += image index (attention)
  int a_stride;          // A tile origin = tile origin * a_stride + shift (2 for the stride-2 Downsample2D conv)
  int64_t out_sn, out_sh, out_sw, out_off;  // output pixel index = n*out_sn + h*out_sh + w*out_sw + out_off
  // UMMA smem descriptor fields (16-byte units), host-provided so they can be overridden for bring-up
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t idesc;
  // epilogue
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
  float* gn_sums;        // GNS kernels: GroupNorm statistics of the output (see EpiArgs)
  int64_t ld_sums;
  int* error_flag;
  int dbg_shift, dbg_boff;  // bring-up experiment: A tile loaded dbg_shift rows early, descriptor start moved back
  int epi_pre;              // bit 0: one-tile kernel, bit 1: persistent kernel fetch the residual before the accumulator wait
};

struct WgradParams {
  int pbw, pbh, pbn;        // pixel box of one k-block (product 64)
  int ptiles_w, ptiles_h;   // k-blocks per image along W / H
  int kblocks_total;        // all k-blocks (NB/pbn * ptiles_h * ptiles_w), or per image when batched
  int b_stride, b_pad;      // B (= X) tile origin = k-block origin * b_stride + tap - b_pad
  int ks;                   // filter size (z = tap -> shift)
  int n_tiles;              // tiles along N
  int batched;              // z = image index (attention dK / dV)
  int splits;
  int Mtot, Ntot;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, idesc;
  void* y;
  int64_t ld_y;
  int out_mode;  // 0: fp32 atomicAdd, 1: fp32 store, 2: fp16 store
  int* error_flag;
};

template <int BN, int kStages>
struct Smem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * kStages + 1) * 8 + 16 + 1024 /*alignment slack*/;
};

// =============================================================================================
// fprop-style kernel
// =============================================================================================
template <int BN, int kStages, bool B_MN, bool GNS = false>   // GNS: the epilogue also accumulates GroupNorm statistics
__global__ void __launch_bounds__(320, BN == 256 ? 1 : 2) umma_fprop_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                            const __grid_constant__ CUtensorMap tmA1,
                                                            const __grid_constant__ CUtensorMap tmB0,
                                                            const __grid_constant__ CUtensorMap tmB1,
                                                            const FpropParams p) {
  using L = Smem<BN, kStages>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n_tile = blockIdx.y;
  const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB0);
    prefetch_tmap(&tmB1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], p.mshare);   // a stage is refilled only when every CTA that receives its B slices has consumed it
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  if (p.mshare > 1) cluster_sync_all();   // every CTA's barriers exist before a peer multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // prologue above overlaps the previous kernel; global memory is touched only below
  // cluster (CM, 1, S): rank = x + CM * z.  split-K: split index z owns k-blocks [it_begin, it_end) of the flattened
  // segment list; the CM CTAs with the same z walk the same k-blocks in lockstep and share the B tiles
  const int S = p.splits, CM = p.mshare;
  const int crank = (S > 1 || CM > 1) ? (int)cluster_ctarank() : 0;
  const int rank = crank / CM, mrank = crank - rank * CM;
  const uint16_t share_mask = (uint16_t)(((1u << CM) - 1u) << (rank * CM));
  const int it_begin = (int)((long long)p.total_kblocks * rank / S), it_end = (int)((long long)p.total_kblocks * (rank + 1) / S);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int s = 0; s < p.nseg && ok; ++s) {
        const KSeg sg = p.seg[s];
        const CUtensorMap* mapA = sg.a_src ? &tmA1 : &tmA0;
        const CUtensorMap* mapB = sg.b_src ? &tmB1 : &tmB0;
        for (int kb = 0; kb < sg.nblk; ++kb, ++it) {
          if (it < it_begin || it >= it_end) continue;   // another rank of the split-K cluster owns this k-block
          ok = mbar_wait(&empty[stage], phase ^ 1, p.error_flag, 1);
          if (!ok) break;
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_expect_tx(&full[stage], L::kStageBytes);
          tma_load_4d(mapA, &full[stage], sa, sg.a_c0 + kb * BK, w0 * p.a_stride + sg.dx - p.dbg_shift, h0 * p.a_stride + sg.dy, n0);
          const int bz = sg.b_z + (p.batched ? n0 : 0);
          if (CM == 1) {
            if (!B_MN) {
              tma_load_3d(mapB, &full[stage], sb, sg.b_k0 + kb * BK, n_tile * BN, bz);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_3d(mapB, &full[stage], sb + i * (64 * BK * 2), n_tile * BN + i * 64, sg.b_k0 + kb * BK, bz);
            }
          } else if (!B_MN) {
            // this CTA's BN/CM rows of the K-major tile, into the same slot of every CTA that shares it
            const int rows = BN / CM;
            tma_load_3d_mc(mapB, &full[stage], sb + mrank * rows * 128, sg.b_k0 + kb * BK, n_tile * BN + mrank * rows, bz, share_mask);
          } else {
            // MN-major tile = BN/64 slabs of [64 k][64 n]: this CTA's 64/CM k-rows of every slab
            const int krows = 64 / CM;
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_3d_mc(mapB, &full[stage], sb + i * (64 * BK * 2) + mrank * krows * 128, n_tile * BN + i * 64,
                             sg.b_k0 + kb * BK + mrank * krows, bz, share_mask);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int it = it_begin; it < it_end; ++it) {
        ok = mbar_wait(&full[stage], phase, p.error_flag, 2);
        if (!ok) break;
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t sb = sa + L::kABytes;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // K-major: advance 16 elements = 32 B inside the 128B swizzle row; MN-major: 16 k-rows = 2048 B
          const uint64_t ad = make_desc(sa + k * 32 + p.dbg_shift * 128, p.a_lbo, p.a_sbo) | ((uint64_t)(p.dbg_boff & 7) << 49);
          const uint64_t bd = make_desc(sb + (B_MN ? k * 2048 : k * 32), p.b_lbo, p.b_sbo);
          umma_f16(tmem_base, ad, bd, p.idesc, (it > it_begin || k > 0) ? 1u : 0u);
        }
        // frees the smem stage when these MMAs retire -- in every CTA that multicasts into it
        if (CM == 1) umma_commit(&empty[stage]); else umma_commit_mc(&empty[stage], share_mask);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (ok) umma_commit(tmem_full);
    }
  } else {
    // ===== epilogue: 8 warps, warp w owns TMEM lanes 32*(w%4) .. +31 and column half (w-2)/4 =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;  // row within the tile
    const int dn = r / (p.bw * p.bh), dh = (r / p.bw) % p.bh, dw = r % p.bw;
    const int n = n0 + dn, h = h0 + dh, w = w0 + dw;
    const bool valid = n < p.NB && h < p.H && w < p.W;
    const int64_t mlin = ((int64_t)n * p.H + h) * p.W + w;                        // dense pixel index (rowbias)
    const int64_t m = (int64_t)n * p.out_sn + h * p.out_sh + w * p.out_sw + p.out_off;  // output / residual pixel
    EpiArgs e{p.bias, p.bias2, p.residual, p.ld_res, p.scale, p.y, p.ld_y, p.out_f32, p.gn_sums, p.ld_sums};
    EpiResidual<BN / 2> pre;
    const bool use_pre = (p.epi_pre & 1) && S == 1 && BN == 128 && p.residual != nullptr;   // in flight while the MMAs run
    if (use_pre) epilogue_prefetch_residual<BN / 2>(pre, lane, m, valid, n_tile * BN + half * (BN / 2), e);
    const bool ok = mbar_wait(tmem_full, 0, p.error_flag, 3);
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * 32 * (BN / 2 + 4);  // operand stages are free now
    if (S == 1) {
      if (ok)
        epilogue_warp<BN / 2, GNS>(tmem_base + ((uint32_t)(q * 32) << 16) + half * (BN / 2), stage, lane, m, mlin, valid,
                                   n_tile * BN + half * (BN / 2), e, p.rowbias, p.ld_rowbias, p.HW, use_pre ? &pre : nullptr);
    } else {
      if (ok) epilogue_stage_warp<BN / 2>(tmem_base + ((uint32_t)(q * 32) << 16) + half * (BN / 2), stage, lane);
      cluster_sync_all();   // every rank's partial tile is staged (all threads of the cluster take part, see below)
      epilogue_splitk_finish_warp<BN / 2, GNS>(stage, lane, S, rank, m, mlin, valid, n_tile * BN + half * (BN / 2), e, p.rowbias,
                                               p.ld_rowbias, p.HW, CM, mrank);
    }
  }
  if (S > 1) {
    if (warp < 2) cluster_sync_all();   // producer / MMA warps: the barrier the epilogue warps passed after staging
    cluster_sync_all();                 // no CTA retires while a peer still reads its staging tile
  } else if (CM > 1) {
    cluster_sync_all();                 // ... or still multicasts into its ring / arrives on its barriers
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// =============================================================================================
// Persistent fprop kernel for grids with many small-K tiles (1x1 convs / Linears / attention GEMMs at 16x16 and up:
// K = 256-768 is 4-12 k-blocks, so a one-tile CTA spends most of its ~10 us life in fixed costs -- barrier init, TMEM
// allocation, first TMA round trip, epilogue drain, teardown).  One CTA per SM walks tiles t = blockIdx.x, += gridDim.x:
//   * the producer streams the operands of consecutive tiles through one ring without a break;
//   * TWO TMEM accumulators (2 x BN columns): the MMA warp fills buffer t&1 while the epilogue warps drain the other,
//     handed over with tmem_full[2] / tmem_empty[2] barriers;
//   * the epilogue's fp32 staging tiles have their own shared memory (in the one-tile kernel they alias the operand
//     ring, which must be idle for that).
// Same operands, segment list and epilogue (`epilogue_warp`) as umma_fprop_kernel: results are bit-identical.
// =============================================================================================
template <int BN, int kStages>
struct SmemP {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiOffset = kStages * kStageBytes;
  static constexpr int kEpiBytes = 8 * 32 * (BN / 2 + 4) * 4;
  static constexpr int kBarOffset = kEpiOffset + kEpiBytes;
  static constexpr int kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16 + 1024;
};

template <int BN, int kStages, bool B_MN, bool GNS = false>
__global__ void __launch_bounds__(320, 1) umma_fprop_persistent_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                                       const __grid_constant__ CUtensorMap tmA1,
                                                                       const __grid_constant__ CUtensorMap tmB0,
                                                                       const __grid_constant__ CUtensorMap tmB1,
                                                                       const FpropParams p, const int m_tiles,
                                                                       const int n_tiles) {
  using L = SmemP<BN, kStages>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB0);
    prefetch_tmap(&tmB1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 8);   // one arrival per epilogue warp
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer: free-running over all tiles of this CTA =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x) {
        const int m_tile = tile / n_tiles, n_tile = tile - m_tile * n_tiles;   // n fastest: neighbours share the A tile in L2
        const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        for (int s = 0; s < p.nseg && ok; ++s) {
          const KSeg sg = p.seg[s];
          const CUtensorMap* mapA = sg.a_src ? &tmA1 : &tmA0;
          const CUtensorMap* mapB = sg.b_src ? &tmB1 : &tmB0;
          for (int kb = 0; kb < sg.nblk; ++kb) {
            ok = mbar_wait(&empty[stage], phase ^ 1, p.error_flag, 1);
            if (!ok) break;
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            mbar_expect_tx(&full[stage], L::kStageBytes);
            tma_load_4d(mapA, &full[stage], sa, sg.a_c0 + kb * BK, w0 * p.a_stride + sg.dx, h0 * p.a_stride + sg.dy, n0);
            const int bz = sg.b_z + (p.batched ? n0 : 0);
            if (!B_MN) {
              tma_load_3d(mapB, &full[stage], sb, sg.b_k0 + kb * BK, n_tile * BN, bz);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_3d(mapB, &full[stage], sb + i * (64 * BK * 2), n_tile * BN + i * 64, sg.b_k0 + kb * BK, bz);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, eph[2] = {0, 0};
      bool ok = true;
      int lt = 0;
      for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        ok = mbar_wait(&tmem_empty[buf], eph[buf] ^ 1, p.error_flag, 4);   // drained by the epilogue two tiles ago
        if (!ok) break;
        eph[buf] ^= 1;
        tc_fence_after();
        for (int it = 0; it < p.total_kblocks; ++it) {
          ok = mbar_wait(&full[stage], phase, p.error_flag, 2);
          if (!ok) break;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = make_desc(sa + k * 32, p.a_lbo, p.a_sbo);
            const uint64_t bd = make_desc(sb + (B_MN ? k * 2048 : k * 32), p.b_lbo, p.b_sbo);
            umma_f16(tmem_base + buf * BN, ad, bd, p.idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (ok) umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ===== epilogue: 8 warps, warp w owns TMEM lanes 32*(w%4) .. +31 and column half (w-2)/4 =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int dn = r / (p.bw * p.bh), dh = (r / p.bw) % p.bh, dw = r % p.bw;
    float* stage = reinterpret_cast<float*>(smem + L::kEpiOffset) + (warp - 2) * 32 * (BN / 2 + 4);
    EpiArgs e{p.bias, p.bias2, p.residual, p.ld_res, p.scale, p.y, p.ld_y, p.out_f32, p.gn_sums, p.ld_sums};
    uint32_t fph[2] = {0, 0};
    bool ok = true;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total && ok; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int m_tile = tile / n_tiles, n_tile = tile - m_tile * n_tiles;
      const int tw = m_tile % p.tiles_w, th = (m_tile / p.tiles_w) % p.tiles_h, tn = m_tile / (p.tiles_w * p.tiles_h);
      const int n = tn * p.bn + dn, h = th * p.bh + dh, w = tw * p.bw + dw;
      const bool valid = n < p.NB && h < p.H && w < p.W;
      const int64_t mlin = ((int64_t)n * p.H + h) * p.W + w;
      const int64_t m = (int64_t)n * p.out_sn + h * p.out_sh + w * p.out_sw + p.out_off;
      EpiResidual<BN / 2> pre;
      const bool use_pre = (p.epi_pre & 2) && BN == 128 && p.residual != nullptr;   // in flight while this tile's MMAs run
      if (use_pre) epilogue_prefetch_residual<BN / 2>(pre, lane, m, valid, n_tile * BN + half * (BN / 2), e);
      ok = mbar_wait(&tmem_full[buf], fph[buf], p.error_flag, 3);
      if (!ok) break;
      fph[buf] ^= 1;
      tc_fence_after();
      epilogue_warp<BN / 2, GNS>(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * (BN / 2), stage, lane, m, mlin, valid,
                                 n_tile * BN + half * (BN / 2), e, p.rowbias, p.ld_rowbias, p.HW, use_pre ? &pre : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// =============================================================================================
// wgrad-style kernel: both operands MN-major, K = pixels.  grid (m_tiles*n_tiles, z, splits)
// =============================================================================================
template <int BN, int kStages>
__global__ void __launch_bounds__(192, 2) umma_wgrad_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const WgradParams p) {
  using L = Smem<BN, kStages>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x / p.n_tiles, n_tile = blockIdx.x % p.n_tiles;
  const int z = blockIdx.y;
  int dx = 0, dy = 0;
  if (!p.batched && p.ks == 3) { dy = z / 3 - p.b_pad; dx = z % 3 - p.b_pad; }
  const int per = (p.kblocks_total + p.splits - 1) / p.splits;
  const int kb_begin = blockIdx.z * per;
  const int kb_end = min(p.kblocks_total, kb_begin + per);
  const int nk = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
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
          const int pw = (kb % p.ptiles_w) * p.pbw, ph = ((kb / p.ptiles_w) % p.ptiles_h) * p.pbh;
          const int pn = p.batched ? z : (kb / (p.ptiles_w * p.ptiles_h)) * p.pbn;
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_expect_tx(&full[stage], L::kStageBytes);
#pragma unroll
          for (int i = 0; i < BM / 64; ++i)
            tma_load_4d(&tmA, &full[stage], sa + i * (64 * BK * 2), m_tile * BM + i * 64, pw, ph, pn);
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            tma_load_4d(&tmB, &full[stage], sb + i * (64 * BK * 2), n_tile * BN + i * 64, pw * p.b_stride + dx, ph * p.b_stride + dy, pn);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
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
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = make_desc(sa + k * 2048, p.a_lbo, p.a_sbo);
            const uint64_t bd = make_desc(sb + k * 2048, p.b_lbo, p.b_sbo);
            umma_f16(tmem_base, ad, bd, p.idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (ok) umma_commit(tmem_full);
      }
    } else {
      const int q = warp & 3;
      const int row = m_tile * BM + q * 32 + lane;
      const bool ok = mbar_wait(tmem_full, 0, p.error_flag, 3);
      tc_fence_after();
      if (ok) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
          const int col = n_tile * BN + c0;
          if (row < p.Mtot && col < p.Ntot) {
            const int64_t off = ((int64_t)z * p.Mtot + row) * p.ld_y + col;
            if (p.out_mode == 0) {
              float* yr = reinterpret_cast<float*>(p.y) + off;   // 16-byte aligned: ld_y and col are multiples of 32
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yr + j), "f"(__uint_as_float(v[j])),
                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                             : "memory");
            } else if (p.out_mode == 1) {
              float* yr = reinterpret_cast<float*>(p.y) + off;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(yr + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
              __half* yr = reinterpret_cast<__half*>(p.y) + off;
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
#pragma unroll
              for (int j = 0; j < 32; j += 8) *reinterpret_cast<half8*>(yr + j) = pack8(f + j);
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
    tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// rank-4 fp16 map: dims {d0,d1,d2,d3} (d0 contiguous), strides in ELEMENTS for d1..d3, box {b0..b3}
bool make_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_elems,
              const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return false; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_elems[i] * 2;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims {%llu,%llu,%llu,%llu} box {%u,%u,%u,%u} ptr %p stride1 %llu",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, ptr,
              (unsigned long long)gs[0]);
    return false;
  }
  return true;
}

// one device int PER DEVICE, lazily allocated on first use there (not per call): a process that drives several GPUs
// (nn.DataParallel-style callers) gets a valid flag on each
static int* g_error_flag[64] = {nullptr};
static int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= 64) ? 0 : dev;
}
int* error_flag() {
  const int d = current_device_slot();
  if (!g_error_flag[d]) {
    cudaMalloc(&g_error_flag[d], sizeof(int));
    cudaMemset(g_error_flag[d], 0, sizeof(int));
  }
  return g_error_flag[d];
}

uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* v = getenv(name);
  return v ? (uint32_t)strtoul(v, nullptr, 0) : dflt;
}

// choose the 128-pixel box for a (H, W) image grid; returns false if the geometry does not tile
static bool pick_box(int H, int W, int pixels, int* bw, int* bh, int* bn) {
  if (W >= pixels) {
    if (W % pixels) return false;
    *bw = pixels; *bh = 1; *bn = 1;
    return true;
  }
  if (pixels % W) return false;
  int rows = pixels / W;
  if (H >= rows) {
    if (H % rows) return false;
    *bw = W; *bh = rows; *bn = 1;
    return true;
  }
  if (rows % H) return false;
  *bw = W; *bh = H; *bn = rows / H;
  return *bw <= 256 && *bh <= 256 && *bn <= 256;
}

// kStages = 3 (3 x 32 KB): two CTAs per SM, so one drains its accumulator while the other streams operands -- for grids
// with more CTAs than SMs.  kStages = 6: grids that put at most one CTA on an SM anyway (the 8x8 / 4x4 levels) are a
// chain of TMA round trips (a k-block's MMAs take 256 cycles, a stage refill ~1 us): twice the loads in flight.
template <int BN, bool B_MN, int kStages, bool GNS = false>
static int launch_fprop_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& b1,
                          const FpropParams& p, dim3 grid, cudaStream_t st) {
  using L = Smem<BN, kStages>;
  auto kern = umma_fprop_kernel<BN, kStages, B_MN, GNS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    attr_set = true;
  }
  if (p.splits > 1 || p.mshare > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = (size_t)L::kTotal;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.mshare;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = p.splits;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = getenv("BD_NO_PDL") ? 1 : 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a0, a1, b, b1, p);
    if (e != cudaSuccess) { set_error("umma fprop: cluster launch (%d,1,%d) failed: %s", p.mshare, p.splits, cudaGetErrorString(e)); return BD_ERR_CUDA; }
  } else {
    launch_pdl(kern, grid, dim3(320), (size_t)L::kTotal, st, a0, a1, b, b1, p);
  }
  count_launch(1);
  return 0;
}

// Generic fprop launch used by conv fwd / dgrad / linear / attention GEMMs.
//   A: fp16 view (NB, H, W, Ca) with ld_a; optional second source for fused K segments.
//   B: fp16 [Z][rows][cols] with row length ldb: K-major (rows = N, cols = K) or MN-major (rows = K, cols = N).
struct FpropCall {
  const void* a; int64_t ld_a; int Ca;
  const void* a2; int64_t ld_a2; int Ca2;
  int NB, H, W;
  const void* b; int b_rows, b_cols, b_z; int64_t ld_b;   // main weights
  const void* b2; int64_t ld_b2;   // weights of the fused second segment [N][Ca2] (K-major), nullable
  int N, ks; bool b_mn; bool batched; bool flip_taps;
  const float* bias; const float* bias2; const float* rowbias; int64_t ld_rowbias; int HW_rowbias;
  const void* residual; int64_t ld_res; float scale; void* y; int64_t ld_y; int out_f32;
  // --- optional generalisations (all zero = plain stride-1 conv) ---
  int a_stride;            // 2: A is sampled at every other pixel (TMA elementStrides); a_H/a_W = dims of the A image
  int a_H, a_W;
  int ntap_override;       // > 0: explicit tap list instead of the ks*ks grid
  int tap_dx[9], tap_dy[9], tap_z[9];
  int64_t out_sn, out_sh, out_sw, out_off;  // output pixel addressing (0 = dense NHWC)
  float* gn_sums; int64_t ld_sums;   // forward only: GroupNorm statistics of y accumulated in the epilogue (nullable)
};

int fprop_supported(const FpropCall& c) {
  int bw, bh, bn;
  if (c.Ca % 64 || (c.a2 && c.Ca2 % 64)) return 0;
  if (c.N % 64) return 0;
  if (c.ld_a % 8 || (c.a2 && c.ld_a2 % 8) || c.ld_b % 8) return 0;
  if (((uintptr_t)c.a & 15) || ((uintptr_t)c.b & 15) || (c.a2 && ((uintptr_t)c.a2 & 15))) return 0;
  if (!pick_box(c.H, c.W, BM, &bw, &bh, &bn)) return 0;
  if (c.batched && bn != 1) return 0;
  if (c.ks != 1 && c.ks != 3) return 0;
  return 1;
}

// in-epilogue GroupNorm statistics on the generic kernels: forward (K-major weights), fp16 output, whole 8-pixel row groups
// of one image per epilogue iteration (HW % 8 == 0), not the batched (attention) GEMMs.  OFF by default (BD_GN_SUMS_GENERIC=1
// enables): with them ALL 51 forward GroupNorms of the CIFAR10 UNet become streaming passes, but measured inside the step
// (scripts/ab_step.py) that LOSES 0.05 ms in training and changes nothing in sampling -- at 8x8 / 4x4 the GroupNorm launches
// are latency-, not reduction-bound, and the atomics make even the tiny test configs run-to-run non-reproducible.
int fprop_gn_sums_supported(const FpropCall& c) {
  const int HW = c.HW_rowbias > 0 ? c.HW_rowbias : c.H * c.W;
  return fprop_supported(c) && !c.b_mn && !c.batched && !c.out_f32 && HW % 8 == 0 && c.N % 128 == 0 && !getenv("BD_NO_GN_SUMS") &&
         getenv("BD_GN_SUMS_GENERIC");
}

int fprop_launch(const FpropCall& c, cudaStream_t st) {
  FpropParams p;
  memset(&p, 0, sizeof(p));
  if (c.gn_sums && !fprop_gn_sums_supported(c)) { set_error("umma fprop: gn_sums is not supported for this launch"); return BD_ERR_UNSUPPORTED; }
  p.gn_sums = c.gn_sums; p.ld_sums = c.ld_sums;
  if (!pick_box(c.H, c.W, BM, &p.bw, &p.bh, &p.bn)) { set_error("umma fprop: geometry %dx%d does not tile", c.H, c.W); return BD_ERR_UNSUPPORTED; }
  p.tiles_w = c.W / p.bw;
  p.tiles_h = c.H / p.bh;
  p.W = c.W; p.H = c.H; p.NB = c.NB;
  p.HW = c.HW_rowbias > 0 ? c.HW_rowbias : c.H * c.W;
  p.N = c.N;
  p.batched = c.batched ? 1 : 0;
  // BN = 256 (one CTA computes both 128-channel halves of a 256-multiple layer from ONE fetch of its A tiles) for the
  // few-tile, long-K launches (the 3x3 convolutions at 8x8 / 4x4): OFF by default.  The idea was that the per-tap A
  // re-fetch makes those launches L2->SM bound; measured inside the train step (scripts/ab_step.py, profiles/
  // r2_ab_step_mshare.log) it LOSES 0.57 ms/step -- half as many, twice as fat CTAs (192 KB of shared memory, one per SM)
  // no longer share SMs with the weight-gradient kernels of the side stream, and the launches are latency- not
  // bandwidth-bound.  BD_BN256=1 enables it for experiments.
  int BN = (c.N % 128 == 0) ? 128 : 64;
  {
    int bw, bh, bn;
    pick_box(c.H, c.W, BM, &bw, &bh, &bn);
    const long long mt = (long long)(c.W / bw) * (c.H / bh) * ceil_div(c.NB, bn);
    const int taps0 = c.ntap_override > 0 ? c.ntap_override : c.ks * c.ks;
    const int kblocks0 = taps0 * (c.Ca / 64) + (c.a2 ? c.Ca2 / 64 : 0);
    if (c.N % 256 == 0 && !c.batched && kblocks0 >= 18 && mt * (c.N / 128) <= num_sms() && getenv("BD_BN256")) BN = 256;
  }
  // K segments
  int nseg = 0, total = 0;
  const int taps = c.ntap_override > 0 ? c.ntap_override : c.ks * c.ks;
  for (int t = 0; t < taps; ++t) {
    KSeg& s = p.seg[nseg++];
    s.a_src = 0; s.a_c0 = 0; s.nblk = c.Ca / 64;
    if (c.ntap_override > 0) {
      s.dx = c.tap_dx[t]; s.dy = c.tap_dy[t]; s.b_z = c.tap_z[t];
    } else {
      int r = t / c.ks, q = t % c.ks;
      s.dy = c.ks == 3 ? r - 1 : 0;
      s.dx = c.ks == 3 ? q - 1 : 0;
      if (c.flip_taps) { s.dy = -s.dy; s.dx = -s.dx; }  // dgrad: dX[q] = sum_tap dY[q - off(tap)] W[tap]^T
      s.b_z = t;
    }
    s.b_src = 0; s.b_k0 = 0;
    total += s.nblk;
  }
  p.a_stride = c.a_stride > 1 ? c.a_stride : 1;
  if (c.out_sn || c.out_sh || c.out_sw) { p.out_sn = c.out_sn; p.out_sh = c.out_sh; p.out_sw = c.out_sw; p.out_off = c.out_off; }
  else { p.out_sn = (int64_t)c.H * c.W; p.out_sh = c.W; p.out_sw = 1; p.out_off = 0; }
  if (c.a2) {
    KSeg& s = p.seg[nseg++];
    s.a_src = 1; s.a_c0 = 0; s.nblk = c.Ca2 / 64; s.dx = s.dy = 0; s.b_src = 1; s.b_k0 = 0; s.b_z = 0;
    total += s.nblk;
  }
  p.nseg = nseg;
  p.total_kblocks = total;
  // descriptors
  p.a_lbo = env_u32("BD_UMMA_AK_LBO", 1);
  p.a_sbo = env_u32("BD_UMMA_AK_SBO", 64);
  if (c.b_mn) { p.b_lbo = env_u32("BD_UMMA_BMN_LBO", 512); p.b_sbo = env_u32("BD_UMMA_BMN_SBO", 64); }
  else        { p.b_lbo = env_u32("BD_UMMA_BK_LBO", 1);   p.b_sbo = env_u32("BD_UMMA_BK_SBO", 64); }
  p.idesc = (1u << 4) | ((c.b_mn ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  p.bias = c.bias; p.bias2 = c.bias2; p.rowbias = c.rowbias; p.ld_rowbias = c.ld_rowbias;
  p.residual = (const __half*)c.residual; p.ld_res = c.ld_res; p.scale = c.scale;
  p.y = c.y; p.ld_y = c.ld_y; p.out_f32 = c.out_f32;
  p.error_flag = error_flag();
  p.dbg_shift = (int)env_u32("BD_UMMA_DBG_SHIFT", 0);
  p.epi_pre = (int)env_u32("BD_EPI_PRE", 3);
  p.dbg_boff = (int)env_u32("BD_UMMA_DBG_BOFF", 0);

  CUtensorMap ma0, ma1, mb, mb1;
  {
    const int aH = c.a_stride > 1 ? c.a_H : c.H, aW = c.a_stride > 1 ? c.a_W : c.W;
    const uint32_t st = (uint32_t)p.a_stride;
    uint64_t dims[4] = {(uint64_t)c.Ca, (uint64_t)aW, (uint64_t)aH, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_a, (uint64_t)aW * c.ld_a, (uint64_t)aH * aW * c.ld_a};
    // with elementStrides s the box spans boxDim input elements and delivers ceil(boxDim / s) of them
    uint32_t box[4] = {64, (uint32_t)p.bw * st, (uint32_t)p.bh * st, (uint32_t)p.bn};
    uint32_t es[4] = {1, st, st, 1};
    if (!make_map(&ma0, c.a, 4, dims, str, box, es)) return BD_ERR_CUDA;
  }
  if (c.a2) {
    uint64_t dims[4] = {(uint64_t)c.Ca2, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_a2, (uint64_t)c.W * c.ld_a2, (uint64_t)c.H * c.W * c.ld_a2};
    uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (!make_map(&ma1, c.a2, 4, dims, str, box)) return BD_ERR_CUDA;
  } else {
    ma1 = ma0;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * ceil_div(c.NB, p.bn);
  // split-K (cluster of S CTAs per output tile) when the tile grid leaves most SMs idle and K is long: the 3x3 convs
  // of the 8x8 / 4x4 levels (128 / 32 tiles, 36-72 k-blocks each)
  p.splits = 1;
  if (!getenv("BD_NO_SPLITK")) {
    const int tiles = m_tiles * (c.N / BN);
    const int budget = (int)env_u32("BD_SPLITK_CTAS", num_sms());   // stay at one CTA per SM: the 6-stage ring below
    const int max_splits = (int)env_u32("BD_SPLITK_MAX", BN == 256 ? 8 : 4), min_kb = BN == 256 ? 4 : 6;
    while (p.splits < max_splits && tiles * p.splits * 2 <= budget && p.total_kblocks / (p.splits * 2) >= min_kb) p.splits *= 2;
  }
  // B-tile sharing: in the few-tile, long-K regime the weight tiles are the larger half of the L2->SM traffic and every M
  // tile of an N tile reads the same ones.  The `mshare` CTAs of a cluster along x fetch 1/mshare of each B tile and
  // multicast it (cluster = (mshare, 1, splits) <= 8 CTAs).  Correct (tests/test_kernels_gpu.py bench-sized cases run it
  // with BD_FPROP_MSHARE=4) but OFF by default: inside the train step it costs +0.4 ms (clusters of 8 one-CTA-per-SM
  // blocks wait for 8 free SMs of one GPC while the side stream's kernels hold some, and the sharing CTAs move in
  // lockstep).  BD_FPROP_MSHARE=2|4 enables it.
  p.mshare = 1;
  {
    const long long ctas = (long long)m_tiles * (c.N / BN) * p.splits;
    const int want = (int)env_u32("BD_FPROP_MSHARE", 1);
    if (!c.batched && c.ks == 3 && ctas <= num_sms() && p.total_kblocks / p.splits >= 4 && p.dbg_shift == 0) {
      int cm = want;
      while (cm > 1 && (m_tiles % cm || cm * p.splits > 8 || (c.b_mn ? 64 % cm : BN % cm))) cm >>= 1;
      p.mshare = cm < 1 ? 1 : cm;
    }
  }
  {
    uint64_t dims[3] = {(uint64_t)c.b_cols, (uint64_t)c.b_rows, (uint64_t)c.b_z};
    uint64_t str[2] = {(uint64_t)c.ld_b, (uint64_t)c.b_rows * c.ld_b};
    uint32_t box[3] = {64, (c.b_mn ? 64u : (uint32_t)BN) / (uint32_t)p.mshare, 1};   // mshare > 1: the slice one CTA multicasts
    if (!make_map(&mb, c.b, 3, dims, str, box)) return BD_ERR_CUDA;
  }
  if (c.a2) {
    uint64_t dims[3] = {(uint64_t)c.Ca2, (uint64_t)c.N, 1};
    uint64_t str[2] = {(uint64_t)c.ld_b2, (uint64_t)c.N * c.ld_b2};
    uint32_t box[3] = {64, (uint32_t)BN / (uint32_t)p.mshare, 1};
    if (!make_map(&mb1, c.b2, 3, dims, str, box)) return BD_ERR_CUDA;
  } else {
    mb1 = mb;
  }
  dim3 grid(m_tiles, c.N / BN, p.splits);
  // many small-K tiles: persistent kernel (one CTA per SM, double-buffered TMEM)
  if (BN == 128 && p.splits == 1 && p.mshare == 1 && (long long)m_tiles * (c.N / BN) > 2 * num_sms() && p.total_kblocks <= 16 && p.dbg_shift == 0 &&
      !getenv("BD_NO_FPROP_PERSIST")) {
    constexpr int kSt = 4;
    using LP = SmemP<128, kSt>;
    const int nt = c.N / BN;
    const int ctas = num_sms();
    static bool pattr[2] = {false, false};
    if (c.gn_sums) {
      static bool gattr = false;
      if (!gattr) { cudaFuncSetAttribute(umma_fprop_persistent_kernel<128, kSt, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LP::kTotal); gattr = true; }
      launch_pdl(umma_fprop_persistent_kernel<128, kSt, false, true>, dim3(ctas), dim3(320), (size_t)LP::kTotal, st, ma0, ma1, mb, mb1, p, m_tiles, nt);
    } else if (c.b_mn) {
      if (!pattr[1]) { cudaFuncSetAttribute(umma_fprop_persistent_kernel<128, kSt, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LP::kTotal); pattr[1] = true; }
      launch_pdl(umma_fprop_persistent_kernel<128, kSt, true>, dim3(ctas), dim3(320), (size_t)LP::kTotal, st, ma0, ma1, mb, mb1, p, m_tiles, nt);
    } else {
      if (!pattr[0]) { cudaFuncSetAttribute(umma_fprop_persistent_kernel<128, kSt, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LP::kTotal); pattr[0] = true; }
      launch_pdl(umma_fprop_persistent_kernel<128, kSt, false>, dim3(ctas), dim3(320), (size_t)LP::kTotal, st, ma0, ma1, mb, mb1, p, m_tiles, nt);
    }
    count_launch(1);
    return BD_OK;
  }
  const bool deep = (long long)grid.x * grid.y * grid.z <= num_sms() && p.total_kblocks / p.splits >= 6 && !getenv("BD_NO_DEEP_RING");
  if (c.gn_sums) {   // forward, BN in {128, 256} (checked above)
    if (BN == 256) return launch_fprop_t<256, false, 4, true>(ma0, ma1, mb, mb1, p, grid, st);
    if (deep) return launch_fprop_t<128, false, 6, true>(ma0, ma1, mb, mb1, p, grid, st);
    return launch_fprop_t<128, false, 3, true>(ma0, ma1, mb, mb1, p, grid, st);
  }
  if (BN == 256) {
    if (c.b_mn) launch_fprop_t<256, true, 4>(ma0, ma1, mb, mb1, p, grid, st); else launch_fprop_t<256, false, 4>(ma0, ma1, mb, mb1, p, grid, st);
  } else if (BN == 128) {
    if (deep) {
      if (c.b_mn) launch_fprop_t<128, true, 6>(ma0, ma1, mb, mb1, p, grid, st); else launch_fprop_t<128, false, 6>(ma0, ma1, mb, mb1, p, grid, st);
    } else {
      if (c.b_mn) launch_fprop_t<128, true, 3>(ma0, ma1, mb, mb1, p, grid, st); else launch_fprop_t<128, false, 3>(ma0, ma1, mb, mb1, p, grid, st);
    }
  } else {
    if (c.b_mn) launch_fprop_t<64, true, 3>(ma0, ma1, mb, mb1, p, grid, st); else launch_fprop_t<64, false, 3>(ma0, ma1, mb, mb1, p, grid, st);
  }
  return BD_OK;
}

// wgrad-style launch.  A = "dY" view (NB,H,W,Ma) -> M rows; B = "X" view (NB,H,W,Nb) -> N cols.
struct WgradCall {
  const void* a; int64_t ld_a; int Mtot;
  const void* b; int64_t ld_b; int Ntot;
  int NB, H, W, ks; bool batched;
  void* y; int64_t ld_y; int out_mode;  // 0 atomic f32, 1 store f32, 2 store f16
  int zcount;                            // taps, or batch count
  int b_stride, b_pad, b_H, b_W;         // stride-2 conv: B (= X) image dims / sampling stride / padding (0 = defaults)
};

int wgrad_supported(const WgradCall& c) {
  int bw, bh, bn;
  if (c.Mtot % 64 || c.Ntot % 64) return 0;
  if (c.ld_a % 8 || c.ld_b % 8) return 0;
  if (((uintptr_t)c.a & 15) || ((uintptr_t)c.b & 15)) return 0;
  if (!pick_box(c.H, c.W, 64, &bw, &bh, &bn)) return 0;
  if (c.batched && bn != 1) return 0;
  if (c.ks != 1 && c.ks != 3) return 0;
  return 1;
}

template <int BN, int kStages>
static int launch_wgrad_t(const CUtensorMap& a, const CUtensorMap& b, const WgradParams& p, dim3 grid, cudaStream_t st) {
  using L = Smem<BN, kStages>;   // 3 stages: two CTAs per SM; 6 stages: one CTA per SM with twice the loads in flight
  auto kern = umma_wgrad_kernel<BN, kStages>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    attr_set = true;
  }
  launch_pdl(kern, grid, dim3(192), (size_t)L::kTotal, st, a, b, p);
  count_launch(1);
  return 0;
}

int wgrad_launch(const WgradCall& c, cudaStream_t st) {
  WgradParams p;
  memset(&p, 0, sizeof(p));
  if (!pick_box(c.H, c.W, 64, &p.pbw, &p.pbh, &p.pbn)) { set_error("umma wgrad: geometry %dx%d does not tile", c.H, c.W); return BD_ERR_UNSUPPORTED; }
  p.ptiles_w = c.W / p.pbw;
  p.ptiles_h = c.H / p.pbh;
  const int per_image = p.ptiles_w * p.ptiles_h;
  p.kblocks_total = c.batched ? per_image : per_image * ceil_div(c.NB, p.pbn);
  p.ks = c.ks;
  p.batched = c.batched ? 1 : 0;
  p.b_stride = c.b_stride > 1 ? c.b_stride : 1;
  p.b_pad = c.b_stride > 1 ? c.b_pad : 1;
  p.Mtot = c.Mtot; p.Ntot = c.Ntot;
  const int BN = (c.Ntot % 128 == 0) ? 128 : 64;
  p.n_tiles = c.Ntot / BN;
  const int m_tiles = ceil_div(c.Mtot, BM);
  const int tiles = m_tiles * p.n_tiles * c.zcount;
  int splits = 1;
  const bool deep_ok = getenv("BD_WGRAD_DEEP") != nullptr;   // measured slower (10.25 vs 10.14 ms/step): off by default
  if (c.out_mode == 0) {
    // split-K over pixel k-blocks.  With the 6-stage ring: as many splits as keep the grid within one CTA per SM (fewer
    // partial tiles to reduce atomically); otherwise fill two CTAs per SM
    splits = deep_ok ? num_sms() / tiles : ceil_div(num_sms(), tiles);
    int maxs = p.kblocks_total / 8 > 0 ? p.kblocks_total / 8 : 1;  // >= 8 k-blocks per CTA
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  p.splits = splits;
  p.a_lbo = env_u32("BD_UMMA_AMN_LBO", 512);
  p.a_sbo = env_u32("BD_UMMA_AMN_SBO", 64);
  p.b_lbo = env_u32("BD_UMMA_BMN_LBO", 512);
  p.b_sbo = env_u32("BD_UMMA_BMN_SBO", 64);
  p.idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  p.y = c.y; p.ld_y = c.ld_y; p.out_mode = c.out_mode;
  p.error_flag = error_flag();
  CUtensorMap ma, mb;
  uint32_t box[4] = {64, (uint32_t)p.pbw, (uint32_t)p.pbh, (uint32_t)p.pbn};
  {
    uint64_t dims[4] = {(uint64_t)c.Mtot, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_a, (uint64_t)c.W * c.ld_a, (uint64_t)c.H * c.W * c.ld_a};
    if (!make_map(&ma, c.a, 4, dims, str, box)) return BD_ERR_CUDA;
  }
  {
    const int bH = c.b_stride > 1 ? c.b_H : c.H, bW = c.b_stride > 1 ? c.b_W : c.W;
    const uint32_t st = (uint32_t)p.b_stride;
    uint64_t dims[4] = {(uint64_t)c.Ntot, (uint64_t)bW, (uint64_t)bH, (uint64_t)c.NB};
    uint64_t str[3] = {(uint64_t)c.ld_b, (uint64_t)bW * c.ld_b, (uint64_t)bH * bW * c.ld_b};
    uint32_t bbox[4] = {64, (uint32_t)p.pbw * st, (uint32_t)p.pbh * st, (uint32_t)p.pbn};
    uint32_t es[4] = {1, st, st, 1};
    if (!make_map(&mb, c.b, 4, dims, str, bbox, es)) return BD_ERR_CUDA;
  }
  dim3 grid(m_tiles * p.n_tiles, c.zcount, splits);
  const bool deep = deep_ok && (long long)grid.x * grid.y * grid.z <= num_sms() && p.kblocks_total / splits >= 6;
  if (BN == 128) {
    if (deep) launch_wgrad_t<128, 6>(ma, mb, p, grid, st); else launch_wgrad_t<128, 3>(ma, mb, p, grid, st);
  } else {
    launch_wgrad_t<64, 3>(ma, mb, p, grid, st);
  }
  return BD_OK;
}

int read_error_flag() {
  int v = 0;
  int* f = g_error_flag[current_device_slot()];
  if (f) {
    cudaMemcpy(&v, f, sizeof(int), cudaMemcpyDeviceToHost);
    if (v) cudaMemset(f, 0, sizeof(int));
  }
  return v;
}

}  // namespace umma
}  // namespace bd
