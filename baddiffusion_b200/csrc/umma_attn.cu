// Fused single-head attention on tcgen05 (D/models/attention.py:135-173): ONE kernel per (image, 128-query block) does
//     phase 1   S = A B1^T            128 x S scores in TMEM (fp32), A = Q rows, B1 = K           [forward]
//     row op    P = softmax(scale S)  in registers: one TMEM read, one ex2 per element, fp16 P -> shared memory in the
//                                     UMMA K-major SWIZZLE_128B layout (the A operand of phase 2) and -> global (saved for backward)
//     phase 2   O = P B2              128 x C in TMEM, B2 = V read in place as an MN-major operand; coalesced fp16 epilogue
// The fp32 score matrix never leaves the SM (the three-launch path wrote B*S*S fp32 to HBM, re-read it in a softmax
// kernel, wrote fp16 P and re-read it for P V).  The same kernel with a different row op is the data-gradient half of
// the backward:   phase 1  dP = dO V^T,   row op  dS = scale P (dP - rowsum(dP P)) (P re-read from the saved fp16 probs),
// phase 2  dQ = dS K; dS also goes to global for dK = dS^T Q, which like dV = P^T dO stays a batched weight-gradient GEMM.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 = row op + epilogue:
// warp w owns TMEM lane quadrant w & 3 (32 query rows, one per thread) and column half (w - 2) >> 2 of the score row,
// so a thread keeps S/2 <= 128 scores in registers; row max / row sum meet through two tiny shared-memory exchanges.
// TMEM: scores at column 0 (S <= 256 columns) + output accumulator at column 256 (C <= 256 columns) = 512 columns.
// Shared memory: 3-stage operand ring (48 KB stages: phase 1 = 16 KB Q k-block + 32 KB K k-block, phase 2 = 32 KB V
// key-block) + 64 KB P + barriers; the epilogue's fp32 staging tiles reuse the ring once every MMA has retired.
#include "umma_common.cuh"

namespace bd {
namespace umma {

constexpr int AT_STAGES = 3;
constexpr int AT_STAGE_BYTES = 48 * 1024;
constexpr int AT_A_BYTES = 128 * 64 * 2;               // phase-1 A k-block (128 rows x 64 channels)
constexpr int AT_P_OFFSET = AT_STAGES * AT_STAGE_BYTES;
constexpr int AT_P_BYTES = 128 * 256 * 2;
constexpr int AT_XCH_OFFSET = AT_P_OFFSET + AT_P_BYTES;   // float[2][2][128] exchange
constexpr int AT_BAR_OFFSET = AT_XCH_OFFSET + 2 * 2 * 128 * 4;
constexpr int AT_SMEM = AT_BAR_OFFSET + (2 * AT_STAGES + 3) * 8 + 16 + 1024;
static_assert(AT_SMEM <= 232448, "attention shared memory");

struct AttnParams {
  int S, C;
  int nkb1, nkb2;          // C / 64 channel blocks (phase 1), S / 64 key blocks (phase 2)
  uint32_t idesc1, idesc2;
  float scale;
  const __half* probs_in;  // MODE 1: saved probabilities (B, S, S)
  __half* p_out;           // MODE 0: probabilities out; MODE 1: dS out (B, S, S)
  void* y;                 // MODE 0: attention output; MODE 1: dQ      rows (B*S), leading dimension ld_y
  int64_t ld_y;
  int* error_flag;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Row op of one 128-row block: thread = query row r (TMEM lane), column half h of the S_-wide score row held in TMEM at
// `tsc` (this thread's lane, first column of the block's scores).  MODE 0: P = softmax(scale S); MODE 1: dS = scale P (dP -
// rowsum(dP P)) with P re-read from the saved probabilities.  The fp16 result goes to shared memory in the UMMA K-major
// SWIZZLE_128B layout (element (r, key j) at  smem_p + (j / 64) * 16 KB + r * 128 + (((j % 64) / 8) ^ (r & 7)) * 16 +
// (j % 8) * 2) and to global.  xch: float[2][2][128] exchange of the row max / sum (or delta) between the two halves;
// both `bar.sync 1, 256` are executed by all 256 row-op threads.
template <int S_, int MODE>
__device__ __forceinline__ void attn_row_op(uint32_t tsc, uint8_t* smem_p, float* xch, int r, int h, int64_t grow,
                                            const AttnParams& p) {
  constexpr int NC = S_ / 2;
  float* xm = xch + h * 128 + r;
  float* xs = xch + 256 + h * 128 + r;
  const float* xm_o = xch + (h ^ 1) * 128 + r;
  __half* prow_out = p.p_out + grow * S_ + h * NC;
  uint8_t* prow_s = smem_p + r * 128;
  const uint32_t trow = tsc - h * NC;   // the code below adds h * NC itself
  if (MODE == 0) {
    uint32_t raw[NC];
#pragma unroll
    for (int c = 0; c < NC; c += 32) tmem_ld32_nowait(trow + h * NC + c, raw + c);
    tmem_wait_ld();
    float v[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = __uint_as_float(raw[j]);
    float m = v[0];
#pragma unroll
    for (int j = 1; j < NC; ++j) m = fmaxf(m, v[j]);
    *xm = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    m = fmaxf(m, *xm_o);
    const float sc = p.scale * 1.4426950408889634f;
    const float ms = m * sc;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      v[j] = ex2_approx(fmaf(v[j], sc, -ms));
      sum += v[j];
    }
    *xs = sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.0f / (xch[256 + r] + xch[256 + 128 + r]);   // same association in both halves
#pragma unroll
    for (int j = 0; j < NC; j += 8) {
      uint4 u;
      u.x = pack_f16x2(v[j] * inv, v[j + 1] * inv);
      u.y = pack_f16x2(v[j + 2] * inv, v[j + 3] * inv);
      u.z = pack_f16x2(v[j + 4] * inv, v[j + 5] * inv);
      u.w = pack_f16x2(v[j + 6] * inv, v[j + 7] * inv);
      const int key = h * NC + j;
      *reinterpret_cast<uint4*>(prow_s + (key >> 6) * 16384 + ((((key & 63) >> 3) ^ (r & 7)) << 4)) = u;
      *reinterpret_cast<uint4*>(prow_out + j) = u;
    }
  } else {
    const __half* prow_in = p.probs_in + grow * S_ + h * NC;
    float dsum = 0.f;
#pragma unroll 1
    for (int c = 0; c < NC; c += 32) {
      uint32_t v[32];
      tmem_ld32_nowait(trow + h * NC + c, v);
      uint4 pu[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pu[i] = *reinterpret_cast<const uint4*>(prow_in + c + i * 8);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __half2* hp = reinterpret_cast<const __half2*>(&pu[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(hp[k]);
          dsum = fmaf(f.x, __uint_as_float(v[i * 8 + 2 * k]), dsum);
          dsum = fmaf(f.y, __uint_as_float(v[i * 8 + 2 * k + 1]), dsum);
        }
      }
    }
    *xm = dsum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float delta = xch[r] + xch[128 + r];
#pragma unroll 1
    for (int c = 0; c < NC; c += 32) {
      uint32_t v[32];
      tmem_ld32_nowait(trow + h * NC + c, v);
      uint4 pu[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pu[i] = *reinterpret_cast<const uint4*>(prow_in + c + i * 8);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __half2* hp = reinterpret_cast<const __half2*>(&pu[i]);
        float d[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(hp[k]);
          d[2 * k] = f.x * (__uint_as_float(v[i * 8 + 2 * k]) - delta) * p.scale;
          d[2 * k + 1] = f.y * (__uint_as_float(v[i * 8 + 2 * k + 1]) - delta) * p.scale;
        }
        uint4 u;
        u.x = pack_f16x2(d[0], d[1]); u.y = pack_f16x2(d[2], d[3]); u.z = pack_f16x2(d[4], d[5]); u.w = pack_f16x2(d[6], d[7]);
        const int key = h * NC + c + i * 8;
        *reinterpret_cast<uint4*>(prow_s + (key >> 6) * 16384 + ((((key & 63) >> 3) ^ (r & 7)) << 4)) = u;
        *reinterpret_cast<uint4*>(prow_out + c + i * 8) = u;
      }
    }
  }
}

// S_ = keys per image (128 or 256), CW = C / 2 (epilogue column window per warp), MODE 0 = forward, 1 = backward (dQ)
template <int S_, int CW, int MODE>
__global__ void __launch_bounds__(320, 1) umma_attn_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB1,
                                                           const __grid_constant__ CUtensorMap tmB2,
                                                           const AttnParams p) {
  constexpr int NC = S_ / 2;   // score columns per row-op thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_p = smem + AT_P_OFFSET;
  float* xch = reinterpret_cast<float*>(smem + AT_XCH_OFFSET);   // [2 quantities][2 halves][128 rows]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + AT_BAR_OFFSET);
  uint64_t* empty = full + AT_STAGES;
  uint64_t* s_full = empty + AT_STAGES;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, b = blockIdx.y;
  const int row0 = b * S_ + qb * 128;      // first query row of this CTA in the (B*S)-row matrices

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB1);
    prefetch_tmap(&tmB2);
    for (int s = 0; s < AT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 8);
    mbar_init(o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer: nkb1 phase-1 stages, then nkb2 phase-2 stages through the same ring =====
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int kb = 0; kb < p.nkb1 && ok; ++kb) {
        ok = mbar_wait(&empty[st], ph ^ 1, p.error_flag, 11);
        if (!ok) break;
        uint8_t* sa = smem + st * AT_STAGE_BYTES;
        mbar_expect_tx(&full[st], AT_A_BYTES + S_ * 128);
        tma_load_3d(&tmA, &full[st], sa, kb * 64, row0, 0);
        tma_load_3d(&tmB1, &full[st], sa + AT_A_BYTES, kb * 64, b * S_, 0);
        if (++st == AT_STAGES) { st = 0; ph ^= 1; }
      }
      const int nslab = p.C / 64;
      for (int kb = 0; kb < p.nkb2 && ok; ++kb) {
        ok = mbar_wait(&empty[st], ph ^ 1, p.error_flag, 11);
        if (!ok) break;
        uint8_t* sb = smem + st * AT_STAGE_BYTES;
        mbar_expect_tx(&full[st], (uint32_t)nslab * 8192u);
        for (int j = 0; j < nslab; ++j)    // MN-major operand: [64 keys][64 channels] slabs, 8 KB apart
          tma_load_3d(&tmB2, &full[st], sb + j * 8192, j * 64, b * S_ + kb * 64, 0);
        if (++st == AT_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int kb = 0; kb < p.nkb1 && ok; ++kb) {
        ok = mbar_wait(&full[st], ph, p.error_flag, 12);
        if (!ok) break;
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + st * AT_STAGE_BYTES), sb = sa + AT_A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base, make_desc(sa + k * 32, 1, 64), make_desc(sb + k * 32, 1, 64), p.idesc1, (kb | k) ? 1u : 0u);
        umma_commit(&empty[st]);
        if (++st == AT_STAGES) { st = 0; ph ^= 1; }
      }
      if (ok) umma_commit(s_full);
      if (ok) ok = mbar_wait(p_ready, 0, p.error_flag, 13);   // P / dS tile written by the row-op warps
      tc_fence_after();
      const uint32_t sp = smem_u32(smem_p);
      for (int kb = 0; kb < p.nkb2 && ok; ++kb) {
        ok = mbar_wait(&full[st], ph, p.error_flag, 12);
        if (!ok) break;
        tc_fence_after();
        const uint32_t sb = smem_u32(smem + st * AT_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + 256, make_desc(sp + kb * 16384 + k * 32, 1, 64), make_desc(sb + k * 2048, 512, 64), p.idesc2,
                   (kb | k) ? 1u : 0u);
        umma_commit(&empty[st]);
        if (++st == AT_STAGES) { st = 0; ph ^= 1; }
      }
      if (ok) umma_commit(o_full);
    }
  } else {
    // ===== row op + epilogue: thread = query row r (TMEM lane), column half h of the score row =====
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t grow = (int64_t)row0 + r;   // global row in (B*S)
    bool ok = mbar_wait(s_full, 0, p.error_flag, 14);
    tc_fence_after();
    if (ok) {
      attn_row_op<S_, MODE>(trow + h * NC, smem_p, xch, r, h, grow, p);
      fence_proxy_async();    // generic-proxy writes of the P tile -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // ---- epilogue: O / dQ accumulator (TMEM column 256..) -> fp16 rows
      ok = mbar_wait(o_full, 0, p.error_flag, 15);
      tc_fence_after();
      if (ok) {
        float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * 32 * (CW + 4);   // ring memory: every MMA has retired
        EpiArgs e{nullptr, nullptr, nullptr, 0, 1.0f, p.y, p.ld_y, 0};
        epilogue_warp<CW>(trow + 256 + h * CW, stage, lane, grow, grow, true, h * CW, e, nullptr, 0, 1);
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

// =============================================================================================================
// Per-image variant for S = 256 ("attn2", the default): ONE CTA owns BOTH 128-query blocks of an image.
//   * K / V cross L2 -> SM once per image instead of once per query block: a phase-1 ring entry holds the k-block of Q_A,
//     Q_B and K (64 KB) and feeds two MMAs chains (S_A in TMEM columns 0..255, S_B in 256..511); a phase-2 entry holds one
//     64-key block of V and feeds O_A and O_B;
//   * the output accumulators REUSE the score columns (O_A over S_A, O_B over S_B): the row op has consumed the scores
//     before p_ready lets the first P V MMA write there, so 2 x (256 + 0) = 512 columns serve both blocks;
//   * 128 CTAs for B = 128 = one wave on 148 SMs (the one-tile kernel ran 256 CTAs in two waves, every CTA a serial chain
//     load -> QK^T -> row op -> PV -> epilogue of ~23 kcycles with nothing to overlap it).
// Shared memory (192 KB): three 64 KB regions R0 | R1 | R2.  Phase 1: a 3-entry operand ring.  After the last score MMA
// retires R1 / R2 become P_A / P_B (written by the row op in the UMMA K-major layout) and R0 the 2-entry V ring; after
// the last P V MMA everything is epilogue staging.
// =============================================================================================================
constexpr int A2_REGION = 64 * 1024;
constexpr int A2_XCH_OFFSET = 3 * A2_REGION;
constexpr int A2_BAR_OFFSET = A2_XCH_OFFSET + 2 * 2 * 128 * 4;
constexpr int A2_SMEM = A2_BAR_OFFSET + (3 + 3 + 1 + 2 + 2 + 2 + 1) * 8 + 16 + 1024;
static_assert(A2_SMEM <= 232448, "attention (per-image) shared memory");

template <int CW, int MODE>
__global__ void __launch_bounds__(320, 1) umma_attn2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB1,
                                                            const __grid_constant__ CUtensorMap tmB2,
                                                            const AttnParams p) {
  constexpr int S_ = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* xch = reinterpret_cast<float*>(smem + A2_XCH_OFFSET);
  uint64_t* full1 = reinterpret_cast<uint64_t*>(smem + A2_BAR_OFFSET);
  uint64_t* empty1 = full1 + 3;
  uint64_t* s_full = empty1 + 3;
  uint64_t* p_ready = s_full + 1;    // [2]: P_A / P_B written
  uint64_t* full2 = p_ready + 2;     // [2]
  uint64_t* empty2 = full2 + 2;      // [2]
  uint64_t* o_full = empty2 + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = b * S_;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB1);
    prefetch_tmap(&tmB2);
    for (int s = 0; s < 3; ++s) { mbar_init(&full1[s], 1); mbar_init(&empty1[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&full2[s], 1); mbar_init(&empty2[s], 1); mbar_init(&p_ready[s], 8); }
    mbar_init(s_full, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      bool ok = true;
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < p.nkb1 && ok; ++kb) {
        ok = mbar_wait(&empty1[st], ph ^ 1, p.error_flag, 21);
        if (!ok) break;
        uint8_t* sq = smem + st * A2_REGION;
        mbar_expect_tx(&full1[st], 2u * AT_A_BYTES + S_ * 128);
        tma_load_3d(&tmA, &full1[st], sq, kb * 64, row0, 0);
        tma_load_3d(&tmA, &full1[st], sq + AT_A_BYTES, kb * 64, row0 + 128, 0);
        tma_load_3d(&tmB1, &full1[st], sq + 2 * AT_A_BYTES, kb * 64, row0, 0);
        if (++st == 3) { st = 0; ph ^= 1; }
      }
      // the V ring lives in R0, which the score MMAs read until the last of them has retired
      if (ok) ok = mbar_wait(s_full, 0, p.error_flag, 22);
      const int nslab = p.C / 64;
      st = 0; ph = 0;
      for (int kb = 0; kb < p.nkb2 && ok; ++kb) {
        ok = mbar_wait(&empty2[st], ph ^ 1, p.error_flag, 21);
        if (!ok) break;
        uint8_t* sv = smem + st * 32768;
        mbar_expect_tx(&full2[st], (uint32_t)nslab * 8192u);
        for (int j = 0; j < nslab; ++j) tma_load_3d(&tmB2, &full2[st], sv + j * 8192, j * 64, row0 + kb * 64, 0);
        if (++st == 2) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      bool ok = true;
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < p.nkb1 && ok; ++kb) {
        ok = mbar_wait(&full1[st], ph, p.error_flag, 23);
        if (!ok) break;
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + st * A2_REGION), sk = sq + 2 * AT_A_BYTES;
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + blk * 256, make_desc(sq + blk * AT_A_BYTES + k * 32, 1, 64), make_desc(sk + k * 32, 1, 64),
                     p.idesc1, (kb | k) ? 1u : 0u);
        umma_commit(&empty1[st]);
        if (++st == 3) { st = 0; ph ^= 1; }
      }
      if (ok) umma_commit(s_full);
      // both P tiles in place (and with them the guarantee that every score has been read out of TMEM)
      if (ok) ok = mbar_wait(&p_ready[0], 0, p.error_flag, 24);
      if (ok) ok = mbar_wait(&p_ready[1], 0, p.error_flag, 24);
      tc_fence_after();
      st = 0; ph = 0;
      for (int kb = 0; kb < p.nkb2 && ok; ++kb) {
        ok = mbar_wait(&full2[st], ph, p.error_flag, 23);
        if (!ok) break;
        tc_fence_after();
        const uint32_t sv = smem_u32(smem + st * 32768);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const uint32_t sp = smem_u32(smem + (1 + blk) * A2_REGION);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + blk * 256, make_desc(sp + kb * 16384 + k * 32, 1, 64), make_desc(sv + k * 2048, 512, 64),
                     p.idesc2, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty2[st]);
        if (++st == 2) { st = 0; ph ^= 1; }
      }
      if (ok) umma_commit(o_full);
    }
  } else {
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    bool ok = mbar_wait(s_full, 0, p.error_flag, 25);
    tc_fence_after();
    if (ok) {
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {
        attn_row_op<S_, MODE>(trow + blk * 256 + h * (S_ / 2), smem + (1 + blk) * A2_REGION, xch, r, h, (int64_t)row0 + blk * 128 + r, p);
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[blk]);
      }
      ok = mbar_wait(o_full, 0, p.error_flag, 26);
      tc_fence_after();
      if (ok) {
        float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * 32 * (CW + 4);   // every MMA has retired: all three regions are free
        EpiArgs e{nullptr, nullptr, nullptr, 0, 1.0f, p.y, p.ld_y, 0};
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
          const int64_t grow = (int64_t)row0 + blk * 128 + r;
          epilogue_warp<CW>(trow + blk * 256 + h * CW, stage, lane, grow, grow, true, h * CW, e, nullptr, 0, 1);
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

template <int CW, int MODE>
static int attn2_launch_t(const CUtensorMap& ma, const CUtensorMap& mb1, const CUtensorMap& mb2, const AttnParams& p, int B,
                          cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(umma_attn2_kernel<CW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM);
    attr_set = true;
  }
  cudaError_t e = launch_pdl(umma_attn2_kernel<CW, MODE>, dim3(1, B), dim3(320), A2_SMEM, st, ma, mb1, mb2, p);
  if (e != cudaSuccess) { set_error("fused attention (per-image) launch failed: %s", cudaGetErrorString(e)); return BD_ERR_CUDA; }
  count_launch(1);
  return BD_OK;
}

// =============================================================================================================
// dK / dV of the same attention in ONE launch (the batched weight-gradient GEMMs they used to be ran 512 one-tile CTAs
// with four k-blocks each -- bound by the ~10 us fixed cost of a one-tile CTA, 45 us per GEMM at B = 128):
// one CTA per (image, 128-key block) accumulates BOTH products over the S queries in the two halves of TMEM,
//     dV[keys, :] = sum_q P[q, keys]^T dO[q, :]          dK[keys, :] = sum_q dS[q, keys]^T Q[q, :]
// A = the 64-query x 128-key block of P / dS (keys contiguous: MN-major A), B = 64 queries x C channels of dO / Q
// (channels contiguous: MN-major B), all read in place by TMA; 48 KB ring entries (16 KB A + 32 KB B), 4 stages.
// =============================================================================================================
constexpr int DKV_STAGES = 4;
constexpr int DKV_BAR_OFFSET = DKV_STAGES * AT_STAGE_BYTES;
constexpr int DKV_SMEM = DKV_BAR_OFFSET + (2 * DKV_STAGES + 1) * 8 + 16 + 1024;
static_assert(DKV_SMEM <= 232448, "attention dK/dV shared memory");

struct AttnDkvParams {
  int S, C, nkb;          // nkb = S / 64 query blocks
  uint32_t idesc;
  void* dv; void* dk;     // (B*S, C) rows with leading dimension ld_y
  int64_t ld_y;
  int* error_flag;
};

template <int CW>
__global__ void __launch_bounds__(320, 1) umma_attn_dkv_kernel(const __grid_constant__ CUtensorMap tmP,
                                                               const __grid_constant__ CUtensorMap tmDO,
                                                               const __grid_constant__ CUtensorMap tmDS,
                                                               const __grid_constant__ CUtensorMap tmQ,
                                                               const AttnDkvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + DKV_BAR_OFFSET);
  uint64_t* empty = full + DKV_STAGES;
  uint64_t* o_full = empty + DKV_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblk = blockIdx.x, b = blockIdx.y;
  const int key0 = kblk * 128;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmP);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDS);
    prefetch_tmap(&tmQ);
    for (int s = 0; s < DKV_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(o_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int nslab = p.C / 64;
  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int e = 0; e < 2 * p.nkb && ok; ++e) {
        const int g = e / p.nkb, kb = e - g * p.nkb;
        const CUtensorMap* mA = g ? &tmDS : &tmP;
        const CUtensorMap* mB = g ? &tmQ : &tmDO;
        ok = mbar_wait(&empty[st], ph ^ 1, p.error_flag, 16);
        if (!ok) break;
        uint8_t* sa = smem + st * AT_STAGE_BYTES;
        mbar_expect_tx(&full[st], 2u * 8192u + (uint32_t)nslab * 8192u);
        const int qrow = b * p.S + kb * 64;
        tma_load_3d(mA, &full[st], sa, key0, qrow, 0);
        tma_load_3d(mA, &full[st], sa + 8192, key0 + 64, qrow, 0);
        for (int j = 0; j < nslab; ++j) tma_load_3d(mB, &full[st], sa + AT_A_BYTES + j * 8192, j * 64, qrow, 0);
        if (++st == DKV_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int e = 0; e < 2 * p.nkb && ok; ++e) {
        const int g = e / p.nkb, kb = e - g * p.nkb;
        ok = mbar_wait(&full[st], ph, p.error_flag, 17);
        if (!ok) break;
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + st * AT_STAGE_BYTES), sb = sa + AT_A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + g * 256, make_desc(sa + k * 2048, 512, 64), make_desc(sb + k * 2048, 512, 64), p.idesc,
                   (kb | k) ? 1u : 0u);
        umma_commit(&empty[st]);
        if (++st == DKV_STAGES) { st = 0; ph ^= 1; }
      }
      if (ok) umma_commit(o_full);
    }
  } else {
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int64_t grow = (int64_t)b * p.S + key0 + q * 32 + lane;
    const bool ok = mbar_wait(o_full, 0, p.error_flag, 18);
    tc_fence_after();
    if (ok) {
      float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * 32 * (CW + 4);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        EpiArgs e{nullptr, nullptr, nullptr, 0, 1.0f, g ? p.dk : p.dv, p.ld_y, 0};
        epilogue_warp<CW>(trow + g * 256 + h * CW, stage, lane, grow, grow, true, h * CW, e, nullptr, 0, 1);
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

template <int CW>
static int attn_dkv_launch_t(const CUtensorMap& mp, const CUtensorMap& mdo, const CUtensorMap& mds, const CUtensorMap& mq,
                             const AttnDkvParams& p, int B, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(umma_attn_dkv_kernel<CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM);
    attr_set = true;
  }
  cudaError_t e = launch_pdl(umma_attn_dkv_kernel<CW>, dim3(p.S / 128, B), dim3(320), DKV_SMEM, st, mp, mdo, mds, mq, p);
  if (e != cudaSuccess) { set_error("fused attention dK/dV launch failed: %s", cudaGetErrorString(e)); return BD_ERR_CUDA; }
  count_launch(1);
  return BD_OK;
}

// probs, ds: (B, S, S) fp16; d_out (B*S, ld_dout); q (B*S, ld_qkv); dk, dv: column slices of the (B*S, ld_y) gradient buffer
int attn_dkv_launch(const void* probs, const void* ds, const void* d_out, int64_t ld_dout, const void* q, int64_t ld_qkv,
                    void* dk, void* dv, int64_t ld_y, int B, int S, int C, cudaStream_t st) {
  AttnDkvParams p;
  memset(&p, 0, sizeof(p));
  p.S = S; p.C = C; p.nkb = S / 64;
  p.idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.dv = dv; p.dk = dk; p.ld_y = ld_y;
  p.error_flag = error_flag();
  CUtensorMap mp, mdo, mds, mq;
  const uint64_t rows = (uint64_t)B * S;
  const uint32_t box[3] = {64, 64, 1};
  {
    uint64_t dims[3] = {(uint64_t)S, rows, 1};
    uint64_t str[2] = {(uint64_t)S, rows * (uint64_t)S};
    if (!make_map(&mp, probs, 3, dims, str, box)) return BD_ERR_CUDA;
    if (!make_map(&mds, ds, 3, dims, str, box)) return BD_ERR_CUDA;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, rows, 1};
    uint64_t str[2] = {(uint64_t)ld_dout, rows * (uint64_t)ld_dout};
    if (!make_map(&mdo, d_out, 3, dims, str, box)) return BD_ERR_CUDA;
    uint64_t str2[2] = {(uint64_t)ld_qkv, rows * (uint64_t)ld_qkv};
    if (!make_map(&mq, q, 3, dims, str2, box)) return BD_ERR_CUDA;
  }
  if (C == 256) return attn_dkv_launch_t<128>(mp, mdo, mds, mq, p, B, st);
  if (C == 128) return attn_dkv_launch_t<64>(mp, mdo, mds, mq, p, B, st);
  if (C == 64) return attn_dkv_launch_t<32>(mp, mdo, mds, mq, p, B, st);
  set_error("fused attention dK/dV: unsupported C=%d", C);
  return BD_ERR_UNSUPPORTED;
}

bool attn_fused_supported(int S, int C, int heads, int64_t ld_a, int64_t ld_qkv) {
  if (getenv("BD_NO_ATTN_FUSED")) return false;
  return heads == 1 && (S == 128 || S == 256) && (C == 64 || C == 128 || C == 256) && ld_a % 8 == 0 && ld_qkv % 8 == 0;
}

template <int S_, int CW, int MODE>
static int attn_launch_t(const CUtensorMap& ma, const CUtensorMap& mb1, const CUtensorMap& mb2, const AttnParams& p, int B,
                         cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(umma_attn_kernel<S_, CW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    attr_set = true;
  }
  cudaError_t e = launch_pdl(umma_attn_kernel<S_, CW, MODE>, dim3(S_ / 128, B), dim3(320), AT_SMEM, st, ma, mb1, mb2, p);
  if (e != cudaSuccess) { set_error("fused attention launch failed: %s", cudaGetErrorString(e)); return BD_ERR_CUDA; }
  count_launch(1);
  return BD_OK;
}

// a: phase-1 A operand rows (Q forward, dO backward), (B*S, C) with leading dimension ld_a
// b1: phase-1 B operand (K forward, V backward) and b2: phase-2 B operand (V forward, K backward): column slices of the
//     (B*S, ld_qkv) projection buffer
int attn_fused_launch(int mode, const void* a, int64_t ld_a, const void* b1, const void* b2, int64_t ld_qkv,
                      const void* probs_in, void* p_out, void* y, int64_t ld_y, int B, int S, int C, float scale,
                      cudaStream_t st) {
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.S = S; p.C = C; p.nkb1 = C / 64; p.nkb2 = S / 64;
  p.idesc1 = (1u << 4) | ((uint32_t)(S >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.scale = scale;
  p.probs_in = (const __half*)probs_in; p.p_out = (__half*)p_out; p.y = y; p.ld_y = ld_y;
  p.error_flag = error_flag();
  CUtensorMap ma, mb1, mb2;
  const uint64_t rows = (uint64_t)B * S;
  {
    uint64_t dims[3] = {(uint64_t)C, rows, 1};
    uint64_t str[2] = {(uint64_t)ld_a, rows * (uint64_t)ld_a};
    uint32_t box[3] = {64, 128, 1};
    if (!make_map(&ma, a, 3, dims, str, box)) return BD_ERR_CUDA;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, rows, 1};
    uint64_t str[2] = {(uint64_t)ld_qkv, rows * (uint64_t)ld_qkv};
    uint32_t box1[3] = {64, (uint32_t)S, 1};
    if (!make_map(&mb1, b1, 3, dims, str, box1)) return BD_ERR_CUDA;
    uint32_t box2[3] = {64, 64, 1};
    if (!make_map(&mb2, b2, 3, dims, str, box2)) return BD_ERR_CUDA;
  }
  if (S == 256 && !getenv("BD_ATTN_V1")) {
#define BD_ATTN2_CASE(CC)                                                                                          \
  if (C == CC) return mode == 0 ? attn2_launch_t<CC / 2, 0>(ma, mb1, mb2, p, B, st) : attn2_launch_t<CC / 2, 1>(ma, mb1, mb2, p, B, st);
    BD_ATTN2_CASE(256)
    BD_ATTN2_CASE(128)
    BD_ATTN2_CASE(64)
#undef BD_ATTN2_CASE
  }
#define BD_ATTN_CASE(SS, CC)                                                                          \
  if (S == SS && C == CC)                                                                             \
    return mode == 0 ? attn_launch_t<SS, CC / 2, 0>(ma, mb1, mb2, p, B, st) : attn_launch_t<SS, CC / 2, 1>(ma, mb1, mb2, p, B, st);
  BD_ATTN_CASE(256, 256)
  BD_ATTN_CASE(256, 128)
  BD_ATTN_CASE(256, 64)
  BD_ATTN_CASE(128, 256)
  BD_ATTN_CASE(128, 128)
  BD_ATTN_CASE(128, 64)
#undef BD_ATTN_CASE
  set_error("fused attention: unsupported shape S=%d C=%d", S, C);
  return BD_ERR_UNSUPPORTED;
}

}  // namespace umma
}  // namespace bd
