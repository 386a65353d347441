// Shared pieces of the tcgen05 kernels: PTX wrappers (mbarrier, TMA, TMEM, UMMA), descriptor builders and the
// host-side tensor-map helpers.
#pragma once
#include "common.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace bd {
namespace umma {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* error_flag, int code) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) {  // ~1 s
      if (error_flag) atomicExch(error_flag, code);
      return false;
    }
  }
  return true;
}
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor is still draining; everything before pdl_wait() (barrier init, TMEM allocation, tensor-map
// prefetch) overlaps the predecessor's tail, everything that touches global memory comes after it.
// launch_dependents lets the NEXT kernel's CTAs take SMs as this kernel's CTAs retire.  Both are no-ops for a
// normally launched kernel.  Rule: every kernel launched through launch_pdl() executes pdl_wait() in all threads.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Multicast variant: the box lands at the SAME CTA-relative shared-memory offset in every CTA of `cta_mask`, and each
// destination CTA's own mbarrier (same offset) receives the complete_tx for the bytes written into it.
__device__ __forceinline__ void tma_load_3d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6}], [%2], %3;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(cta_mask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the arrive lands on the mbarrier at the same CTA-relative offset in every CTA of `cta_mask` (this CTA included)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// issue only: pair with tmem_wait_ld() before the registers are read (lets global loads overlap the TMEM read)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// 16 TMEM lanes x 16 columns in the mma-fragment layout (cute SM100_TMEM_LOAD_16dp256b2x): with L = lane of the
// address, thread t receives  v[0,1] = (lane L + t/4,     columns 2(t%4), 2(t%4)+1),  v[2,3] = (lane L + t/4 + 8, same),
//                             v[4,5] = (lane L + t/4, columns 8 + 2(t%4), +1),        v[6,7] = (lane L + t/4 + 8, same).
// Every {v[2i], v[2i+1]} pair is one stmatrix fragment register once packed to f16x2.
__device__ __forceinline__ void tmem_ld_16x256b_x2_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
// four 8x8 b16 matrices, stored TRANSPOSED: fragment element (row r, column c) of matrix i lands in the 16-byte
// shared-memory row addressed by thread 8i + c, at position r.
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t smem_row_addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};"
               ::"r"(smem_row_addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo16, uint32_t sbo16) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) |
         (1ull << 46) | (2ull << 61);
}


// ---------------------------------------------------------------------------------------------
// Shared epilogue: one warp drains 32 accumulator rows x BN columns.
//   phase 1  TMEM -> registers (tcgen05.ld) -> + bias (+bias2, +per-sample rowbias) -> fp32 staging tile in smem
//            (row pitch BN+4 floats: conflict-free 16-byte st.shared per quarter warp);
//   phase 2  coalesced read-back: 8 channels per lane, BN/8 lanes per pixel row, so every global load (residual)
//            and store is a run of full 32-byte sectors of ONE pixel row instead of 32 different rows.
// A per-thread strided epilogue measured 35-43 kcycles per CTA in the 3x3 kernel (LSU wavefront bound); this is the fix.
// ---------------------------------------------------------------------------------------------
struct EpiArgs {
  const float* bias;
  const float* bias2;
  const __half* residual;
  int64_t ld_res;
  float scale;
  void* y;
  int64_t ld_y;
  int out_f32;
  // GroupNorm statistics of the output (GNS instantiations of the epilogues only): gn_sums[sample * ld_sums + 2 * column
  // + {0, 1}] += sum / sum of squares of the fp16-rounded outputs; sample = dense pixel index / HW
  float* gn_sums = nullptr;
  int64_t ld_sums = 0;
};

// Fold the per-lane channel partials of one sample over the lanes that hold the same 8 columns (lane = rsub * LPR + piece)
// and add them to the statistics buffer (one red.global.add per (warp, sample, column, quantity)).  Warp-uniform call.
template <int LPR>
__device__ __forceinline__ void gn_sums_flush(float* gs1, float* gs2, float* dst, int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      gs1[k] += __shfl_xor_sync(0xffffffffu, gs1[k], o);
      gs2[k] += __shfl_xor_sync(0xffffffffu, gs2[k], o);
    }
  }
  if (lane < LPR && dst) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(dst + 2 * k, gs1[k]);
      atomicAdd(dst + 2 * k + 1, gs2[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) gs1[k] = gs2[k] = 0.f;
}

// Residual operand of the rows / columns this lane will finish in epilogue_warp<CW>, fetched BEFORE the warp waits for
// the accumulator: issued inside the row loop, each of the CW/8 iterations paid one dependent HBM round trip (the
// attention proj GEMM ran 7.8 us per tile against 2.9 us for the same tile without a residual).
template <int CW>
struct EpiResidual {
  half8 v[CW / 8];
};
template <int CW>
__device__ __forceinline__ void epilogue_prefetch_residual(EpiResidual<CW>& r, int lane, int64_t m_own, bool valid_own, int col0,
                                                           const EpiArgs& e) {
  constexpr int LPR = CW / 8, RPI = 32 / LPR;
  const int piece = lane % LPR, rsub = lane / LPR;
#pragma unroll
  for (int it = 0; it < LPR; ++it) {
    const int row = it * RPI + rsub;
    const int64_t m = __shfl_sync(0xffffffffu, m_own, row);
    const int valid = __shfl_sync(0xffffffffu, (int)valid_own, row);
    if (valid) r.v[it] = *reinterpret_cast<const half8*>(e.residual + m * e.ld_res + col0 + piece * 8);
  }
}

// CW = number of accumulator columns this warp drains (a window of the tile starting at TMEM address `taddr`,
// global column `col0`).  rowbias: base pointer (nullable); the per-row sample index is mlin / HW.
// pre: optional residual values fetched by epilogue_prefetch_residual (nullptr: loaded in the loop).
template <int CW, bool GNS = false>
__device__ __forceinline__ void epilogue_warp(uint32_t taddr, float* __restrict__ stage, int lane, int64_t m_own,
                                              int64_t mlin_own, bool valid_own, int col0, const EpiArgs& e,
                                              const float* __restrict__ rowbias, int64_t ld_rowbias, int HW,
                                              const EpiResidual<CW>* pre = nullptr) {
  constexpr int PITCH = CW + 4;
  float* myrow = stage + lane * PITCH;
#pragma unroll
  for (int c0 = 0; c0 < CW; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(myrow + c0 + j) =
          make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
  }
  __syncwarp();
  constexpr int LPR = CW / 8;        // lanes per pixel row
  constexpr int RPI = 32 / LPR;      // rows per iteration
  const int piece = lane % LPR, rsub = lane / LPR;
  const int col = col0 + piece * 8;
  float bsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // bias (+bias2) of this lane's 8 columns: loaded once
  if (e.bias) {
    const float4 a = *reinterpret_cast<const float4*>(e.bias + col), b = *reinterpret_cast<const float4*>(e.bias + col + 4);
    bsum[0] += a.x; bsum[1] += a.y; bsum[2] += a.z; bsum[3] += a.w; bsum[4] += b.x; bsum[5] += b.y; bsum[6] += b.z; bsum[7] += b.w;
  }
  if (e.bias2) {
    const float4 a = *reinterpret_cast<const float4*>(e.bias2 + col), b = *reinterpret_cast<const float4*>(e.bias2 + col + 4);
    bsum[0] += a.x; bsum[1] += a.y; bsum[2] += a.z; bsum[3] += a.w; bsum[4] += b.x; bsum[5] += b.y; bsum[6] += b.z; bsum[7] += b.w;
  }
  // GNS: the RPI rows of one iteration belong to ONE sample (hosts enable it only for HW % 8 == 0 and tiles that hold whole
  // pixel groups of an image), so the sample index is warp-uniform and a change of it flushes the partials of the previous one
  float gs1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, gs2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int cur_smp = -1;
  // sample index of this lane's own row, once (a 64-bit division per row and iteration otherwise)
  const int smp_own = ((GNS || rowbias) && valid_own) ? (int)(mlin_own / HW) : -1;
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += RPI) {
    const int row = r0 + rsub;
    const int64_t m = __shfl_sync(0xffffffffu, m_own, row);
    const int valid = __shfl_sync(0xffffffffu, (int)valid_own, row);
    const int smp_row = __shfl_sync(0xffffffffu, smp_own, row);
    if (GNS) {
      const int smp = __shfl_sync(0xffffffffu, smp_row, 0);   // lane 0's row speaks for the iteration
      if (smp != cur_smp) {
        if (cur_smp >= 0) gn_sums_flush<LPR>(gs1, gs2, e.gn_sums + (int64_t)cur_smp * e.ld_sums + 2 * col, lane);
        cur_smp = smp;
      }
    }
    if (!valid) continue;
    const float* sp = stage + row * PITCH + piece * 8;
    const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 4);
    float f[8] = {a.x + bsum[0], a.y + bsum[1], a.z + bsum[2], a.w + bsum[3], b.x + bsum[4], b.y + bsum[5], b.z + bsum[6], b.w + bsum[7]};
    if (rowbias) {
      const float* rb = rowbias + (int64_t)smp_row * ld_rowbias + col;
      const float4 ra = *reinterpret_cast<const float4*>(rb), rc = *reinterpret_cast<const float4*>(rb + 4);
      f[0] += ra.x; f[1] += ra.y; f[2] += ra.z; f[3] += ra.w; f[4] += rc.x; f[5] += rc.y; f[6] += rc.z; f[7] += rc.w;
    }
    if (e.residual) {
      float g[8];
      if (pre) unpack8(pre->v[r0 / RPI], g);
      else unpack8(*reinterpret_cast<const half8*>(e.residual + m * e.ld_res + col), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += g[k];
    }
    if (e.scale != 1.0f) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] *= e.scale;
    }
    if (e.out_f32) {
      float* yr = reinterpret_cast<float*>(e.y) + m * e.ld_y + col;
      *reinterpret_cast<float4*>(yr) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(yr + 4) = make_float4(f[4], f[5], f[6], f[7]);
    } else {
      const half8 hv = pack8(f);
      *reinterpret_cast<half8*>(reinterpret_cast<__half*>(e.y) + m * e.ld_y + col) = hv;
      if (GNS) {   // statistics of the rounded values the consumer will read
        float g[8];
        unpack8(hv, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) { gs1[k] += g[k]; gs2[k] = fmaf(g[k], g[k], gs2[k]); }
      }
    }
  }
  if (GNS) {
    __syncwarp();
    if (cur_smp >= 0) gn_sums_flush<LPR>(gs1, gs2, e.gn_sums + (int64_t)cur_smp * e.ld_sums + 2 * col, lane);
  }
  __syncwarp();
}

// host: launch with the programmatic-dependent-launch attribute (BD_NO_PDL=1 -> plain launch)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = getenv("BD_NO_PDL") ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------
// Split-K inside a thread-block cluster: the S CTAs of a cluster (cluster dims (1,1,S)) accumulate disjoint k-block
// ranges of the SAME output tile.  Each stages its fp32 partial tile in its own shared memory (phase 1), the cluster
// meets at a barrier, and rank r finishes every S-th group of rows: it adds the S partials in rank order through
// distributed shared memory (deterministic), applies the usual epilogue and writes the result (phase 2).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_smem_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
  return v;
}

// phase 1: TMEM -> this warp's fp32 staging tile [32 rows][CW + 4]
template <int CW>
__device__ __forceinline__ void epilogue_stage_warp(uint32_t taddr, float* __restrict__ stage, int lane) {
  constexpr int PITCH = CW + 4;
  float* myrow = stage + lane * PITCH;
#pragma unroll
  for (int c0 = 0; c0 < CW; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(myrow + c0 + j) =
          make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
  }
}

// phase 2 (after the cluster barrier): rows r0 = RPI * i with i % S == rank
template <int CW, bool GNS = false>
__device__ __forceinline__ void epilogue_splitk_finish_warp(const float* __restrict__ stage, int lane, int S, int rank,
                                                            int64_t m_own, int64_t mlin_own, bool valid_own, int col0,
                                                            const EpiArgs& e, const float* __restrict__ rowbias,
                                                            int64_t ld_rowbias, int HW, int rank_stride = 1, int rank_off = 0) {
  constexpr int PITCH = CW + 4;
  constexpr int LPR = CW / 8;
  constexpr int RPI = 32 / LPR;
  const int piece = lane % LPR, rsub = lane / LPR;
  const int col = col0 + piece * 8;
  float bsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (e.bias) {
    const float4 a = *reinterpret_cast<const float4*>(e.bias + col), b = *reinterpret_cast<const float4*>(e.bias + col + 4);
    bsum[0] += a.x; bsum[1] += a.y; bsum[2] += a.z; bsum[3] += a.w; bsum[4] += b.x; bsum[5] += b.y; bsum[6] += b.z; bsum[7] += b.w;
  }
  if (e.bias2) {
    const float4 a = *reinterpret_cast<const float4*>(e.bias2 + col), b = *reinterpret_cast<const float4*>(e.bias2 + col + 4);
    bsum[0] += a.x; bsum[1] += a.y; bsum[2] += a.z; bsum[3] += a.w; bsum[4] += b.x; bsum[5] += b.y; bsum[6] += b.z; bsum[7] += b.w;
  }
  float gs1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, gs2[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // GNS: see epilogue_warp
  int64_t cur_smp = -1;
  for (int it = rank; it < 32 / RPI; it += S) {
    const int row = it * RPI + rsub;
    const int64_t m = __shfl_sync(0xffffffffu, m_own, row);
    const int64_t mlin = __shfl_sync(0xffffffffu, mlin_own, row);
    const int valid = __shfl_sync(0xffffffffu, (int)valid_own, row);
    if (GNS) {
      const int64_t smp = __shfl_sync(0xffffffffu, valid ? mlin / HW : (int64_t)-1, 0);
      if (smp != cur_smp) {
        if (cur_smp >= 0) gn_sums_flush<LPR>(gs1, gs2, e.gn_sums + cur_smp * e.ld_sums + 2 * col, lane);
        cur_smp = smp;
      }
    }
    if (!valid) continue;
    const uint32_t sp = smem_u32(stage + row * PITCH + piece * 8);
    float f[8] = {bsum[0], bsum[1], bsum[2], bsum[3], bsum[4], bsum[5], bsum[6], bsum[7]};
    for (int rk = 0; rk < S; ++rk) {
      // split index rk lives in cluster rank rank_off + rank_stride * rk (clusters that also share B tiles along x)
      const uint32_t cr = (uint32_t)(rank_off + rank_stride * rk);
      const float4 a = ld_dsmem_f4(sp, cr), b = ld_dsmem_f4(sp + 16, cr);
      f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
    }
    if (rowbias) {
      const float* rb = rowbias + (mlin / HW) * ld_rowbias + col;
      const float4 ra = *reinterpret_cast<const float4*>(rb), rc = *reinterpret_cast<const float4*>(rb + 4);
      f[0] += ra.x; f[1] += ra.y; f[2] += ra.z; f[3] += ra.w; f[4] += rc.x; f[5] += rc.y; f[6] += rc.z; f[7] += rc.w;
    }
    if (e.residual) {
      float g[8];
      unpack8(*reinterpret_cast<const half8*>(e.residual + m * e.ld_res + col), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += g[k];
    }
    if (e.scale != 1.0f) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] *= e.scale;
    }
    if (e.out_f32) {
      float* yr = reinterpret_cast<float*>(e.y) + m * e.ld_y + col;
      *reinterpret_cast<float4*>(yr) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(yr + 4) = make_float4(f[4], f[5], f[6], f[7]);
    } else {
      const half8 hv = pack8(f);
      *reinterpret_cast<half8*>(reinterpret_cast<__half*>(e.y) + m * e.ld_y + col) = hv;
      if (GNS) {
        float g[8];
        unpack8(hv, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) { gs1[k] += g[k]; gs2[k] = fmaf(g[k], g[k], gs2[k]); }
      }
    }
  }
  if (GNS) {
    __syncwarp();
    if (cur_smp >= 0) gn_sums_flush<LPR>(gs1, gs2, e.gn_sums + cur_smp * e.ld_sums + 2 * col, lane);
  }
}

// host helpers (defined in umma.cu)
bool make_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_elems,
              const uint32_t* box, const uint32_t* elem_strides = nullptr);
int* error_flag();
uint32_t env_u32(const char* name, uint32_t dflt);

}  // namespace umma
}  // namespace bd
