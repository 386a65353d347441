// Microbenchmark: tcgen05.mma SS-mode issue throughput for M=128, N in {64,128,256}, K=16 fp16, operands resident in
// shared memory (SWIZZLE_128B K-major), accumulating into 1..4 TMEM accumulators.  Prints cycles per MMA.
#include "../../baddiffusion_b200/csrc/umma_common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace bd::umma;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int reps, int naccum, int kadv, int b_mn, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); fence_proxy_async(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const int k = kadv ? (i & 3) : 0;
      const uint64_t ad = make_desc(sa + k * 32, 1, 64);
      const uint64_t bd = b_mn ? make_desc(sb + k * 2048, 512, 64) : make_desc(sb + k * 32, 1, 64);
      umma_f16(tmem + (uint32_t)((i / 4) % naccum) * (uint32_t)N, ad, bd, idesc, i >= 4 * naccum ? 1u : 0u);
    }
    long long t1 = clock64();
    umma_commit(bar);
    while (!mbar_try_wait(bar, 0)) {}
    long long t2 = clock64();
    out[blockIdx.x * 2 + 0] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 2 * 148 * sizeof(long long));
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 4096;
  for (int grid : {1, 148})
    for (int b_mn : {0, 1})
      for (int N : {64, 128, 256})
        for (int naccum : {1, 2})
          for (int kadv : {0, 1}) {
            if (N * naccum > 512) continue;
            mma_rate_kernel<<<grid, 128, 200 * 1024>>>(N, reps, naccum, kadv, b_mn, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double cyc = (double)h[1] / reps;
            printf("grid=%3d b_mn=%d N=%3d naccum=%d kadv=%d : issue %.1f cyc/mma, complete %.1f cyc/mma -> %.0f MAC/cyc/SM (%s)\n", grid, b_mn, N,
                   naccum, kadv, (double)h[0] / reps, cyc, 128.0 * N * 16 / cyc, cudaGetErrorString(e));
          }
  return 0;
}
