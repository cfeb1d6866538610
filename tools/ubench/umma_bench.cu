// Micro-benchmark: tcgen05.mma issue/execute rate for M=128, K=16 bf16, N in {64,128,256}, A from TMEM or SMEM,
// 1 or 2 accumulators, 1 or 2 CTAs per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N, bool A_TMEM, int NACC>
__global__ void __launch_bounds__(128) k(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 65536 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  asm volatile("fence.proxy.async.shared::cta;");
  const uint32_t tmem = tslot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
    const uint32_t b_addr = smem_u32(smem);            // B tile: N rows x 128 B
    const uint32_t a_addr = smem_u32(smem + 32768);    // A tile (SS mode): 128 rows x 128 B
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t bdesc = desc_hi | (uint64_t)((((b_addr + kk * 32) >> 4) & 0x3FFF) | (1u << 16));
        const uint32_t d = tmem + ((it * 4 + kk) % NACC) * N;
        if (A_TMEM) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                       ::"r"(d), "r"(tmem + 448 + kk * 8), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
        } else {
          const uint64_t adesc = desc_hi | (uint64_t)((((a_addr + kk * 32) >> 4) & 0x3FFF) | (1u << 16));
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                       ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
        }
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, bool A_TMEM, int NACC>
void run(const char* name, int ctas) {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 256;
  auto kern = k<N, A_TMEM, NACC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  kern<<<ctas, 128, 66 * 1024>>>(d, iters);
  kern<<<ctas, 128, 66 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-34s ctas=%3d: issue %6.1f cyc/MMA, issue+drain %6.1f cyc/MMA (%s)\n", name, ctas, (double)h[0] / (iters * 4),
         (double)h[1] / (iters * 4), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int ctas : {1, 148}) {
    run<64, true, 1>("N=64  A=TMEM 1 acc", ctas);
    run<64, true, 2>("N=64  A=TMEM 2 acc", ctas);
    run<64, true, 4>("N=64  A=TMEM 4 acc", ctas);
    run<64, false, 1>("N=64  A=SMEM 1 acc", ctas);
    run<128, true, 1>("N=128 A=TMEM 1 acc", ctas);
    run<128, true, 2>("N=128 A=TMEM 2 acc", ctas);
    run<256, true, 1>("N=256 A=TMEM 1 acc", ctas);
    run<256, false, 1>("N=256 A=SMEM 1 acc", ctas);
  }
  // two CTAs per SM (smem 66 KB each, 512 TMEM columns each would not fit: use 296 CTAs only for N<=64 with smaller alloc) skipped
  return 0;
}
