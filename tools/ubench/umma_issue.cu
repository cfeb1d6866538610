// Micro-benchmark: tcgen05.mma issue cost with per-MMA operand arithmetic (M=128, N=64, K=16, A in TMEM), as issued by the
// weight-only GEMM: (0) operands computed right before each MMA, (1) software-pipelined: MMA i+1's operands are computed
// (pinned with asm volatile) before MMA i issues.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue umma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 bd, {%2, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d), "r"(a), "r"(blo), "r"(idesc), "r"(acc), "r"(0u), "r"(bhi) : "memory");
}
__device__ __forceinline__ uint32_t padd(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(128) k(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tslot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 98304 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  asm volatile("fence.proxy.async.shared::cta;");
  const uint32_t tmem = __shfl_sync(0xffffffffu, tslot, 0);
  if (warp == 1) {
    const uint32_t idesc = 0x8100490;
    const uint32_t bhi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t xbase = smem_u32(smem);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int g = __shfl_sync(0xffffffffu, it, 0);
      const int ab = g % 3, xs = g % 3;
      if (lane == 0) {
        const uint32_t a0 = tmem + 64 + ab * 128;
        const uint32_t b0 = (((xbase + xs * 32768) >> 4) & 0x3FFF) | (1u << 16);
        if (MODE == 1) {
          uint32_t a_cur = a0, b_cur = b0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t a_nxt = padd(a_cur, 8);
            const uint32_t b_nxt = padd(b_cur, (i & 3) == 3 ? 512 - 6 : 2);
            mma(tmem, a_cur, b_cur, bhi, idesc, (it > 0 || i > 0) ? 1u : 0u);
            a_cur = a_nxt; b_cur = b_nxt;
          }
        } else {
#pragma unroll
          for (int ti = 0; ti < 4; ++ti)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              mma(tmem, a0 + ti * 32 + kk * 8, b0 + ti * 512 + 2 * kk, bhi, idesc, (it > 0 || ti > 0 || kk > 0) ? 1u : 0u);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int MODE> void run() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024);
  const int iters = 64;
  for (int r = 0; r < 2; ++r) k<MODE><<<148, 128, 98304 + 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("mode %d: issue %6.1f cyc/MMA, issue+drain %6.1f cyc/MMA (%s)\n", MODE, (double)h[0] / (iters * 16), (double)h[1] / (iters * 16), cudaGetErrorString(e));
}
int main() { run<0>(); run<1>(); return 0; }
