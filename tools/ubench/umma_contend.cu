// Micro-benchmark: what slows tcgen05.mma (M=128, N=64, K=16, A in TMEM, B in smem) below its 44.8-cycle floor inside the
// weight-only GEMM?  One thread issues MMAs; the other warps generate one kind of contention each (flag bitmask):
//   1  tcgen05.commit after every 8 MMAs          2  warps 1-3 spin on an mbarrier (try_wait polling)
//   4  warps 4-7 stream tcgen05.st (32x32b.x8) + wait::st      8  warps 2-3 stream LDS.128 over a 16 KB tile
//   16 B tile address rotates over 4 tiles x 4 k-steps (as in the kernel) instead of one tile
//   32 one accumulator chain broken into 2 alternating accumulators
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_contend umma_contend.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512) k(long long* out, int iters, int flags, int a_off, int d_off, int a_step, int layout = 2, int sbo = 1024, int lbo = 0, int kstep = 32, int base_off = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tslot;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    done = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 65536 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  asm volatile("fence.proxy.async.shared::cta;");
  const uint32_t tmem = tslot;
  constexpr int N = 64;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t desc_hi = (uint64_t)(((uint32_t)sbo >> 4) | (1u << 14) | ((uint32_t)layout << 29)) << 32;
    const uint32_t b_addr = smem_u32(smem) + base_off;
    if (blockIdx.x == 0) out[3] = b_addr;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t tile = (flags & 16) ? ((it * 2 + (kk >> 2)) & 3) * 8192u : 0u;
        const uint64_t bdesc = desc_hi | (uint64_t)((((b_addr + tile + (kk & 3) * kstep) >> 4) & 0x3FFF) | ((uint64_t)(((uint32_t)lbo >> 4) & 0x3FFF) << 16));
        const uint32_t d = tmem + ((flags & 32) ? (kk & 1) * 64 : 0);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(d + d_off), "r"(tmem + a_off + kk * a_step), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u) : "memory");
      }
      if (flags & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar3)) : "memory");
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    done = 1;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp >= 1 && warp <= 3 && (flags & 2)) {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(0u) : "memory");
    }
  } else if (warp >= 2 && warp <= 3 && (flags & 8)) {
    float acc = 0.f;
    const uint32_t base = smem_u32(smem + 32768) + (tid - 64) * 128;
    while (!done) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + ((c ^ (lane & 7)) << 4)));
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
      }
    }
    if (acc == 123.f) out[2] = 1;
  } else if (warp >= 4 && (flags & 4)) {
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    uint32_t a = tid;
    while (!done) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(trow + j * 8), "r"(a) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      a += 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  const int iters = 128;
  for (int threads : {32, 64, 128, 160, 192, 256, 288, 416, 512}) {
    k<<<148, threads, 66 * 1024>>>(d, iters, 0, 448, 0, 8);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("threads=%3d: issue %6.1f cyc/MMA (%s)\n", threads, (double)h[0] / (iters * 8), cudaGetErrorString(e));
  }
  struct L { const char* name; int layout, sbo, lbo, kstep, base; };
  const L ls[] = {{"SW128 aligned", 2, 1024, 16, 32, 0}, {"SW128 lbo=0", 2, 1024, 0, 32, 0}, {"SW128 base+16", 2, 1024, 16, 32, 16}, {"SW128 base+128", 2, 1024, 16, 32, 128},
                  {"SW128 base+512", 2, 1024, 16, 32, 512}, {"SW64", 4, 512, 16, 32, 0}, {"SW32", 6, 256, 16, 2048, 0}, {"none lbo128 sbo256", 0, 256, 128, 2048, 0},
                  {"none lbo=1024 sbo=128 (k-core planes)", 0, 128, 1024, 2048, 0}, {"SW128_base32B", 1, 1024, 16, 32, 0}};
  for (const L& l : ls) {
    k<<<148, 256, 66 * 1024>>>(d, iters, 0, 448, 0, 8, l.layout, l.sbo, l.lbo, l.kstep, l.base);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4];
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("%-40s: issue %6.1f cyc/MMA  b_addr=0x%llx (%s)\n", l.name, (double)h[0] / (iters * 8), h[3], cudaGetErrorString(e));
  }
  for (int a_off : {448})
    for (int d_off : {0, 64, 448}) {
      if (d_off + 64 > a_off && d_off < a_off + 64) continue;
      for (int a_step : {8}) {
        k<<<148, 256, 66 * 1024>>>(d, iters, 0, a_off, d_off, a_step);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("a_off=%3d d_off=%3d a_step=%d: issue %6.1f cyc/MMA (%s)\n", a_off, d_off, a_step, (double)h[0] / (iters * 8), cudaGetErrorString(e));
      }
    }
  for (int a_step : {0, 4, 16}) {
    k<<<148, 256, 66 * 1024>>>(d, iters, 0, 256, 0, a_step);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("a_off=256 d_off=0 a_step=%d: issue %6.1f cyc/MMA\n", a_step, (double)h[0] / (iters * 8));
  }
  for (int ctas : {148}) {
    for (int flags : {0, 1, 2, 4, 8, 16, 32, 1 | 16, 1 | 2 | 4 | 8 | 16, 1 | 2 | 4 | 8 | 16 | 32, 4 | 32, 4 | 8}) {
      k<<<ctas, 256, 66 * 1024>>>(d, iters, flags, 448, 0, 8);
      k<<<ctas, 256, 66 * 1024>>>(d, iters, flags, 448, 0, 8);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("ctas=%3d flags=%2d: issue %6.1f cyc/MMA, issue+drain %6.1f cyc/MMA (%s)\n", ctas, flags, (double)h[0] / (iters * 8),
             (double)h[1] / (iters * 8), cudaGetErrorString(e));
    }
  }
  return 0;
}
