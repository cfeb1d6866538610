// b200spark — weight-only quantized GEMM for decode batches 17..64 on the 5th-gen tensor cores (tcgen05), sm_100a.
//
// Replaces the reference's "dequantize the whole weight to a [K,N] bf16 workspace, then cuBLAS" fallback
// (csrc/core/operator/general/gemm_lowp/gemm_a16w4_gpu.cpp:193-210, gemm_a16w8_gpu.cpp:210-237): 4.5 B of HBM
// traffic per weight there, 0.5 B (int4) / 1 B (int8) here.
//
//   C^T[128 n x NM m] (fp32, TMEM) += W^T[128 n x 16 k] (bf16, TMEM) * A^T[16 k x NM m] (bf16, shared memory)
//
//   warp 0      : TMA producer — int4/int8 weight tiles (same init-time image as the mma.sync kernel) stream
//                 HBM -> shared memory with cp.async.bulk, ahead of the previous kernel's completion (PDL)
//   warps 2..5  : dequant — one thread per output channel: LDS.128 -> lop3/shf -> exact bf16 integers (16+q)
//                 -> tcgen05.st into the A-operand region of TMEM (the dequantized weights never touch shared
//                 memory: its bandwidth could not carry 2 B/weight at HBM rate)
//   warps 6..7  : activation tiles (64 k x NM m) via cp.async into the 128B-swizzled K-major UMMA layout, plus
//                 the per-row sums sum_k a[m][k] needed by the zero-point term
//   warp 1      : one elected thread issues tcgen05.mma (A from TMEM, B from shared memory, D in TMEM) and
//                 tcgen05.commit's the pipeline barriers
//   epilogue    : warps 2..5 read D with tcgen05.ld, apply s * (acc - (16+z) * sum a), split-K partial or final
//                 alpha/bias/activation/residual, bf16 store.
//
// Roofline: HBM-bound up to M ~ 64 (256 FLOP/B ~ the tensor/HBM ridge); report both.
#include <cstdlib>

#include "b2_common.cuh"
#include "wq_gemm_shared.cuh"

namespace b2 {

constexpr int kTcThreads = 256;
constexpr int kTcNM = 64;            // batch columns per MMA (UMMA N)
constexpr int kTcNSW = 8;            // weight stages
constexpr int kTcNSX = 4;            // activation stages
constexpr int kTcXTile = kTcNM * 128;  // bytes: NM rows x 64 k bf16
constexpr int kTcColsD = 0;          // TMEM columns [0, 64): accumulator
constexpr int kTcColsA = 64;         // TMEM columns [64, ...): two A buffers (W4: 32 columns each, W8: 64 = lo+hi planes)
constexpr int kTcTmemCols = 256;

// ---- tcgen05 wrappers (forms as in cute/arch/{mma_sm100_umma,copy_sm100,tmem_allocator_sm100}.hpp) ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

struct TcParams {
  const uint8_t* packed;
  const float2* sz;
  const __nv_bfloat16* A;
  int64_t lda;
  __nv_bfloat16* C;
  int64_t ldc;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;
  float* ws;
  unsigned* counters;
  int M, N, K, Np, KT, NG, S;
  int act;
  float alpha;
};

template <int WBITS>
__global__ void __launch_bounds__(kTcThreads, 1) wq_gemm_tc_kernel(const TcParams p) {
  constexpr int TILE_BYTES = WBITS == 4 ? 4096 : 8192;
  constexpr int NCH = WBITS == 4 ? 2 : 4;  // 16B chunks per row per k-tile
  constexpr int ABUF = WBITS == 4 ? 32 : 64;  // TMEM columns per A buffer
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xring = smem;                                   // NSX x 8 KB, 1024B aligned (SWIZZLE_128B atoms)
  uint8_t* wring = xring + kTcNSX * kTcXTile;              // NSW x TILE_BYTES
  float* suma = reinterpret_cast<float*>(wring + kTcNSW * TILE_BYTES);  // [NM]
  uint64_t* bars = reinterpret_cast<uint64_t*>(suma + kTcNM);
  uint64_t* wfull = bars;
  uint64_t* wfree = wfull + kTcNSW;
  uint64_t* xfull = wfree + kTcNSW;
  uint64_t* xfree = xfull + kTcNSX;
  uint64_t* afull = xfree + kTcNSX;
  uint64_t* afree = afull + 2;
  uint64_t* dfull = afree + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + 1);
  __shared__ int s_is_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ng = blockIdx.x / p.S;
  const int s = blockIdx.x - ng * p.S;
  const int kt0 = (int)((int64_t)s * p.KT / p.S), kt1 = (int)((int64_t)(s + 1) * p.KT / p.S);
  const int nt = kt1 - kt0;

  if (tid == 0) {
    for (int i = 0; i < kTcNSW; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wfree[i], 4); }
    for (int i = 0; i < kTcNSX; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xfree[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&afull[i], 4); mbar_init(&afree[i], 1); }
    mbar_init(dfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {  // TMEM allocation (this warp also frees it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTcTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== weight producer (does not wait for the previous kernel) =====================
    if (lane == 0) {
      const uint8_t* wsrc = p.packed + ((size_t)ng * p.KT + kt0) * TILE_BYTES;
      for (int j = 0; j < nt; ++j) {
        const int slot = j % kTcNSW;
        if (j >= kTcNSW) mbar_wait(&wfree[slot], ((j / kTcNSW) & 1) ^ 1);
        mbar_arrive_expect_tx(&wfull[slot], TILE_BYTES);
        bulk_g2s(wring + slot * TILE_BYTES, wsrc + (size_t)j * TILE_BYTES, TILE_BYTES, &wfull[slot]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N = NM, M = 128 (cute::UMMA::InstrDescriptor)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcNM >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // B smem descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_128B, SBO = 1024 B (8-row groups), version 1
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    for (int j = 0; j < nt; ++j) {
      const int ab = j & 1, xs = j % kTcNSX;
      mbar_wait(&afull[ab], (j >> 1) & 1);
      mbar_wait(&xfull[xs], (j / kTcNSX) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t xaddr = smem_u32(xring + xs * kTcXTile);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)((((xaddr + kk * 32) >> 4) & 0x3FFF) | (1u << 16));
          if (WBITS == 4) {
            tc_mma_ts(tmem + kTcColsD, tmem + kTcColsA + ab * ABUF + kk * 8, bdesc, idesc, (j > 0 || kk > 0) ? 1u : 0u);
          } else {  // W8: per k16 step the buffer holds [lo plane | hi plane], 8 columns each
            tc_mma_ts(tmem + kTcColsD, tmem + kTcColsA + ab * ABUF + kk * 16, bdesc, idesc, (j > 0 || kk > 0) ? 1u : 0u);
            tc_mma_ts(tmem + kTcColsD, tmem + kTcColsA + ab * ABUF + kk * 16 + 8, bdesc, idesc, 1u);
          }
        }
        tc_commit(&afree[ab]);
        tc_commit(&xfree[xs]);
        if (j == nt - 1) tc_commit(dfull);
      }
      __syncwarp();
    }
  } else if (warp >= 6) {
    // ===================== activation tiles + row sums =====================
    const int xt = tid - 192;  // 0..63 == row m owned for the sums
    pdl_wait();
    float rsum = 0.f;
    auto issue = [&](int jj) {
      const int slot = jj % kTcNSX;
      if (jj >= kTcNSX) mbar_wait(&xfree[slot], ((jj / kTcNSX) & 1) ^ 1);
      uint8_t* dst = xring + slot * kTcXTile;
      const int64_t k0 = (int64_t)(kt0 + jj) * kBK;
#pragma unroll
      for (int i = 0; i < kTcNM * 8 / 64; ++i) {
        const int idx = xt + i * 64;
        const int row = idx >> 3, c = idx & 7;
        const bool valid = row < p.M && (k0 + c * 8) < p.K;
        const __nv_bfloat16* src = valid ? p.A + (int64_t)row * p.lda + k0 + c * 8 : p.A;
        cp_async16_zfill(dst + row * 128 + ((c ^ (row & 7)) << 4), src, valid);
      }
    };
    for (int jj = 0; jj < kTcNSX - 1; ++jj) {
      if (jj < nt) issue(jj);
      cp_async_commit();
    }
    for (int j = 0; j < nt; ++j) {
      if (j + kTcNSX - 1 < nt) issue(j + kTcNSX - 1);
      cp_async_commit();
      cp_async_wait<kTcNSX - 1>();
      fence_proxy_async();               // generic-proxy writes -> visible to the tensor core (async proxy)
      asm volatile("bar.sync 2, 64;" ::: "memory");
      const int slot = j % kTcNSX;
      if (xt == 0) mbar_arrive(&xfull[slot]);
      // row sum of this tile (row = xt): 8 swizzled 16B chunks
      const uint32_t rbase = smem_u32(xring + slot * kTcXTile) + xt * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = lds128(rbase + ((c ^ (xt & 7)) << 4));
        rsum += (bf16_lo(v.x) + bf16_hi(v.x)) + (bf16_lo(v.y) + bf16_hi(v.y)) + (bf16_lo(v.z) + bf16_hi(v.z)) +
                (bf16_lo(v.w) + bf16_hi(v.w));
      }
      asm volatile("bar.sync 2, 64;" ::: "memory");  // all sums of this slot done before it can be refilled (issue waits xfree too)
    }
    suma[xt] = rsum;
    asm volatile("bar.sync 3, 192;" ::: "memory");  // hand the sums to the epilogue warps
  } else {
    // ===================== dequant (warps 2..5) then epilogue =====================
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;        // output channel (row of the 128-row tile)
    const int n = ng * kBN + r;
    const float2 sz = p.sz[n];          // per-channel (scale, zero + bias constant): immutable, read before the wait
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t wring_u = smem_u32(wring);
    uint32_t woff[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) woff[c] = c * 2048 + ((r ^ tile_swz(WBITS, c)) << 4);

    for (int j = 0; j < nt; ++j) {
      const int slot = j % kTcNSW, ab = j & 1;
      mbar_wait(&wfull[slot], (j / kTcNSW) & 1);
      if (j >= 2) mbar_wait(&afree[ab], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t wt = wring_u + slot * TILE_BYTES;
      if (WBITS == 4) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint4 wv = lds128(wt + woff[c]);
          const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // k16 step kk = 2c + h: words 2h, 2h+1
            uint32_t a[8];
#pragma unroll
            for (int jw = 0; jw < 2; ++jw) {
              const uint32_t w = ww[2 * h + jw];
              a[4 * jw + 0] = lop3_and_or(w, kMask4, kMagic);
              a[4 * jw + 1] = lop3_and_or(__funnelshift_r(w, w, 4), kMask4, kMagic);
              a[4 * jw + 2] = lop3_and_or(__funnelshift_r(w, w, 8), kMask4, kMagic);
              a[4 * jw + 3] = lop3_and_or(__funnelshift_r(w, w, 12), kMask4, kMagic);
            }
            tc_st8(trow + kTcColsA + ab * ABUF + (2 * c + h) * 8, a);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // chunk c = k16 step kk
          const uint4 wv = lds128(wt + woff[c]);
          const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
          uint32_t lo[8], hi[8];
#pragma unroll
          for (int jw = 0; jw < 4; ++jw) {
            const uint32_t w = ww[jw];
            lo[2 * jw + 0] = lop3_and_or(w, kMask4, kMagic);
            hi[2 * jw + 0] = lop3_and_or(__funnelshift_r(w, w, 4), kMask4, kMagicHi);
            lo[2 * jw + 1] = lop3_and_or(__funnelshift_r(w, w, 8), kMask4, kMagic);
            hi[2 * jw + 1] = lop3_and_or(__funnelshift_r(w, w, 12), kMask4, kMagicHi);
          }
          tc_st8(trow + kTcColsA + ab * ABUF + c * 16, lo);
          tc_st8(trow + kTcColsA + ab * ABUF + c * 16 + 8, hi);
        }
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&wfree[slot]);
        mbar_arrive(&afull[ab]);
      }
    }

    // ---------------- epilogue ----------------
    pdl_wait();  // workspace / counters / C belong to the previous kernels until here
    mbar_wait(dfull, 0);
    tc_fence_after();
    asm volatile("bar.sync 3, 192;" ::: "memory");  // row sums ready
    uint32_t d0[32], d1[32];
    tc_ld32(trow + kTcColsD, d0);
    tc_ld32(trow + kTcColsD + 32, d1);
    tc_wait_ld();
    const int et = tid - 64;  // 0..127
    const int MPK = kTcNM * kBN;
    float v[kTcNM];
#pragma unroll
    for (int m = 0; m < 32; ++m) {
      v[m] = sz.x * (__uint_as_float(d0[m]) - sz.y * suma[m]);
      v[m + 32] = sz.x * (__uint_as_float(d1[m]) - sz.y * suma[m + 32]);
    }
    if (p.S > 1) {
      float* wsu = p.ws + ((size_t)ng * p.S + s) * MPK;
#pragma unroll
      for (int m = 0; m < kTcNM; ++m)
        if (m < p.M) wsu[m * kBN + r] = v[m];
      __threadfence();
      asm volatile("bar.sync 4, 128;" ::: "memory");
      if (et == 0) {
        const unsigned prev = atomicAdd(&p.counters[ng], 1u);
        s_is_last = (prev == (unsigned)(p.S - 1));
      }
      asm volatile("bar.sync 4, 128;" ::: "memory");
      if (s_is_last) {
        __threadfence();
        const float* wsg = p.ws + (size_t)ng * p.S * MPK;
#pragma unroll 4
        for (int m = 0; m < kTcNM; ++m) {
          if (m >= p.M) break;
          float a = 0.f;
          for (int s0 = 0; s0 < p.S; s0 += 8) {
            float b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) b[u] = (s0 + u < p.S) ? __ldcg(wsg + (size_t)(s0 + u) * MPK + m * kBN + r) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) a += b[u];
          }
          v[m] = a;
        }
        if (et == 0) p.counters[ng] = 0;
      }
    }
    if (p.S == 1 || s_is_last) {
      if (n < p.N) {
        const float bv = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
#pragma unroll
        for (int m = 0; m < kTcNM; ++m) {
          if (m < p.M) {
            float o = apply_act_rt(v[m] * p.alpha + bv, p.act);
            if (p.residual) o += __bfloat162float(p.residual[(int64_t)m * p.ldc + n]);
            p.C[(int64_t)m * p.ldc + n] = __float2bfloat16(o);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols) : "memory");
  }
}

int tc_smem_bytes(int wbits) {
  const int tile = wbits == 4 ? 4096 : 8192;
  return 1024 + kTcNSX * kTcXTile + kTcNSW * tile + kTcNM * 4 + 32 * 8 + 64;
}

cudaError_t tc_configure(int wbits) {
  if (wbits == 4) return cudaFuncSetAttribute(wq_gemm_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(4));
  return cudaFuncSetAttribute(wq_gemm_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(8));
}

cudaError_t tc_launch(int wbits, const TcLaunch& a, cudaStream_t stream) {
  TcParams p;
  p.packed = a.packed; p.sz = a.sz; p.A = a.A; p.lda = a.lda; p.C = a.C; p.ldc = a.ldc; p.bias = a.bias; p.residual = a.residual;
  p.ws = a.ws; p.counters = a.counters; p.M = a.M; p.N = a.N; p.K = a.K; p.Np = a.Np; p.KT = a.KT; p.NG = a.NG; p.S = a.S;
  p.act = a.act; p.alpha = a.alpha;
  if (wbits == 4) return launch(wq_gemm_tc_kernel<4>, dim3(a.NG * a.S), dim3(kTcThreads), (size_t)tc_smem_bytes(4), stream, true, p);
  return launch(wq_gemm_tc_kernel<8>, dim3(a.NG * a.S), dim3(kTcThreads), (size_t)tc_smem_bytes(8), stream, true, p);
}

}  // namespace b2
