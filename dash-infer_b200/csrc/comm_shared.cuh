// b200spark — device view of a tensor-parallel communicator (comm.cu), shared with the GEMV kernel's fused
// all-reduce epilogue (wq_gemm.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace b2 {

constexpr int kCommMaxRanks = 8;        // one NVSwitch domain
constexpr int kCommMaxChunks = 128;     // flags per (parity, source rank): all-reduce chunks or 128-channel GEMV tiles
constexpr int kCommFlagStride = 32;     // bytes between flags (a flag is one 32-bit epoch; its own sector)

// Exchange buffer of one rank:  [control 256 B | data: 2 parities x nranks slots x slot_bytes | flags: 2 x nranks x chunks]
struct CommDev {
  uint8_t* peer[kCommMaxRanks];  // every rank's exchange buffer as mapped into THIS process (peer[rank] = local)
  unsigned* epoch;               // local: exchanges completed on this communicator (advanced by the kernels)
  unsigned* done;                // local: CTAs of the running exchange that finished
  int* error;                    // local: set when a peer's flag did not arrive within timeout_ns
  size_t slot_bytes, data_off, flag_off, max_bytes;
  unsigned long long timeout_ns;
  int rank, nranks;
};

__host__ __device__ __forceinline__ size_t comm_slot_offset(const CommDev& c, int parity, int src_rank) {
  return c.data_off + ((size_t)parity * c.nranks + src_rank) * c.slot_bytes;
}
__host__ __device__ __forceinline__ size_t comm_flag_offset(const CommDev& c, int parity, int src_rank, int chunk) {
  return c.flag_off + (((size_t)parity * c.nranks + src_rank) * kCommMaxChunks + chunk) * kCommFlagStride;
}

#ifdef __CUDACC__
// spin on a flag in local memory until it reaches `want`; a peer that never shows up trips the timeout (device-side: the
// box must not hang) and the launch reports through comm->error
__device__ __forceinline__ bool wait_flag(const unsigned* flag, unsigned want, unsigned long long timeout_ns, int* error) {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  while (true) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int)(v - want) >= 0) return true;
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) {
        atomicExch(error, 1);
        return false;
      }
    }
  }
}

#endif

}  // namespace b2
