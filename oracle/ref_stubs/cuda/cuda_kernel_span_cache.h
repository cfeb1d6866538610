// oracle/_ref build stub (TEST INFRASTRUCTURE): declarations of the two span-cache launchers that oracle/ref_shim.cu
// instantiates from the reference's own .cuh files (the original header also declares prefix-cache copies that need the
// engine's DataType enum).  Signatures follow csrc/core/kernel/cuda/cuda_kernel_span_cache.h:11-27.
#pragma once
#include <cuda_runtime.h>
#include <span_attn.h>
#include <stdint.h>

namespace allspark {
namespace cuda {
template <typename T>
void ContextSpanCopyLauncher(void* const* spanPtrs, const T* src, int nGroups, int headSize, int spanLen, int seqLen,
                             span::QuantMode cacheMode, cudaStream_t stream);
template <typename T>
void DecoderCacheAppendLauncher(void* const* kSpanArray, void* const* vSpanArray, T* queryOut, const T* src,
                                const uint32_t* oldSeqLens, int batchSize, int nHeads, int nGroups, int headSize,
                                int spanLen, int nSpansPerBatch, span::QuantMode cacheMode, cudaStream_t stream);
}  // namespace cuda
}  // namespace allspark
