// DeviceContext / CUDAContext / SpanCacheConfig — the accessors the hot-path operators call
// (csrc/common/device_context.h:17-209, csrc/device/cuda/cuda_context.h:40-136, csrc/runtime/cache/span_cache_config.cpp).
#pragma once
#include <cuda_runtime.h>

#include <memory>

#include "../../include/b200spark.h"
#include "common.h"

namespace allspark {

struct SpanCacheConfig {
  using Ptr = std::shared_ptr<SpanCacheConfig>;
  AsCacheMode mode;
  int span_size, span_num_init, span_num_grow;
  static Ptr Create(AsCacheMode mode, int span_size, int span_num_init = 0, int span_num_grow = 0) {
    // span_cache_config.cpp:32-48: only 16/32/64/128 (0 disables span cache)
    if (span_size != 0 && span_size != 16 && span_size != 32 && span_size != 64 && span_size != 128) return nullptr;
    if (span_num_init < 0 || span_num_grow < 0) return nullptr;
    return std::make_shared<SpanCacheConfig>(SpanCacheConfig{mode, span_size, span_num_init, span_num_grow});
  }
};

class DeviceContext {
 public:
  virtual ~DeviceContext() = default;
  virtual DeviceType GetDeviceType() const = 0;
  virtual int GetRank() const = 0;
  virtual int GetNranks() const = 0;
  virtual void Synchronize() const = 0;
  void SetModelMaxLength(int v) { max_length_ = v; }
  int GetModelMaxLength() const { return max_length_; }
  void SetModelMaxBatch(int v) { max_batch_ = v; }
  int GetModelMaxBatch() const { return max_batch_; }
  void SetNumberHeads(int v) { num_heads_ = v; }
  int GetNumberHeads() const { return num_heads_; }
  void SetNumberGroups(int v) { num_groups_ = v; }
  int GetNumberGroups() const { return num_groups_; }
  void SetSizePerHead(int v) { size_per_head_ = v; }
  int GetSizePerHead() const { return size_per_head_; }
  void SetDecoderLayer(int v) { dec_layer_ = v; }
  int GetDecoderLayer() const { return dec_layer_; }
  void SetDtype(DataType t) { dtype_ = t; }
  DataType GetDtype() const { return dtype_; }
  void SetCacheConfig(SpanCacheConfig::Ptr c) { cache_config_ = c; }
  SpanCacheConfig::Ptr GetCacheConfig() const {
    if (!cache_config_) throw AsException("DeviceContext: cache config uninitialized");
    return cache_config_;
  }
  AsCacheMode GetCacheMode() const { return GetCacheConfig()->mode; }
  int GetCacheSpanSize() const { return GetCacheConfig()->span_size; }

 private:
  int max_length_ = 0, max_batch_ = 0, num_heads_ = 0, num_groups_ = 0, size_per_head_ = 0, dec_layer_ = 0;
  DataType dtype_ = DATATYPE_UNDEFINED;
  SpanCacheConfig::Ptr cache_config_;
};

// One stream per rank, like the reference.  Where the reference's CUDAContext carries an ncclComm_t (GetNCCLComm), this one
// carries the rank's b2_comm_t: the NVLink peer-memory communicator of include/b200spark.h (the engine creates it once per
// rank next to where it calls ncclCommInitRank today and exchanges the 64-byte IPC handles over its existing bootstrap).
class CUDAContext : public DeviceContext {
 public:
  CUDAContext() { cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking); }
  ~CUDAContext() override { if (stream_) cudaStreamDestroy(stream_); }
  DeviceType GetDeviceType() const override { return DeviceType::CUDA; }
  int GetRank() const override { return rank_; }
  int GetNranks() const override { return nranks_; }
  void SetRank(int rank, int nranks) { rank_ = rank; nranks_ = nranks; }
  void SetDeviceId(int id) { device_id_ = id; cudaSetDevice(id); }
  int GetDeviceId() const { return device_id_; }
  cudaStream_t GetStream() const { return stream_; }
  void Synchronize() const override { cudaStreamSynchronize(stream_); }
  void SetB2Comm(b2_comm_t c) { comm_ = c; }     // not owned
  b2_comm_t GetB2Comm() const { return comm_; }

 private:
  cudaStream_t stream_ = nullptr;
  int rank_ = 0, nranks_ = 1, device_id_ = 0;
  b2_comm_t comm_ = nullptr;
};

}  // namespace allspark
