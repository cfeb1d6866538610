// Span cache plumbing the decode attention operator reads:
//   VirtualCache::GetCache(layer, increment) -> AsTensor of span pointers   csrc/runtime/cache/virtual_cache.h:98-141
//   CacheUtils::GetSpanSizeInBytes                                          csrc/runtime/cache/virtual_cache.cpp:202-232
//   GenerateContext / RuntimeContext                                        csrc/common/generate_context.h:32-141
// The reference's frame/span managers (csrc/runtime/cache/, 4.7k lines) stay out of scope; the pool
// below is the smallest stand-in that hands out real device spans so the operator can be exercised like in a model.
#pragma once
#include <memory>
#include <vector>

#include "../../include/b200spark.h"
#include "device_context.h"
#include "tensor.h"

namespace allspark {

struct CacheUtils {
  static size_t GetSpanSizeInBytes(const SpanCacheConfig& cfg, DataType dtype, int num_heads, int per_head_size) {
    b2_span_cfg c{};
    c.ft = dtype; c.quant_mode = (int)cfg.mode; c.n_heads = num_heads; c.n_groups = num_heads;
    c.head_size = per_head_size; c.span_len = cfg.span_size; c.max_spans_per_seq = 1;
    return b2_span_bytes(&c);
  }
};

// A slab of equally sized device spans (stand-in for CacheFrameManager + CacheSpanManager).
class CacheSpanPool {
 public:
  CacheSpanPool(size_t span_bytes, int num_spans) : span_bytes_((span_bytes + 255) / 256 * 256), free_() {
    if (cudaMalloc(&base_, span_bytes_ * (size_t)num_spans) != cudaSuccess) throw AsException("CacheSpanPool: out of device memory");
    cudaMemset(base_, 0, span_bytes_ * (size_t)num_spans);
    for (int i = num_spans - 1; i >= 0; --i) free_.push_back(static_cast<char*>(base_) + (size_t)i * span_bytes_);
  }
  ~CacheSpanPool() { if (base_) cudaFree(base_); }
  void* Claim() {
    if (free_.empty()) return nullptr;
    void* p = free_.back();
    free_.pop_back();
    return p;
  }
  void Release(void* p) { free_.push_back(p); }
  size_t FreeSpans() const { return free_.size(); }

 private:
  size_t span_bytes_;
  void* base_ = nullptr;
  std::vector<void*> free_;
};

class VirtualCache {
 public:
  virtual ~VirtualCache() = default;
  virtual const AsTensor& GetCache(int layer_id, int increment) = 0;
  virtual size_t GetSeqLength(int layer_id) const = 0;
  virtual int GetLayerNum() const = 0;
};

// One request's K (or V) cache over all layers: per layer a growing vector of span pointers kept in a host
// POINTER tensor, claimed on demand when the sequence crosses a span boundary.
class SpannedVirtualCache : public VirtualCache {
 public:
  SpannedVirtualCache(std::shared_ptr<CacheSpanPool> pool, int layers, int span_len, int max_spans)
      : pool_(std::move(pool)), span_len_(span_len), max_spans_(max_spans), len_(layers, 0) {
    for (int l = 0; l < layers; ++l)
      ptrs_.push_back(std::make_unique<AsTensor>("span_ptrs", DeviceType::CPU, DataType::POINTER, DataMode::DENSE, Shape{0}));
  }
  ~SpannedVirtualCache() override {
    for (auto& t : ptrs_) {
      void** p = static_cast<void**>(t->GetDataPtr());
      for (int64_t i = 0; i < t->GetShape().Count(); ++i) pool_->Release(p[i]);
    }
  }
  const AsTensor& GetCache(int layer_id, int increment) override {
    if (layer_id < 0 || layer_id >= (int)ptrs_.size() || increment < 0) throw AsException("VirtualCache: bad layer/increment");
    const size_t new_len = len_[layer_id] + increment;
    const int64_t need = (int64_t)((new_len + span_len_ - 1) / span_len_);
    AsTensor& t = *ptrs_[layer_id];
    const int64_t have = t.GetShape().Count();
    if (need > max_spans_) throw AsException("VirtualCache: sequence longer than the engine max length");
    if (need > have) {
      std::vector<void*> v((void**)t.GetDataPtr(), (void**)t.GetDataPtr() + have);
      for (int64_t i = have; i < need; ++i) {
        void* s = pool_->Claim();
        if (!s) throw AsException("ALLSPARK_CACHE_MEMORY_OUT");
        v.push_back(s);
      }
      t.SetShape(Shape{need});
      t.CopyDataFrom(v.data(), v.size() * sizeof(void*), DeviceType::CPU);
    }
    len_[layer_id] = new_len;
    return t;
  }
  size_t GetSeqLength(int layer_id) const override { return len_[layer_id]; }
  int GetLayerNum() const override { return (int)ptrs_.size(); }

 private:
  std::shared_ptr<CacheSpanPool> pool_;
  int span_len_, max_spans_;
  std::vector<size_t> len_;
  std::vector<std::unique_ptr<AsTensor>> ptrs_;
};

struct GenerateContext {
  int step = 0;  // tokens already cached
  int prefix_len = 0;
  int num_beams = 1;
  int current_batch = 0;
  std::unique_ptr<VirtualCache> virtual_k_cache;
  std::unique_ptr<VirtualCache> virtual_v_cache;
};

class RuntimeContext {
 public:
  explicit RuntimeContext(bool is_prefill) : is_context(is_prefill) {}
  const bool is_context;
  int current_batch = 0;
  GenerateContext* GetContextGenCtx() const { return list_[current_batch].get(); }
  GenerateContext* GetGenCtx(int i) const { return list_[i].get(); }
  int GetGenCtxListSize() const { return (int)list_.size(); }
  void PushBackGenCtx(std::unique_ptr<GenerateContext> g) {
    list_.push_back(std::move(g));
    list_.back()->current_batch = (int)list_.size() - 1;
  }

 private:
  std::vector<std::unique_ptr<GenerateContext>> list_;
};

}  // namespace allspark
