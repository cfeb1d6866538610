// SpanAttnOpCUDA (op types DecOptMHA / DecOptMQA), decode branch — forwarder to b2_span_cache_append + b2_span_attn_run.
// Reference: SpanAttnOp::{Init,Reshape,Alloc,Forward} csrc/core/operator/generate_opt/span_attn/span_attn_op.cpp:172-368,
//            SpanAttnOpCUDA::{decoderAppendCacheLauncher,decoderAttnLauncher} span_attn_op_cuda.cpp:287-392,542-587.
// Differences by design: the attention handle is created once (not per layer per step), the span pointer tables live
// on the device and are re-uploaded only when a sequence claims a new span (through a pinned staging ring: Forward never
// synchronises the stream), and the prefill branch is out of scope.
#pragma once
#include "operator.h"

namespace allspark {

class SpanAttnOpCUDA : public AsOperator {
 public:
  using AsOperator::AsOperator;
  ~SpanAttnOpCUDA() override;
  AsStatus InitV2(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map,
                  TensorMap& weights_buffer, TensorMap* tensor_map, RuntimeContext* runtime_ctx) override;
  AsStatus Reshape(RuntimeContext* runtime_ctx) override;
  AsStatus Alloc(RuntimeContext* runtime_ctx) override;
  AsStatus Forward(RuntimeContext* runtime_ctx) override;

 private:
  b2_span_cfg cfg_{};
  b2_span_attn_t handle_ = nullptr;
  int layer_num_ = 0, batch_size_ = 0, max_batch_ = 0, max_spans_ = 0;
  float alpha_ = -1.0f;
  std::unique_ptr<AsTensor> q_tensor_, k_tab_, v_tab_, old_lens_, new_lens_;
  // pinned host staging, kStages deep (span_attn_op_cuda.cpp:287-392 stages through pinned "host workspace" too)
  static constexpr int kStages = 4;
  struct Stage { void* host = nullptr; cudaEvent_t done = nullptr; bool busy = false; };
  Stage stage_[kStages];
  int stage_i_ = 0;
  size_t slot_bytes_ = 0;
  std::vector<int> span_counts_;  // spans uploaded per sequence (re-upload only on change)
};

}  // namespace allspark
