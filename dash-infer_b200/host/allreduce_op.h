// AllReduceOp (op type "AllReduce") — forwarder to b2_allreduce.
// Reference: csrc/core/operator/nccl/allreduce/allreduce_op.cpp:25-115 (ncclAllReduce on the op's stream followed by a
// ctx_->Synchronize() of the whole device every call).  Here: one-shot exchange over NVLink peer memory, stream-ordered, no
// host synchronisation, CUDA-graph replayable; sums are fp32 in rank order (deterministic), rounded to the model dtype once.
#pragma once
#include "operator.h"

namespace allspark {

class AllReduceOp : public AsOperator {
 public:
  using AsOperator::AsOperator;
  AsStatus Init(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map, TensorMap* tensor_map) override;
  AsStatus Reshape(RuntimeContext* runtime_ctx) override;
  AsStatus Reshape() override { return Reshape(nullptr); }
  AsStatus Forward(RuntimeContext* runtime_ctx) override;
  AsStatus Forward() override { return Forward(nullptr); }

 private:
  int64_t count_ = 0;
  int nranks_ = 1;
};

}  // namespace allspark
