// GemmA16W4GPU / GemmA16W8GPU / GemmOpGPU — thin forwarders from the allspark operator contract to b2_gemm_wq_*.
// Reference classes replaced:
//   GemmA16W4Base/GPU  csrc/core/operator/general/gemm_lowp/gemm_a16w4.cpp:21-115, gemm_a16w4_gpu.cpp:92-255
//   GemmA16W8Base/GPU  csrc/core/operator/general/gemm_lowp/gemm_a16w8.cpp:21-107, gemm_a16w8_gpu.cpp:30-299
//   GemmOpGPU (dense)  csrc/core/operator/general/gemm/gemm_op_gpu.cpp (cuBLAS) — lm_head / unquantized projections
// Weights arrive in the reference order [qdata, scales, zeros, (bias)] and layouts ([K,N/2] uint4x2 / [K,N] int8);
// InitV2 re-lays them out once (the reference pads / reorders at init too).
#pragma once
#include "operator.h"

namespace allspark {

class GemmLowpGPUBase : public AsOperator {
 public:
  using AsOperator::AsOperator;
  ~GemmLowpGPUBase() override;
  AsStatus InitV2(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map,
                  TensorMap& weights_buffer, TensorMap* tensor_map, RuntimeContext* runtime_ctx) override;
  AsStatus Reshape() override;
  AsStatus Forward() override;
  // re-run the init-time re-layout from the (restored) weight tensors: the counterpart of the reference's swap-in path
  // that copies weights_buffer back to the device (gemm_a16w4_gpu.cpp:296-300)
  AsStatus ReloadWeights();

 protected:
  virtual int wbits() const = 0;
  b2_gemm_wq_t handle_ = nullptr;
  int64_t m_ = 0, n_ = 0, k_ = 0, lda_ = 0;
  int binary_type_ = BINARYTYPE_UNDEFINED;
  bool is_split_k_ = false;
  int group_size_ = -1;
  int activation_ = UNARYTYPE_UNDEFINED;
  float alpha_ = 1.0f;
  bool transB_ = false, is_pooler_ = false;
  DataType qtype_ = DATATYPE_UNDEFINED;
};

class GemmA16W4GPU : public GemmLowpGPUBase {
 public:
  using GemmLowpGPUBase::GemmLowpGPUBase;
 protected:
  int wbits() const override { return 4; }
};
class GemmA16W8GPU : public GemmLowpGPUBase {
 public:
  using GemmLowpGPUBase::GemmLowpGPUBase;
 protected:
  int wbits() const override { return 8; }
};
class GemmOpGPU : public GemmLowpGPUBase {
 public:
  using GemmLowpGPUBase::GemmLowpGPUBase;
 protected:
  int wbits() const override { return 16; }
};

}  // namespace allspark
