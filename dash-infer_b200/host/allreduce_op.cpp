#include "allreduce_op.h"

namespace allspark {

AsStatus AllReduceOp::Init(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map, TensorMap* tensor_map) {
  AS_CHECK_STATUS(AsOperator::Init(op_proto, ctx, weights_map, tensor_map));
  DataType dtype = tensor_map_->at(in_names_[0])->GetDataType();
  if (dtype == DATATYPE_UNDEFINED) dtype = ctx.GetDtype();
  if (dtype != DataType::BFLOAT16) {  // allreduce_op.cpp:35-52 maps FLOAT32 / FLOAT16 / BFLOAT16; this build exchanges bf16
    AS_LOG_ERROR("AllReduce: only BFLOAT16 activations are exchanged by this build");
    return AsStatus::ALLSPARK_PARAM_ERROR;
  }
  tensor_map_->at(out_names_[0])->SetDataType(dtype);
  nranks_ = ctx.GetNranks() > 0 ? ctx.GetNranks() : 1;
  if (nranks_ > 1 && !static_cast<const CUDAContext*>(ctx_)->GetB2Comm()) {
    AS_LOG_ERROR("AllReduce: the context carries no communicator (CUDAContext::SetB2Comm)");
    return AsStatus::ALLSPARK_RUNTIME_ERROR;
  }
  return AsStatus::ALLSPARK_SUCCESS;
}

AsStatus AllReduceOp::Reshape(RuntimeContext*) {
  Shape s = tensor_map_->at(in_names_[0])->GetShape();
  count_ = s.Count();
  return tensor_map_->at(out_names_[0])->SetShape(std::move(s));
}

AsStatus AllReduceOp::Forward(RuntimeContext*) {
  if (nranks_ == 1) return AsStatus::ALLSPARK_SUCCESS;  // allreduce_op.cpp:78-80 (in and out alias in a one-rank graph)
  const CUDAContext* cc = static_cast<const CUDAContext*>(ctx_);
  void* in = tensor_map_->at(in_names_[0])->GetDataPtr();
  void* out = tensor_map_->at(out_names_[0])->GetDataPtr();
  // no ctx_->Synchronize(): errors of the exchange surface through b2_comm_error / the model's next CUDA error poll
  return FromB2(b2_allreduce(cc->GetB2Comm(), out, in, nullptr, count_, B2_DT_BF16, cc->GetStream()));
}

REGISTER_OP(AllReduce, CUDA, AllReduceOp)

}  // namespace allspark
