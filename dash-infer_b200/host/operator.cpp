#include "operator.h"

namespace allspark {

AsStatus AsOperator::Init(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map,
                          TensorMap* tensor_map) {
  tensor_map_ = tensor_map;
  ctx_ = &ctx;
  op_name_ = op_proto.op_name();
  in_names_.clear(); out_names_.clear(); weights_.clear();
  const DeviceType dev = ctx.GetDeviceType();
  for (auto& t : op_proto.inputs()) {
    in_names_.push_back(t.name());
    if (!tensor_map_->count(t.name())) (*tensor_map_)[t.name()] = std::make_shared<AsTensor>(t.name(), dev);
  }
  for (auto& t : op_proto.outputs()) {
    out_names_.push_back(t.name());
    if (!tensor_map_->count(t.name())) (*tensor_map_)[t.name()] = std::make_shared<AsTensor>(t.name(), dev);
  }
  for (auto& t : op_proto.weights()) {
    auto it = weights_map.find(t.name());
    if (it == weights_map.end()) {
      AS_LOG_ERROR("%s: weight %s not found", op_name_.c_str(), t.name().c_str());
      return AsStatus::ALLSPARK_PARAM_ERROR;
    }
    weights_.push_back(it->second.get());
  }
  return AsStatus::ALLSPARK_SUCCESS;
}

OpFactory& OpFactory::getInstance() {
  static OpFactory f;
  return f;
}
OpConstructor OpFactory::GetOperator(const OpRegistType& t) {
  auto it = ops_.find(t);
  if (it == ops_.end()) throw AsException("OpFactory: op " + t.op_type_str + " not registered for this device");
  return it->second;
}
void OpFactory::Register(const OpRegistType& t, OpConstructor c) { ops_[t] = std::move(c); }

}  // namespace allspark
