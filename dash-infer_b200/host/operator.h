// AsOperator / OpFactory / REGISTER_OP — the operator contract of csrc/core/operator/operator.h:38-201, with the
// protobuf OperatorProto replaced by a plain struct that keeps the same accessor names (op_type(), attr(), ...;
// protoc is not available in this build image).  Attributes travel as raw bytes exactly like in the reference
// (`*(T*)attr.at(k).c_str()`, gemm_a16w4.cpp:44-64).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "cache.h"
#include "device_context.h"
#include "tensor.h"

namespace allspark {

struct TensorProto {
  std::string name_;
  const std::string& name() const { return name_; }
};

struct OperatorProto {
  std::string op_type_, op_name_;
  std::vector<TensorProto> inputs_, outputs_, weights_;
  std::map<std::string, std::string> attr_;
  const std::string& op_type() const { return op_type_; }
  const std::string& op_name() const { return op_name_; }
  const std::vector<TensorProto>& inputs() const { return inputs_; }
  const std::vector<TensorProto>& outputs() const { return outputs_; }
  const std::vector<TensorProto>& weights() const { return weights_; }
  const std::map<std::string, std::string>& attr() const { return attr_; }
  template <typename T>
  void SetAttr(const std::string& k, T v) { attr_[k] = std::string(reinterpret_cast<const char*>(&v), sizeof(T)); }
};

class AsOperator {
 public:
  explicit AsOperator(const std::string& op_type = "") : op_type_(op_type) {}
  virtual ~AsOperator() = default;

  // model-facing entry points (operator.h:43-53)
  AsStatus CallForward(RuntimeContext* runtime_ctx) { return runtime_ctx ? Forward(runtime_ctx) : Forward(); }
  AsStatus CallReshape(RuntimeContext* runtime_ctx) { return runtime_ctx ? Reshape(runtime_ctx) : Reshape(); }
  AsStatus CallAlloc(RuntimeContext* runtime_ctx) { return Alloc(runtime_ctx); }

  virtual AsStatus InitV2(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map,
                          TensorMap& weights_buffer, TensorMap* tensor_map, RuntimeContext* runtime_ctx) {
    (void)weights_buffer; (void)runtime_ctx;
    return Init(op_proto, ctx, weights_map, tensor_map);
  }
  // public like the reference tests use them (op.Reshape(); op.Forward(); operator_gemm_lowp_test.cpp:708-713)
  virtual AsStatus Forward() { return AsStatus::ALLSPARK_INVALID_CALL_ERROR; }
  virtual AsStatus Forward(RuntimeContext*) { return Forward(); }
  virtual AsStatus Reshape() { return AsStatus::ALLSPARK_INVALID_CALL_ERROR; }
  virtual AsStatus Reshape(RuntimeContext*) { return Reshape(); }
  virtual AsStatus Alloc(RuntimeContext*) { return AsStatus::ALLSPARK_SUCCESS; }

  const std::string& GetOpType() const { return op_type_; }
  const std::string& GetOpName() const { return op_name_; }
  const std::vector<std::string>& GetInNames() const { return in_names_; }
  const std::vector<std::string>& GetOutNames() const { return out_names_; }

 protected:
  // resolves input/output names, creates missing tensors on the context's device, collects weights in proto order
  // (operator.cpp:330-336: weights come from weights_map by name)
  virtual AsStatus Init(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map,
                        TensorMap* tensor_map);

  std::string op_type_, op_name_;
  std::vector<std::string> in_names_, out_names_;
  std::vector<AsTensor*> weights_;
  TensorMap* tensor_map_ = nullptr;
  const DeviceContext* ctx_ = nullptr;
};

struct OpRegistType {
  std::string op_type_str;
  DeviceType device_type;
  bool operator==(const OpRegistType& o) const { return op_type_str == o.op_type_str && device_type == o.device_type; }
};
struct OpRegistTypeHash {
  size_t operator()(const OpRegistType& p) const { return std::hash<std::string>{}(p.op_type_str) * 31 + (size_t)p.device_type; }
};
using OpConstructor = std::function<std::unique_ptr<AsOperator>()>;

class OpFactory {
 public:
  static OpFactory& getInstance();
  OpConstructor GetOperator(const OpRegistType& t);
  void Register(const OpRegistType& t, OpConstructor c);

 private:
  std::unordered_map<OpRegistType, OpConstructor, OpRegistTypeHash> ops_;
};

struct OpRegisterHelper {
  OpRegisterHelper(const OpRegistType& t, OpConstructor c) { OpFactory::getInstance().Register(t, std::move(c)); }
};

#define REGISTER_OP(op_name, device_type, typed_class)                                        \
  static ::allspark::OpRegisterHelper op_name##_##typed_class##Register##_##device_type(      \
      ::allspark::OpRegistType{#op_name, ::allspark::DeviceType::device_type},                \
      []() -> std::unique_ptr<::allspark::AsOperator> { return std::make_unique<typed_class>(#op_name); });

}  // namespace allspark
