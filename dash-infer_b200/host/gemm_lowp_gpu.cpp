#include "gemm_lowp_gpu.h"

namespace allspark {

GemmLowpGPUBase::~GemmLowpGPUBase() {
  if (handle_) b2_gemm_wq_destroy(handle_);
}

AsStatus GemmLowpGPUBase::InitV2(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map,
                                 TensorMap& weights_buffer, TensorMap* tensor_map, RuntimeContext* runtime_ctx) {
  (void)runtime_ctx;
  AS_CHECK_STATUS(AsOperator::Init(op_proto, ctx, weights_map, tensor_map));
  const int wb = wbits();
  // weights: quantized ops carry [weight, scales, zeros, (bias)] (gemm_a16w4.cpp:30-36), dense Gemm [weight, (bias)]
  const size_t nw = weights_.size();
  if (wb != 16 ? (nw != 3 && nw != 4) : (nw != 1 && nw != 2)) {
    AS_LOG_ERROR("%s has %zu weights; expected [weight], [scales], [zeros], (optional) [bias]", op_type_.c_str(), nw);
    return AsStatus::ALLSPARK_PARAM_ERROR;
  }
  DataType dtype = tensor_map_->at(in_names_[0])->GetDataType();
  if (dtype == DATATYPE_UNDEFINED) dtype = ctx.GetDtype();
  tensor_map_->at(out_names_[0])->SetDataType(dtype);

  auto& attr = op_proto.attr();
  if (attr.count("transB")) transB_ = *(const bool*)attr.at("transB").c_str();
  if (attr.count("is_pooler")) is_pooler_ = *(const bool*)attr.at("is_pooler").c_str();
  if (attr.count("activation")) activation_ = *(const int*)attr.at("activation").c_str();
  if (attr.count("alpha")) alpha_ = *(const float*)attr.at("alpha").c_str();
  // dense Gemm only (gemm_op.cpp:73-84): residual fused as a second input, K-split activations
  if (attr.count("binary_type")) binary_type_ = *(const int*)attr.at("binary_type").c_str();
  if (attr.count("splitk")) is_split_k_ = *(const bool*)attr.at("splitk").c_str();
  if (binary_type_ != BINARYTYPE_UNDEFINED && binary_type_ != ADD) {
    AS_LOG_ERROR("%s: binary_type %d is not implemented (only ADD, like GemmOpGPU)", op_type_.c_str(), binary_type_);
    return AsStatus::ALLSPARK_PARAM_ERROR;
  }
  if (binary_type_ == ADD && in_names_.size() < 2) {
    AS_LOG_ERROR("%s: binary_type ADD needs the residual as a second input", op_type_.c_str());
    return AsStatus::ALLSPARK_PARAM_ERROR;
  }
  if (attr.count("GroupSize")) {
    group_size_ = *(const int*)attr.at("GroupSize").c_str();
    // gemm_a16w4.cpp:57-63 (>= 32, % 8) and gemm_a16w8.cpp (64/128/256/512); this build streams k in 64-wide tiles
    if (group_size_ % 8 != 0 || group_size_ < 32) {
      AS_LOG_ERROR("%s: SubChannel only supports GroupSize >= 32 and divisible by 8", op_type_.c_str());
      return AsStatus::ALLSPARK_PARAM_ERROR;
    }
  }
  if (transB_ || is_pooler_) {
    AS_LOG_ERROR("%s: transB / is_pooler are not supported (same as the reference lowp ops)", op_type_.c_str());
    return AsStatus::ALLSPARK_PARAM_ERROR;
  }
  const Shape& ws = weights_[0]->GetShape();
  if (ws.Size() != 2) return AsStatus::ALLSPARK_PARAM_ERROR;
  k_ = ws[0];
  qtype_ = weights_[0]->GetDataType();
  if (wb == 4) {
    n_ = weights_[1]->GetShape()[1];  // the quantization params carry the real N; weight holds (N+1)/2 columns
    if (ws[1] != (n_ + 1) / 2) {
      AS_LOG_ERROR("GemmA16W4: N_PACK size error");
      return AsStatus::ALLSPARK_PARAM_ERROR;
    }
    if (qtype_ != DataType::UINT8) {
      AS_LOG_ERROR("GemmA16W4: packed weight must be uint8 (uint4x2)");
      return AsStatus::ALLSPARK_PARAM_ERROR;
    }
  } else {
    n_ = ws[1];
    if (wb == 8 && qtype_ != DataType::INT8 && qtype_ != DataType::UINT8) return AsStatus::ALLSPARK_PARAM_ERROR;
  }
  b2_gemm_wq_desc d{};
  d.K = (int)k_; d.N = (int)n_; d.wbits = wb; d.group_size = wb == 16 ? -1 : group_size_;
  d.ft = dtype; d.qtype = wb == 16 ? B2_DT_U8 : (int)qtype_;
  d.max_m = ctx.GetModelMaxBatch() > 0 ? ctx.GetModelMaxBatch() : 1024;
  AS_CHECK_STATUS(FromB2(b2_gemm_wq_create(&handle_, &d)));
  cudaStream_t stream = static_cast<const CUDAContext*>(ctx_)->GetStream();
  const void* sc = wb == 16 ? nullptr : weights_[1]->GetDataPtr();
  const void* zr = wb == 16 ? nullptr : weights_[2]->GetDataPtr();
  AS_CHECK_STATUS(FromB2(b2_gemm_wq_prepare_weights(handle_, weights_[0]->GetDataPtr(), sc, zr, nullptr, stream)));
  cudaStreamSynchronize(stream);
  // Weight-swap contract (gemm_a16w4_gpu.cpp:296-300, util::SyncWeightsBuffer): the reference overwrites the op's weight
  // tensors with padded / reordered copies and must mirror them into `weights_buffer` (the host image the swap-in path
  // restores from).  Here the caller's weight tensors are left untouched — the re-laid-out image is handle-owned — so
  // whatever `weights_buffer` holds for them stays valid; after a swap-in the engine calls ReloadWeights() below.
  for (const AsTensor* w : weights_) {
    auto it = weights_buffer.find(w->GetName());
    if (it != weights_buffer.end() && it->second && it->second->GetSizeInByte() != w->GetSizeInByte()) {
      AS_LOG_ERROR("%s: weights_buffer entry '%s' does not match the weight tensor", op_type_.c_str(), w->GetName().c_str());
      return AsStatus::ALLSPARK_PARAM_ERROR;
    }
  }
  return AsStatus::ALLSPARK_SUCCESS;
}

AsStatus GemmLowpGPUBase::ReloadWeights() {
  cudaStream_t stream = static_cast<const CUDAContext*>(ctx_)->GetStream();
  const int wb = wbits();
  const void* sc = wb == 16 ? nullptr : weights_[1]->GetDataPtr();
  const void* zr = wb == 16 ? nullptr : weights_[2]->GetDataPtr();
  AS_CHECK_STATUS(FromB2(b2_gemm_wq_prepare_weights(handle_, weights_[0]->GetDataPtr(), sc, zr, nullptr, stream)));
  cudaStreamSynchronize(stream);
  return AsStatus::ALLSPARK_SUCCESS;
}

AsStatus GemmLowpGPUBase::Reshape() {
  const Shape& xs = tensor_map_->at(in_names_[0])->GetShape();
  const int nd = xs.Size();
  const int nranks = ctx_->GetNranks() > 0 ? ctx_->GetNranks() : 1;
  // splitk (gemm_op.cpp:96-98): the activation rows are nranks * K wide and this rank multiplies its own K-slice
  lda_ = is_split_k_ ? k_ * nranks : k_;
  if (nd < 1 || xs[nd - 1] != lda_) return AsStatus::ALLSPARK_PARAM_ERROR;
  m_ = xs.Count(0, nd - 1);
  Shape ys;
  for (int i = 0; i < nd - 1; ++i) ys.Append(xs[i]);
  ys.Append(n_);
  AsTensor* out = tensor_map_->at(out_names_[0]).get();
  out->SetDataType(tensor_map_->at(in_names_[0])->GetDataType());
  AS_CHECK_STATUS(out->SetShape(std::move(ys)));
  // the shared "workspace" tensor: grow to what this op needs (never shrink: other ops share it)
  auto ws_it = tensor_map_->find("workspace");
  if (ws_it == tensor_map_->end()) return AsStatus::ALLSPARK_PARAM_ERROR;
  const int64_t need = (int64_t)b2_gemm_wq_workspace_bytes(handle_, (int)m_);
  if (ws_it->second->GetDataType() == DATATYPE_UNDEFINED) ws_it->second->SetDataType(DataType::INT8);
  if ((int64_t)ws_it->second->GetSizeInByte() < need) AS_CHECK_STATUS(ws_it->second->SetShape(Shape{need}));
  return AsStatus::ALLSPARK_SUCCESS;
}

AsStatus GemmLowpGPUBase::Forward() {
  AsTensor* in = tensor_map_->at(in_names_[0]).get();
  AsTensor* out = tensor_map_->at(out_names_[0]).get();
  AsTensor* ws = tensor_map_->at("workspace").get();
  const size_t bias_idx = wbits() == 16 ? 1 : 3;
  const void* bias = weights_.size() > bias_idx ? weights_[bias_idx]->GetDataPtr() : nullptr;
  cudaStream_t stream = static_cast<const CUDAContext*>(ctx_)->GetStream();
  // binary_type ADD (gemm_op.cpp:121-136, gemm_op_gpu.cpp): out = act(alpha * x W + bias) + in[1]; under tensor
  // parallelism only rank 0 adds it (the all-reduce that follows sums the partial outputs)
  const void* residual = nullptr;
  if (binary_type_ == ADD && ctx_->GetRank() == 0) {
    AsTensor* res = tensor_map_->at(in_names_[1]).get();
    if (res->GetShape().Count() != m_ * n_) return AsStatus::ALLSPARK_PARAM_ERROR;
    residual = res->GetDataPtr();
  }
  const char* a = (const char*)in->GetDataPtr();
  if (is_split_k_) a += (size_t)ctx_->GetRank() * k_ * SizeofType(in->GetDataType());
  return FromB2(b2_gemm_wq_run(handle_, a, lda_, out->GetDataPtr(), n_, (int)m_, bias, residual, activation_, alpha_,
                               ws->GetDataPtr(), ws->GetSizeInByte(), stream));
}

REGISTER_OP(GemmA16W4, CUDA, GemmA16W4GPU)
REGISTER_OP(GemmA16W8, CUDA, GemmA16W8GPU)
REGISTER_OP(Gemm, CUDA, GemmOpGPU)

}  // namespace allspark
