#include "span_attn_op_cuda.h"

#include <cmath>

namespace allspark {

SpanAttnOpCUDA::~SpanAttnOpCUDA() {
  if (handle_) b2_span_attn_destroy(handle_);
}

AsStatus SpanAttnOpCUDA::InitV2(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map,
                                TensorMap& weights_buffer, TensorMap* tensor_map, RuntimeContext* runtime_ctx) {
  (void)weights_buffer; (void)runtime_ctx;
  AS_CHECK_STATUS(AsOperator::Init(op_proto, ctx, weights_map, tensor_map));
  auto& attr = op_proto.attr();
  if (attr.count("alpha")) alpha_ = *(const float*)attr.at("alpha").c_str();
  if (attr.count("layer_num")) layer_num_ = *(const int*)attr.at("layer_num").c_str();  // reference: parsed from the op name
  DataType dtype = ctx.GetDtype() != DATATYPE_UNDEFINED ? ctx.GetDtype() : DataType::BFLOAT16;
  tensor_map_->at(out_names_[0])->SetDataType(dtype);
  const int nranks = ctx.GetNranks() > 0 ? ctx.GetNranks() : 1;
  cfg_.ft = dtype;
  cfg_.quant_mode = (int)ctx.GetCacheMode();
  cfg_.n_heads = ctx.GetNumberHeads() / nranks;                     // heads are split across TP ranks (head_gqa.h:29-50)
  cfg_.n_groups = (ctx.GetNumberGroups() > 0 ? ctx.GetNumberGroups() : ctx.GetNumberHeads()) / nranks;
  cfg_.head_size = ctx.GetSizePerHead();
  cfg_.span_len = ctx.GetCacheSpanSize();
  max_spans_ = (ctx.GetModelMaxLength() + cfg_.span_len - 1) / cfg_.span_len;
  cfg_.max_spans_per_seq = max_spans_;
  max_batch_ = ctx.GetModelMaxBatch();
  if (alpha_ < 0) alpha_ = 1.0f / std::sqrt((float)cfg_.head_size);
  AS_CHECK_STATUS(FromB2(b2_span_attn_create(&handle_, &cfg_, max_batch_)));
  const DeviceType dev = DeviceType::CUDA;
  q_tensor_ = std::make_unique<AsTensor>("decoder_q", dev, dtype, DataMode::DENSE, Shape{max_batch_, cfg_.n_heads * cfg_.head_size});
  k_tab_ = std::make_unique<AsTensor>("k_span_array", dev, DataType::POINTER, DataMode::DENSE, Shape{(int64_t)max_batch_ * max_spans_});
  v_tab_ = std::make_unique<AsTensor>("v_span_array", dev, DataType::POINTER, DataMode::DENSE, Shape{(int64_t)max_batch_ * max_spans_});
  old_lens_ = std::make_unique<AsTensor>("old_seq_lens", dev, DataType::INT32, DataMode::DENSE, Shape{max_batch_});
  new_lens_ = std::make_unique<AsTensor>("new_seq_lens", dev, DataType::INT32, DataMode::DENSE, Shape{max_batch_});
  k_host_.assign((size_t)max_batch_ * max_spans_, nullptr);
  v_host_.assign((size_t)max_batch_ * max_spans_, nullptr);
  span_counts_.assign(max_batch_, -1);
  return AsStatus::ALLSPARK_SUCCESS;
}

AsStatus SpanAttnOpCUDA::Reshape(RuntimeContext* runtime_ctx) {
  if (!runtime_ctx || runtime_ctx->is_context) {
    AS_LOG_ERROR("SpanAttnOpCUDA: only the decode branch is built on this path (prefill is out of scope)");
    return AsStatus::ALLSPARK_INVALID_CALL_ERROR;
  }
  const Shape& xs = tensor_map_->at(in_names_[0])->GetShape();  // [batch, 1, (nH + 2 nG) * head]
  batch_size_ = (int)xs[0];
  if (batch_size_ > max_batch_) return AsStatus::ALLSPARK_EXCEED_LIMIT_ERROR;
  const int64_t width = xs[xs.Size() - 1];
  if (width != (int64_t)(cfg_.n_heads + 2 * cfg_.n_groups) * cfg_.head_size) return AsStatus::ALLSPARK_PARAM_ERROR;
  AS_CHECK_STATUS(tensor_map_->at(out_names_[0])->SetShape(Shape{batch_size_, 1, (int64_t)cfg_.n_heads * cfg_.head_size}));
  auto ws_it = tensor_map_->find("workspace");
  if (ws_it == tensor_map_->end()) return AsStatus::ALLSPARK_PARAM_ERROR;
  const int64_t need = (int64_t)b2_span_attn_workspace_bytes(handle_, batch_size_, ctx_->GetModelMaxLength());
  if (ws_it->second->GetDataType() == DATATYPE_UNDEFINED) ws_it->second->SetDataType(DataType::INT8);
  if ((int64_t)ws_it->second->GetSizeInByte() < need) AS_CHECK_STATUS(ws_it->second->SetShape(Shape{need}));
  std::fill(span_counts_.begin(), span_counts_.end(), -1);  // batch membership may have changed
  return AsStatus::ALLSPARK_SUCCESS;
}

AsStatus SpanAttnOpCUDA::Alloc(RuntimeContext* runtime_ctx) {
  // claim the span that will hold this step's token (span_attn_op.cpp:315-368)
  for (int b = 0; b < batch_size_; ++b) {
    GenerateContext* g = runtime_ctx->GetGenCtx(b);
    const size_t old_len = (size_t)g->step;
    if (old_len != g->virtual_k_cache->GetSeqLength(layer_num_) || old_len != g->virtual_v_cache->GetSeqLength(layer_num_)) {
      AS_LOG_ERROR("SpanAttnOp: gen_ctx step and cached seq len mismatch (layer %d)", layer_num_);
      return AsStatus::ALLSPARK_RUNTIME_ERROR;
    }
    try {
      (void)g->virtual_k_cache->GetCache(layer_num_, 1);
      (void)g->virtual_v_cache->GetCache(layer_num_, 1);
    } catch (const AsException&) {
      return AsStatus::ALLSPARK_CACHE_MEMORY_OUT;
    }
  }
  return AsStatus::ALLSPARK_SUCCESS;
}

AsStatus SpanAttnOpCUDA::Forward(RuntimeContext* runtime_ctx) {
  cudaStream_t stream = static_cast<const CUDAContext*>(ctx_)->GetStream();
  // host -> device staging only for what changed: lengths (4 B / sequence) and the span table rows that grew
  lens_host_.resize(batch_size_);
  bool tables_dirty = false;
  for (int b = 0; b < batch_size_; ++b) {
    GenerateContext* g = runtime_ctx->GetGenCtx(b);
    lens_host_[b] = g->step;
    const AsTensor& kp = g->virtual_k_cache->GetCache(layer_num_, 0);
    const AsTensor& vp = g->virtual_v_cache->GetCache(layer_num_, 0);
    const int ns = (int)kp.GetShape().Count();
    if (ns != span_counts_[b]) {
      std::memcpy(&k_host_[(size_t)b * max_spans_], kp.GetDataPtr(), ns * sizeof(void*));
      std::memcpy(&v_host_[(size_t)b * max_spans_], vp.GetDataPtr(), ns * sizeof(void*));
      span_counts_[b] = ns;
      tables_dirty = true;
    }
  }
  if (tables_dirty) {
    const size_t bytes = (size_t)batch_size_ * max_spans_ * sizeof(void*);
    cudaMemcpyAsync(k_tab_->GetDataPtr(), k_host_.data(), bytes, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(v_tab_->GetDataPtr(), v_host_.data(), bytes, cudaMemcpyHostToDevice, stream);
  }
  cudaMemcpyAsync(old_lens_->GetDataPtr(), lens_host_.data(), batch_size_ * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
  for (auto& l : lens_host_) l += 1;
  cudaMemcpyAsync(new_lens_->GetDataPtr(), lens_host_.data(), batch_size_ * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
  cudaStreamSynchronize(stream);  // lens_host_/k_host_ are pageable and reused next call

  AsTensor* in = tensor_map_->at(in_names_[0]).get();
  AsTensor* out = tensor_map_->at(out_names_[0]).get();
  AsTensor* ws = tensor_map_->at("workspace").get();
  AS_CHECK_STATUS(FromB2(b2_span_cache_append(&cfg_, (void* const*)k_tab_->GetDataPtr(), (void* const*)v_tab_->GetDataPtr(),
                                              q_tensor_->GetDataPtr(), in->GetDataPtr(), (const int32_t*)old_lens_->GetDataPtr(),
                                              batch_size_, nullptr, stream)));
  return FromB2(b2_span_attn_run(handle_, out->GetDataPtr(), q_tensor_->GetDataPtr(), (const void* const*)k_tab_->GetDataPtr(),
                                 (const void* const*)v_tab_->GetDataPtr(), (const int32_t*)new_lens_->GetDataPtr(), batch_size_,
                                 ctx_->GetModelMaxLength(), ws->GetDataPtr(), ws->GetSizeInByte(), alpha_, stream));
}

REGISTER_OP(DecOptMHA, CUDA, SpanAttnOpCUDA)
REGISTER_OP(DecOptMQA, CUDA, SpanAttnOpCUDA)

}  // namespace allspark
