#include "span_attn_op_cuda.h"

#include <cmath>
#include <cstring>
#include <string>

namespace allspark {

SpanAttnOpCUDA::~SpanAttnOpCUDA() {
  if (handle_) b2_span_attn_destroy(handle_);
  for (auto& s : stage_) {
    if (s.done) cudaEventDestroy(s.done);
    if (s.host) cudaFreeHost(s.host);
  }
}

// csrc/common/common.h:239-257: the first '.'-separated field of the op name that is all digits
// ("decoder.layer.17.attention" -> 17); -1 when there is none
static int layer_num_from_name(const std::string& name) {
  size_t a = 0;
  while (a <= name.size()) {
    size_t b = name.find('.', a);
    if (b == std::string::npos) b = name.size();
    bool digits = b > a;
    for (size_t i = a; i < b; ++i) digits = digits && name[i] >= '0' && name[i] <= '9';
    if (digits) return std::stoi(name.substr(a, b - a));
    a = b + 1;
  }
  return -1;
}

AsStatus SpanAttnOpCUDA::InitV2(const OperatorProto& op_proto, const DeviceContext& ctx, const TensorMap& weights_map,
                                TensorMap& weights_buffer, TensorMap* tensor_map, RuntimeContext* runtime_ctx) {
  (void)weights_buffer; (void)runtime_ctx;
  AS_CHECK_STATUS(AsOperator::Init(op_proto, ctx, weights_map, tensor_map));
  auto& attr = op_proto.attr();
  if (attr.count("alpha")) alpha_ = *(const float*)attr.at("alpha").c_str();
  layer_num_ = layer_num_from_name(op_name_);  // span_attn_op.cpp:183
  if (attr.count("layer_num")) layer_num_ = *(const int*)attr.at("layer_num").c_str();  // optional override (tests)
  if (layer_num_ < 0) {
    AS_LOG_ERROR("SpanAttnOp: cannot get layer_num from op name '%s'", op_name_.c_str());
    return AsStatus::ALLSPARK_PARAM_ERROR;
  }
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
  // pinned staging ring: [old lens | new lens | k table | v table] per slot; a slot is reused only after the copies that
  // read it have completed (event), so Forward never blocks on the stream in steady state
  slot_bytes_ = (size_t)2 * max_batch_ * sizeof(int32_t) + (size_t)2 * max_batch_ * max_spans_ * sizeof(void*);
  for (auto& sl : stage_) {
    if (cudaMallocHost(&sl.host, slot_bytes_) != cudaSuccess || cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) != cudaSuccess)
      return AsStatus::ALLSPARK_MEMORY_ERROR;
    sl.busy = false;
  }
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
  Stage& sl = stage_[stage_i_];
  stage_i_ = (stage_i_ + 1) % kStages;
  if (sl.busy) cudaEventSynchronize(sl.done);  // kStages Forward calls ago: long finished unless the host runs far ahead
  int32_t* old_h = reinterpret_cast<int32_t*>(sl.host);
  int32_t* new_h = old_h + max_batch_;
  void** k_h = reinterpret_cast<void**>(new_h + max_batch_);
  void** v_h = k_h + (size_t)max_batch_ * max_spans_;
  // host -> device staging only for what changed: lengths (4 B / sequence) and the span-table rows that grew
  int dirty_lo = batch_size_, dirty_hi = -1;
  for (int b = 0; b < batch_size_; ++b) {
    GenerateContext* g = runtime_ctx->GetGenCtx(b);
    old_h[b] = g->step;
    new_h[b] = g->step + 1;
    const int ns = (int)g->virtual_k_cache->GetCache(layer_num_, 0).GetShape().Count();
    if (ns > max_spans_ || (int)g->virtual_v_cache->GetCache(layer_num_, 0).GetShape().Count() != ns)
      return AsStatus::ALLSPARK_EXCEED_LIMIT_ERROR;
    if (ns != span_counts_[b]) {
      span_counts_[b] = ns;
      dirty_lo = b < dirty_lo ? b : dirty_lo;
      dirty_hi = b;
    }
  }
  if (dirty_hi >= 0) {  // stage and upload the contiguous row range that contains every changed row
    for (int b = dirty_lo; b <= dirty_hi; ++b) {
      GenerateContext* g = runtime_ctx->GetGenCtx(b);
      std::memcpy(k_h + (size_t)b * max_spans_, g->virtual_k_cache->GetCache(layer_num_, 0).GetDataPtr(), span_counts_[b] * sizeof(void*));
      std::memcpy(v_h + (size_t)b * max_spans_, g->virtual_v_cache->GetCache(layer_num_, 0).GetDataPtr(), span_counts_[b] * sizeof(void*));
    }
    const size_t off = (size_t)dirty_lo * max_spans_, cnt = (size_t)(dirty_hi - dirty_lo + 1) * max_spans_;
    cudaMemcpyAsync((void**)k_tab_->GetDataPtr() + off, k_h + off, cnt * sizeof(void*), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync((void**)v_tab_->GetDataPtr() + off, v_h + off, cnt * sizeof(void*), cudaMemcpyHostToDevice, stream);
  }
  cudaMemcpyAsync(old_lens_->GetDataPtr(), old_h, batch_size_ * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(new_lens_->GetDataPtr(), new_h, batch_size_ * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
  cudaEventRecord(sl.done, stream);  // no stream synchronisation: the slot is only reused kStages calls later
  sl.busy = true;

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
