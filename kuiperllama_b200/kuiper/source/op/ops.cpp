// The compute layers: thin shape/device validation + registry dispatch
// (reference kuiper/source/op/{add,swiglu,rmsnorm,rope,mha,matmul,embedding}.cpp).
#include "kernels/kernels_interface.h"
#include "op/decoder_layers.h"

namespace op {
namespace {
inline void* stream_of(const std::shared_ptr<kernel::CudaConfig>& c) { return c ? c->stream : nullptr; }
}  // namespace

// ---- VecAdd --------------------------------------------------------------------------------------
VecAddLayer::VecAddLayer(base::DeviceType device_type) : Layer(device_type, LayerType::kLayerAdd, "Add") {
  reset_input_size(2);
  reset_output_size(1);
}
base::Status VecAddLayer::check() const {
  const int32_t n = static_cast<int32_t>(get_input(0).size());
  for (const tensor::Tensor* t : {&get_input(0), &get_input(1), &get_output(0)}) {
    base::Status st = check_tensor_with_dim(*t, device_type_, data_type_, n);
    if (!st) {
      LOG(ERROR) << "A tensor of the add layer has the wrong shape, type or device.";
      return st;
    }
  }
  return base::error::Success();
}
base::Status VecAddLayer::forward() {
  base::Status st = check();
  if (!st) return st;
  if (device_type_ == base::DeviceType::kDeviceCUDA) CHECK(cuda_config_ != nullptr);
  kernel::get_add_kernel(device_type_)(get_input(0), get_input(1), get_output(0), stream_of(cuda_config_));
  return base::error::Success();
}

// ---- SwiGLU --------------------------------------------------------------------------------------
SwiGLULayer::SwiGLULayer(base::DeviceType device_type, int32_t hidden_dim)
    : Layer(device_type, LayerType::kLayerSwiGLU, "SwiGLU"), hidden_dim_(hidden_dim) {
  reset_input_size(2);
  reset_output_size(1);
}
base::Status SwiGLULayer::check() const {
  for (const tensor::Tensor* t : {&get_input(0), &get_input(1), &get_output(0)}) {
    base::Status st = check_tensor_with_dim(*t, device_type_, data_type_, hidden_dim_);
    if (!st) {
      LOG(ERROR) << "A tensor of the swiglu layer has the wrong shape, type or device.";
      return st;
    }
  }
  return base::error::Success();
}
base::Status SwiGLULayer::forward() {
  base::Status st = check();
  if (!st) return st;
  if (device_type_ == base::DeviceType::kDeviceCUDA) CHECK(cuda_config_ != nullptr);
  kernel::get_swiglu_kernel(device_type_)(get_input(0), get_input(1), get_output(0), stream_of(cuda_config_));
  return base::error::Success();
}

// ---- RMSNorm -------------------------------------------------------------------------------------
RmsNormLayer::RmsNormLayer(base::DeviceType device_type, int32_t dim)
    : LayerParam(device_type, LayerType::kLayerRMSNorm, false, "RMSNorm"), dim_(dim) {
  reset_input_size(1);
  reset_output_size(1);
  reset_weight_size(1);
}
base::Status RmsNormLayer::check() const {
  for (const tensor::Tensor* t : {&get_input(0), &get_weight(0), &get_output(0)}) {
    base::Status st = check_tensor_with_dim(*t, device_type_, data_type_, dim_);
    if (!st) {
      LOG(ERROR) << "A tensor of the rmsnorm layer has the wrong shape, type or device.";
      return st;
    }
  }
  return base::error::Success();
}
base::Status RmsNormLayer::forward() {
  base::Status st = check();
  if (!st) return st;
  if (device_type_ == base::DeviceType::kDeviceCUDA) CHECK(cuda_config_ != nullptr);
  kernel::get_rmsnorm_kernel(device_type_)(get_input(0), get_weight(0), get_output(0), stream_of(cuda_config_));
  return base::error::Success();
}

// ---- RoPE ----------------------------------------------------------------------------------------
RoPELayer::RoPELayer(base::DeviceType device_type, int32_t dim, int32_t kv_dim, int32_t head_size)
    : Layer(device_type, LayerType::kLayerRoPe, "RoPe"), dim_(dim), kv_dim_(kv_dim), head_size_(head_size) {
  reset_input_size(5);
  reset_output_size(1);
}
base::Status RoPELayer::check() const {
  base::Status st = check_tensor_with_dim(get_input(2), base::DeviceType::kDeviceCPU,
                                          base::DataType::kDataTypeInt32, 1);
  if (!st) {
    LOG(ERROR) << "The position tensor of the rope layer must be a CPU int32 [1].";
    return st;
  }
  st = check_tensor_with_dim(get_input(1), device_type_, data_type_, kv_dim_);
  if (!st) {
    LOG(ERROR) << "The key tensor of the rope layer is wrong.";
    return st;
  }
  st = check_tensor_with_dim(get_input(0), device_type_, data_type_, dim_);
  if (!st) LOG(ERROR) << "The query tensor of the rope layer is wrong.";
  return st;
}
base::Status RoPELayer::forward() {
  base::Status st = check();
  if (!st) return st;
  if (device_type_ == base::DeviceType::kDeviceCUDA) CHECK(cuda_config_ != nullptr);
  kernel::get_rope_kernel(device_type_)(dim_, kv_dim_, head_size_, get_input(0), get_input(1), get_input(2),
                                        get_input(3), get_input(4), stream_of(cuda_config_));
  return base::error::Success();
}

// ---- MHA -----------------------------------------------------------------------------------------
MultiHeadAttention::MultiHeadAttention(base::DeviceType device_type, int32_t layer_index, int32_t kv_mul,
                                       int32_t kv_dim, int32_t seq_len, int32_t head_num, int32_t head_size)
    : Layer(device_type, LayerType::kLayerMHA, "MultiHead"),
      layer_index_(layer_index),
      kv_mul_(kv_mul),
      kv_dim_(kv_dim),
      seq_len_(seq_len),
      head_num_(head_num),
      head_size_(head_size) {
  reset_input_size(5);
  reset_output_size(1);
}
base::Status MultiHeadAttention::check() const {
  for (int32_t i = 0; i < 4; ++i) {  // query, score, key cache, value cache
    base::Status st = check_tensor(get_input(i), device_type_, data_type_);
    if (!st) {
      LOG(ERROR) << "The input tensor " << i << " error in the mha layer.";
      return st;
    }
  }
  return check_tensor(get_output(0), device_type_, data_type_);
}
base::Status MultiHeadAttention::forward() {
  base::Status st = check();
  if (!st) return st;
  if (device_type_ == base::DeviceType::kDeviceCUDA) CHECK(cuda_config_ != nullptr);
  kernel::get_mha_kernel(device_type_)(pos_, head_num_, layer_index_, seq_len_, kv_dim_, kv_mul_, head_size_,
                                       get_output(0), get_input(0), get_input(1), get_input(2), get_input(3),
                                       device_type_, cuda_config_ ? cuda_config_.get() : nullptr);
  return base::error::Success();
}

// ---- Matmul --------------------------------------------------------------------------------------
MatmulLayer::MatmulLayer(base::DeviceType device_type, int32_t dim0, int32_t dim1, bool is_quant_layer,
                         bool has_bias)
    : LayerParam(device_type, LayerType::kLayerMatmul, is_quant_layer, "Matmul"),
      dim0_(dim0),
      dim1_(dim1),
      has_bias_(has_bias) {
  reset_input_size(1);
  reset_output_size(1);
  reset_weight_size(1);
  if (has_bias_) bias_.resize(1);
}
base::Status MatmulLayer::check() const {
  base::Status st = check_tensor_with_dim(get_input(0), device_type_, data_type_, dim1_);
  if (!st) {
    LOG(ERROR) << "The input tensor error in the matmul layer.";
    return st;
  }
  st = check_tensor_with_dim(get_weight(0), device_type_,
                             is_quant_layer_ ? base::DataType::kDataTypeInt8 : data_type_, dim0_, dim1_);
  if (!st) {
    LOG(ERROR) << "The weight tensor error in the matmul layer.";
    return st;
  }
  if (is_quant_layer_) {
    st = check_tensor_with_dim(scales_, device_type_, base::DataType::kDataTypeFp32,
                               static_cast<int32_t>(scales_.size()));
    if (!st) {
      LOG(ERROR) << "The scale tensor error in the matmul layer.";
      return st;
    }
  }
  st = check_tensor_with_dim(get_output(0), device_type_, data_type_, dim0_);
  if (!st) LOG(ERROR) << "The output tensor error in the matmul layer.";
  return st;
}
base::Status MatmulLayer::forward() {
  base::Status st = check();
  if (!st) return st;
  if (device_type_ == base::DeviceType::kDeviceCUDA) CHECK(cuda_config_ != nullptr);
  const kernel::CudaConfig* cfg = cuda_config_ ? cuda_config_.get() : nullptr;
  if (is_quant_layer_) {
    kernel::get_matmul_kernel_quant8(device_type_)(get_input(0), get_weight(0), get_output(0), group_size_,
                                                   scales_, cfg);
  } else {
    kernel::get_matmul_kernel(device_type_)(get_input(0), get_weight(0), get_output(0), 1.f, cfg);
  }
  if (has_bias_)
    kernel::get_add_kernel(device_type_)(get_output(0), get_bias(0), get_output(0), stream_of(cuda_config_));
  return base::error::Success();
}
base::Status MatmulLayer::set_bias(int32_t idx, int32_t& dim, const void* bias_ptr, base::DeviceType device_type) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(bias_.size()));
  CHECK_NE(bias_ptr, nullptr);
  // Biases are fp32 vectors in every format the exporters write (export_qwen2.py:103-110); the
  // reference's int8 branch here reinterprets them as int8 + scales, which no file provides.
  auto view = std::make_shared<base::Buffer>(static_cast<size_t>(dim) * sizeof(float), nullptr,
                                             const_cast<void*>(bias_ptr), true);
  if (device_type != base::DeviceType::kDeviceUnknown) view->set_device_type(device_type);
  tensor::Tensor bias(base::DataType::kDataTypeFp32, dim);
  CHECK(bias.assign(view));
  bias_[idx] = bias;
  return base::error::Success();
}
tensor::Tensor& MatmulLayer::get_bias(int32_t idx) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(bias_.size()));
  return bias_[idx];
}
const tensor::Tensor& MatmulLayer::get_bias(int32_t idx) const {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(bias_.size()));
  return bias_[idx];
}
void MatmulLayer::to_cuda() {
  LayerParam::to_cuda();
  for (auto& b : bias_) b.to_cuda(cuda_config_ ? cuda_config_->stream : nullptr);
}

// ---- Embedding -----------------------------------------------------------------------------------
EmbeddingLayer::EmbeddingLayer(base::DeviceType device_type, int32_t dim, int32_t seq_len, int32_t vocab_size)
    : LayerParam(device_type, LayerType::kLayerEmbedding, false, "Embedding"),
      dim_(dim),
      seq_len_(seq_len),
      vocab_size_(vocab_size) {
  reset_weight_size(1);
  reset_input_size(2);
  reset_output_size(1);
}
base::Status EmbeddingLayer::check() const {
  const tensor::Tensor& ids = get_input(0);
  const int32_t n = static_cast<int32_t>(get_input(1).size());
  if (static_cast<size_t>(n) > ids.size())
    return base::error::InvalidArgument("The number of input tensor is greater than seq len.");
  base::Status st = check_tensor_with_dim(ids, base::DeviceType::kDeviceCPU, base::DataType::kDataTypeInt32, n);
  if (!st) {
    LOG(ERROR) << "The input tensor error in the embedding layer.";
    return st;
  }
  st = check_tensor_with_dim(get_weight(0), device_type_, data_type_, vocab_size_, dim_);
  if (!st) {
    LOG(ERROR) << "The weight tensor error in the embedding layer.";
    return st;
  }
  st = check_tensor_with_dim(get_output(0), device_type_, data_type_, n, dim_);
  if (!st) LOG(ERROR) << "The output tensor error in the embedding layer.";
  return st;
}
base::Status EmbeddingLayer::forward() {
  base::Status st = check();
  if (!st) return st;
  if (device_type_ == base::DeviceType::kDeviceCUDA) CHECK(cuda_config_ != nullptr);
  kernel::get_emb_kernel(device_type_)(get_input(0), get_weight(0), get_output(0), vocab_size_,
                                       stream_of(cuda_config_));
  return base::error::Success();
}
}  // namespace op
