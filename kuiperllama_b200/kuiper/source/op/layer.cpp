#include "op/layer.h"

#include <cstdarg>
#include <utility>

namespace op {
// ---- BaseLayer -----------------------------------------------------------------------------------
BaseLayer::BaseLayer(base::DeviceType device_type, LayerType layer_type, base::DataType data_type,
                     std::string layer_name)
    : layer_name_(std::move(layer_name)),
      layer_type_(layer_type),
      data_type_(data_type),
      device_type_(device_type) {}

base::DataType BaseLayer::data_type() const { return data_type_; }
LayerType BaseLayer::layer_type() const { return layer_type_; }
const std::string& BaseLayer::get_layer_name() const { return layer_name_; }
void BaseLayer::set_layer_name(const std::string& layer_name) { layer_name_ = layer_name; }
base::DeviceType BaseLayer::device_type() const { return device_type_; }
void BaseLayer::set_device_type(base::DeviceType device_type) { device_type_ = device_type; }

base::Status BaseLayer::set_weight(int32_t, const tensor::Tensor&) {
  return base::error::FunctionNotImplement();
}
base::Status BaseLayer::set_weight(int32_t, const std::vector<int32_t>&, const void*, base::DeviceType) {
  return base::error::FunctionNotImplement();
}

// ---- Layer ---------------------------------------------------------------------------------------
Layer::Layer(base::DeviceType device_type, LayerType layer_type, std::string layer_name)
    : BaseLayer(device_type, layer_type, base::DataType::kDataTypeFp32, std::move(layer_name)) {}

base::Status Layer::init() { return base::error::Success(); }
base::Status Layer::forward() { return base::error::FunctionNotImplement(""); }
base::Status Layer::check() const {
  return base::error::FunctionNotImplement("The check function is not implement yet");
}

base::Status Layer::check_tensor(const tensor::Tensor& tensor, base::DeviceType device_type,
                                 base::DataType data_type) const {
  if (tensor.is_empty()) return base::error::InvalidArgument("The tensor parameter is empty.");
  if (tensor.device_type() != device_type)
    return base::error::InvalidArgument("The tensor has a wrong device type.");
  if (tensor.data_type() != data_type)
    return base::error::InvalidArgument("The tensor has a wrong data type.");
  return base::error::Success();
}

base::Status Layer::check_tensor_with_dim(const tensor::Tensor& tensor, base::DeviceType device_type,
                                          base::DataType data_type, ...) const {
  base::Status basic = check_tensor(tensor, device_type, data_type);
  if (!basic) return basic;
  std::va_list args;
  va_start(args, data_type);
  base::Status result = base::error::Success();
  for (int32_t i = 0; i < tensor.dims_size(); ++i) {
    const int32_t want = va_arg(args, int32_t);
    if (want != tensor.get_dim(i)) {
      result = base::error::InvalidArgument("The tensor has a wrong dim in dim" + std::to_string(i));
      break;
    }
  }
  va_end(args);
  return result;
}

void Layer::set_input(int32_t idx, const tensor::Tensor& input) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(inputs_.size()));
  inputs_[idx] = input;
}
void Layer::set_output(int32_t idx, const tensor::Tensor& output) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(outputs_.size()));
  outputs_[idx] = output;
}
const tensor::Tensor& Layer::get_input(int32_t idx) const {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(inputs_.size()));
  return inputs_[idx];
}
tensor::Tensor& Layer::get_input(int32_t idx) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(inputs_.size()));
  return inputs_[idx];
}
const tensor::Tensor& Layer::get_output(int32_t idx) const {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(outputs_.size()));
  return outputs_[idx];
}
tensor::Tensor& Layer::get_output(int32_t idx) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(outputs_.size()));
  return outputs_[idx];
}
size_t Layer::input_size() const { return inputs_.size(); }
size_t Layer::output_size() const { return outputs_.size(); }
void Layer::reset_input_size(size_t size) { inputs_.resize(size); }
void Layer::reset_output_size(size_t size) { outputs_.resize(size); }

void Layer::to_cuda() {
  cudaStream_t s = cuda_config_ ? cuda_config_->stream : nullptr;
  for (auto& t : inputs_)
    if (!t.is_empty()) t.to_cuda(s);
  for (auto& t : outputs_)
    if (!t.is_empty()) t.to_cuda(s);
}

void Layer::set_cuda_config(std::shared_ptr<kernel::CudaConfig> config) {
  if (config) cuda_config_ = std::move(config);
}
std::shared_ptr<kernel::CudaConfig> Layer::cuda_config() const { return cuda_config_; }

// the N-input overloads only park the tensors in the slots and run forward()
base::Status Layer::forward(const tensor::Tensor& input1, const tensor::Tensor& output1) {
  set_input(0, input1);
  set_output(0, output1);
  return forward();
}
base::Status Layer::forward(const tensor::Tensor& input1, const tensor::Tensor& input2,
                            const tensor::Tensor& output1) {
  set_input(0, input1);
  set_input(1, input2);
  set_output(0, output1);
  return forward();
}
base::Status Layer::forward(const tensor::Tensor& input1, const tensor::Tensor& input2,
                            const tensor::Tensor& input3, const tensor::Tensor& output1) {
  set_input(0, input1);
  set_input(1, input2);
  set_input(2, input3);
  set_output(0, output1);
  return forward();
}
base::Status Layer::forward(const tensor::Tensor& input1, const tensor::Tensor& input2,
                            const tensor::Tensor& input3, const tensor::Tensor& input4,
                            const tensor::Tensor& output1) {
  set_input(0, input1);
  set_input(1, input2);
  set_input(2, input3);
  set_input(3, input4);
  set_output(0, output1);
  return forward();
}
base::Status Layer::forward(const tensor::Tensor& input1, const tensor::Tensor& input2,
                            const tensor::Tensor& input3, const tensor::Tensor& input4,
                            const tensor::Tensor& input5, const tensor::Tensor& output1) {
  set_input(0, input1);
  set_input(1, input2);
  set_input(2, input3);
  set_input(3, input4);
  set_input(4, input5);
  set_output(0, output1);
  return forward();
}

// ---- LayerParam ----------------------------------------------------------------------------------
LayerParam::LayerParam(base::DeviceType device_type, LayerType layer_type, bool is_quant_layer,
                       std::string layer_name)
    : Layer(device_type, layer_type, std::move(layer_name)), is_quant_layer_(is_quant_layer) {}

size_t LayerParam::weight_size() const { return weights_.size(); }
void LayerParam::reset_weight_size(size_t size) { weights_.resize(size); }

tensor::Tensor& LayerParam::get_weight(int32_t idx) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(weights_.size()));
  return weights_[idx];
}
const tensor::Tensor& LayerParam::get_weight(int32_t idx) const {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(weights_.size()));
  return weights_[idx];
}

void LayerParam::to_cuda() {
  Layer::to_cuda();
  cudaStream_t s = cuda_config_ ? cuda_config_->stream : nullptr;
  for (auto& w : weights_) w.to_cuda(s);
  if (!scales_.is_empty()) scales_.to_cuda(s);
}

base::Status LayerParam::set_weight(int32_t idx, const tensor::Tensor& weight) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(weights_.size()));
  CHECK(weight.data_type() == base::DataType::kDataTypeFp32);
  if (!weight.is_empty()) CHECK(weight.device_type() == device_type_);
  weights_[idx] = weight;
  return base::error::Success();
}

base::Status LayerParam::set_weight(int32_t idx, const std::vector<int32_t>& dims,
                                    const void* weight_ptr, base::DeviceType device_type) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(weights_.size()));
  CHECK_NE(weight_ptr, nullptr);
  size_t numel = 1;
  for (int32_t d : dims) numel *= static_cast<size_t>(d);
  const base::DataType dt = is_quant_layer_ ? base::DataType::kDataTypeInt8 : base::DataType::kDataTypeFp32;
  auto view = std::make_shared<base::Buffer>(numel * base::DataTypeSize(dt), nullptr,
                                             const_cast<void*>(weight_ptr), true);
  if (device_type != base::DeviceType::kDeviceUnknown) view->set_device_type(device_type);
  tensor::Tensor weight(dt, dims);
  CHECK(weight.assign(view));
  weights_[idx] = weight;

  if (is_quant_layer_) {
    CHECK(group_size_ > 0 && numel % static_cast<size_t>(group_size_) == 0);
    const int32_t n_scales = static_cast<int32_t>(numel / static_cast<size_t>(group_size_));
    const auto* after = static_cast<const int8_t*>(weight_ptr) + numel;
    scales_ = tensor::Tensor(base::DataType::kDataTypeFp32, n_scales, false, nullptr,
                             const_cast<int8_t*>(after));
    scales_.set_device_type(device_type);
  }
  return base::error::Success();
}

void LayerParam::set_scales(const tensor::Tensor& scales) {
  CHECK(!scales.is_empty());
  scales_ = scales;
}
void LayerParam::set_group_size(int32_t group_size) { group_size_ = group_size; }
int32_t LayerParam::get_scale_num() const {
  CHECK(!scales_.is_empty());
  return static_cast<int32_t>(scales_.size());
}
}  // namespace op
