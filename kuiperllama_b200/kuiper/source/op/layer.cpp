// op::Layer / op::LayerParam (see op/layer.h).
#include "op/layer.h"

#include <cstdarg>
#include <utility>

namespace op {
namespace {
template <typename Vec>
auto& slot(Vec& v, int32_t idx) {
  CHECK_GE(idx, 0);
  CHECK_LT(idx, static_cast<int32_t>(v.size()));
  return v[idx];
}
cudaStream_t stream_of(const std::shared_ptr<kernel::CudaConfig>& cfg) { return cfg ? cfg->stream : nullptr; }
}  // namespace

Layer::Layer(base::DeviceType device_type, LayerType layer_type, std::string layer_name)
    : layer_name_(std::move(layer_name)), layer_type_(layer_type), device_type_(device_type) {}

void Layer::set_input(int32_t idx, const Tensor& input) { slot(in_slots_, idx) = input; }
void Layer::set_output(int32_t idx, const Tensor& output) { slot(out_slots_, idx) = output; }
tensor::Tensor& Layer::get_input(int32_t idx) { return slot(in_slots_, idx); }
tensor::Tensor& Layer::get_output(int32_t idx) { return slot(out_slots_, idx); }
const tensor::Tensor& Layer::get_input(int32_t idx) const { return slot(in_slots_, idx); }
const tensor::Tensor& Layer::get_output(int32_t idx) const { return slot(out_slots_, idx); }

base::Status Layer::init() { return base::error::Success(); }
base::Status Layer::forward() { return base::error::FunctionNotImplement(""); }
base::Status Layer::check() const {
  return base::error::FunctionNotImplement("The check function is not implement yet");
}

base::Status Layer::bind_and_run(std::initializer_list<const Tensor*> inputs, const Tensor& output) {
  int32_t idx = 0;
  for (const Tensor* t : inputs) set_input(idx++, *t);
  set_output(0, output);
  return forward();
}

void Layer::to_cuda() {
  for (auto* slots : {&in_slots_, &out_slots_})
    for (Tensor& t : *slots)
      if (!t.is_empty()) t.to_cuda(stream_of(cuda_config_));
}
void Layer::set_cuda_config(std::shared_ptr<kernel::CudaConfig> config) {
  if (config) cuda_config_ = std::move(config);
}

base::Status Layer::set_weight(int32_t, const Tensor&) { return base::error::FunctionNotImplement(); }
base::Status Layer::set_weight(int32_t, const std::vector<int32_t>&, const void*, base::DeviceType) {
  return base::error::FunctionNotImplement();
}

base::Status Layer::check_tensor(const Tensor& tensor, base::DeviceType device_type,
                                 base::DataType data_type) const {
  if (tensor.is_empty()) return base::error::InvalidArgument("The tensor parameter is empty.");
  if (tensor.device_type() != device_type) return base::error::InvalidArgument("The tensor has a wrong device type.");
  if (tensor.data_type() != data_type) return base::error::InvalidArgument("The tensor has a wrong data type.");
  return base::error::Success();
}

base::Status Layer::check_tensor_with_dim(const Tensor& tensor, base::DeviceType device_type,
                                          base::DataType data_type, ...) const {
  base::Status st = check_tensor(tensor, device_type, data_type);
  if (!st) return st;
  std::va_list extents;
  va_start(extents, data_type);
  for (int32_t d = 0; d < tensor.dims_size() && st; ++d)
    if (va_arg(extents, int32_t) != tensor.get_dim(d))
      st = base::error::InvalidArgument("The tensor has a wrong dim in dim" + std::to_string(d));
  va_end(extents);
  return st;
}

// ---- LayerParam ----------------------------------------------------------------------------------------
LayerParam::LayerParam(base::DeviceType device_type, LayerType layer_type, bool is_quant_layer,
                       std::string layer_name)
    : Layer(device_type, layer_type, std::move(layer_name)), is_quant_layer_(is_quant_layer) {}

tensor::Tensor& LayerParam::get_weight(int32_t idx) { return slot(weight_slots_, idx); }
const tensor::Tensor& LayerParam::get_weight(int32_t idx) const { return slot(weight_slots_, idx); }

void LayerParam::to_cuda() {
  Layer::to_cuda();
  for (Tensor& w : weight_slots_) w.to_cuda(stream_of(cuda_config_));
  if (!scales_.is_empty()) scales_.to_cuda(stream_of(cuda_config_));
}

base::Status LayerParam::set_weight(int32_t idx, const Tensor& weight) {
  CHECK(weight.data_type() == base::DataType::kDataTypeFp32);
  if (!weight.is_empty()) CHECK(weight.device_type() == device_type_);
  slot(weight_slots_, idx) = weight;
  return base::error::Success();
}

base::Status LayerParam::set_weight(int32_t idx, const std::vector<int32_t>& dims, const void* weight_ptr,
                                    base::DeviceType device_type) {
  CHECK_NE(weight_ptr, nullptr);
  Tensor& dst = slot(weight_slots_, idx);
  size_t numel = 1;
  for (int32_t d : dims) numel *= static_cast<size_t>(d);
  const base::DataType dt = is_quant_layer_ ? base::DataType::kDataTypeInt8 : base::DataType::kDataTypeFp32;
  // a view: the checkpoint mapping (or whoever owns weight_ptr) outlives the layer
  auto view = std::make_shared<base::Buffer>(numel * base::DataTypeSize(dt), nullptr,
                                             const_cast<void*>(weight_ptr), true);
  if (device_type != base::DeviceType::kDeviceUnknown) view->set_device_type(device_type);
  Tensor weight(dt, dims);
  CHECK(weight.assign(view));
  dst = weight;

  if (is_quant_layer_) {
    CHECK(group_size_ > 0 && numel % static_cast<size_t>(group_size_) == 0);
    const auto* behind = static_cast<const int8_t*>(weight_ptr) + numel;
    scales_ = Tensor(base::DataType::kDataTypeFp32, static_cast<int32_t>(numel / static_cast<size_t>(group_size_)),
                     false, nullptr, const_cast<int8_t*>(behind));
    scales_.set_device_type(device_type);
  }
  return base::error::Success();
}

void LayerParam::set_scales(const Tensor& scales) {
  CHECK(!scales.is_empty());
  scales_ = scales;
}
int32_t LayerParam::get_scale_num() const {
  CHECK(!scales_.is_empty());
  return static_cast<int32_t>(scales_.size());
}
}  // namespace op
