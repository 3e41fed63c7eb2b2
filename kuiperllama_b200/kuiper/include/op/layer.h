// The operator interface of the kuiper:: API surface (what reference kuiper/include/op/layer.h
// exposes as BaseLayer / Layer / LayerParam), redesigned around ONE base class:
//
//   op::Layer       numbered input / output tensor slots, a device + stream, check() and forward().
//                   forward(in..., out) is shorthand for "bind the slots, then forward()".
//   op::LayerParam  adds numbered weight slots.  set_weight(idx, dims, ptr) wraps memory the
//                   caller keeps alive (a view into the mmap'd checkpoint); for an int8 layer the
//                   fp32 group scales are taken from right behind the int8 block, which is how
//                   export.py --version 3 lays a tensor out.  to_cuda() uploads slots and weights.
//
// `op::BaseLayer` remains as an alias so code written against the three-level hierarchy compiles.
#ifndef KLLM_KUIPER_OP_LAYER_H_
#define KLLM_KUIPER_OP_LAYER_H_
#include <base/cuda_config.h>

#include <initializer_list>
#include <memory>
#include <string>
#include <vector>

#include "base/base.h"
#include "tensor/tensor.h"

namespace op {
enum class LayerType : uint8_t {
  kLayerUnknown = 0, kLayerLinear = 1, kLayerEncode = 2, kLayerEmbedding = 3, kLayerRMSNorm = 4,
  kLayerMatmul = 5, kLayerRoPe = 6, kLayerMHA = 7, kLayerSoftmax = 8, kLayerAdd = 9, kLayerSwiGLU = 10,
};

class Layer {
 public:
  using Tensor = tensor::Tensor;

  Layer(base::DeviceType device_type, LayerType layer_type, std::string layer_name = "");
  virtual ~Layer() = default;

  // ---- what the layer is ---------------------------------------------------------------------
  LayerType layer_type() const { return layer_type_; }
  base::DataType data_type() const { return data_type_; }  // activations are fp32 on this path
  base::DeviceType device_type() const { return device_type_; }
  void set_device_type(base::DeviceType device_type) { device_type_ = device_type; }
  const std::string& get_layer_name() const { return layer_name_; }
  void set_layer_name(const std::string& layer_name) { layer_name_ = layer_name; }

  // ---- tensor slots (copies share the caller's buffer) ------------------------------------------
  void reset_input_size(size_t size) { in_slots_.resize(size); }
  void reset_output_size(size_t size) { out_slots_.resize(size); }
  size_t input_size() const { return in_slots_.size(); }
  size_t output_size() const { return out_slots_.size(); }
  virtual void set_input(int32_t idx, const Tensor& input);
  virtual void set_output(int32_t idx, const Tensor& output);
  Tensor& get_input(int32_t idx);
  Tensor& get_output(int32_t idx);
  const Tensor& get_input(int32_t idx) const;
  const Tensor& get_output(int32_t idx) const;

  // ---- running it ------------------------------------------------------------------------------
  virtual base::Status init();
  virtual base::Status check() const;  // shapes / dtypes / devices of everything bound
  virtual base::Status forward();      // on whatever is bound; concrete layers override this one
  base::Status forward(const Tensor& in0, const Tensor& out) { return bind_and_run({&in0}, out); }
  base::Status forward(const Tensor& in0, const Tensor& in1, const Tensor& out) {
    return bind_and_run({&in0, &in1}, out);
  }
  base::Status forward(const Tensor& in0, const Tensor& in1, const Tensor& in2, const Tensor& out) {
    return bind_and_run({&in0, &in1, &in2}, out);
  }
  base::Status forward(const Tensor& in0, const Tensor& in1, const Tensor& in2, const Tensor& in3,
                       const Tensor& out) {
    return bind_and_run({&in0, &in1, &in2, &in3}, out);
  }
  base::Status forward(const Tensor& in0, const Tensor& in1, const Tensor& in2, const Tensor& in3,
                       const Tensor& in4, const Tensor& out) {
    return bind_and_run({&in0, &in1, &in2, &in3, &in4}, out);
  }

  // ---- device ------------------------------------------------------------------------------------
  virtual void to_cuda();  // uploads whatever the slots hold
  void set_cuda_config(std::shared_ptr<kernel::CudaConfig> config);
  std::shared_ptr<kernel::CudaConfig> cuda_config() const { return cuda_config_; }

  // ---- weights: only LayerParam has any ------------------------------------------------------------
  virtual base::Status set_weight(int32_t idx, const Tensor& weight);
  virtual base::Status set_weight(int32_t idx, const std::vector<int32_t>& dims, const void* weight_ptr,
                                  base::DeviceType device_type = base::DeviceType::kDeviceUnknown);

  // ---- helpers for check() implementations -----------------------------------------------------------
  base::Status check_tensor(const Tensor& tensor, base::DeviceType device_type, base::DataType data_type) const;
  // ... followed by one int per dimension: the extent `tensor` must have there
  base::Status check_tensor_with_dim(const Tensor& tensor, base::DeviceType device_type,
                                     base::DataType data_type, ...) const;

 protected:
  base::Status bind_and_run(std::initializer_list<const Tensor*> inputs, const Tensor& output);

  std::string layer_name_;
  LayerType layer_type_ = LayerType::kLayerUnknown;
  base::DataType data_type_ = base::DataType::kDataTypeFp32;
  base::DeviceType device_type_ = base::DeviceType::kDeviceUnknown;
  std::vector<Tensor> in_slots_, out_slots_;
  std::shared_ptr<kernel::CudaConfig> cuda_config_;
};
using BaseLayer = Layer;

class LayerParam : public Layer {
 public:
  LayerParam(base::DeviceType device_type, LayerType layer_type, bool is_quant_layer = false,
             std::string layer_name = "");
  using Layer::forward;

  void reset_weight_size(size_t size) { weight_slots_.resize(size); }
  size_t weight_size() const { return weight_slots_.size(); }
  Tensor& get_weight(int32_t idx);
  const Tensor& get_weight(int32_t idx) const;
  base::Status set_weight(int32_t idx, const Tensor& weight) override;
  base::Status set_weight(int32_t idx, const std::vector<int32_t>& dims, const void* weight_ptr,
                          base::DeviceType device_type = base::DeviceType::kDeviceUnknown) override;
  void to_cuda() override;

  // int8 group quantisation (export.py --version 3): one fp32 scale per `group_size` weights
  bool is_quant_layer() const { return is_quant_layer_; }
  void set_group_size(int32_t group_size) { group_size_ = group_size; }
  int32_t group_size() const { return group_size_; }
  void set_scales(const Tensor& scales);
  const Tensor& get_scales() const { return scales_; }
  int32_t get_scale_num() const;

 protected:
  bool is_quant_layer_ = false;
  int32_t group_size_ = 0;
  Tensor scales_;
  std::vector<Tensor> weight_slots_;
};
}  // namespace op
#endif  // KLLM_KUIPER_OP_LAYER_H_
