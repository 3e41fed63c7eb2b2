#ifndef KLLM_KUIPER_OP_MATMUL_H_
#define KLLM_KUIPER_OP_MATMUL_H_
#include <base/cuda_config.h>

#include "layer.h"
namespace op {
// out[dim0] = W[dim0, dim1] . in[dim1] (+ bias).  fp32 weights, or int8 weights with fp32 group
// scales when is_quant_layer (reference op/matmul.h, matmul.cpp:19-118).
class MatmulLayer : public LayerParam {
 public:
  explicit MatmulLayer(base::DeviceType device_type, int32_t dim0, int32_t dim1,
                       bool is_quant_layer = false, bool has_bias = false);
  base::Status check() const override;
  base::Status forward() override;
  base::Status set_bias(int32_t idx, int32_t& dims, const void* bias_ptr, base::DeviceType device_type);
  tensor::Tensor& get_bias(int32_t idx);
  const tensor::Tensor& get_bias(int32_t idx) const;
  bool has_bias() const { return has_bias_; }
  void to_cuda() override;

 private:
  int32_t dim0_ = 0;
  int32_t dim1_ = 0;
  bool has_bias_ = false;
  std::vector<tensor::Tensor> bias_;
};
}  // namespace op
#endif
