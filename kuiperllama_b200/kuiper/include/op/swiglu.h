#ifndef KLLM_KUIPER_OP_SWIGLU_H_
#define KLLM_KUIPER_OP_SWIGLU_H_
#include "layer.h"
namespace op {
// out = silu(in0) * in1 over hidden_dim elements; reference op/swiglu.h.
class SwiGLULayer : public op::Layer {
 public:
  explicit SwiGLULayer(base::DeviceType device_type, int32_t hidden_dim);
  base::Status check() const override;
  base::Status forward() override;

 private:
  int32_t hidden_dim_ = 0;
};
}  // namespace op
#endif
