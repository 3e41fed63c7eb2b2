#ifndef KLLM_KUIPER_OP_RMSNORM_H_
#define KLLM_KUIPER_OP_RMSNORM_H_
#include "layer.h"
namespace op {
// out = w * x * rsqrt(mean(x^2) + eps); one weight [dim]; reference op/rmsnorm.h.
class RmsNormLayer : public LayerParam {
 public:
  explicit RmsNormLayer(base::DeviceType device_type, int32_t dim);
  base::Status check() const override;
  base::Status forward() override;

 private:
  int32_t dim_ = 0;
};
}  // namespace op
#endif
