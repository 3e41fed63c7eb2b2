#ifndef KLLM_KUIPER_OP_ADD_H_
#define KLLM_KUIPER_OP_ADD_H_
#include "base/base.h"
#include "layer.h"
namespace op {
// out = in0 + in1, same shapes (residual adds, Qwen2 bias); reference op/add.h.
class VecAddLayer : public Layer {
 public:
  explicit VecAddLayer(base::DeviceType device_type);
  base::Status check() const override;
  base::Status forward() override;
};
}  // namespace op
#endif
