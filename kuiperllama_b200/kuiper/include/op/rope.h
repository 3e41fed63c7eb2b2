#ifndef KLLM_KUIPER_OP_ROPE_H_
#define KLLM_KUIPER_OP_ROPE_H_
#include "layer.h"
namespace op {
// Rotates q [dim] and k [kv_dim] in place.  Inputs: q, k, pos (CPU int32 [1]), sin table, cos
// table; the output slot is unused (the reference calls it with an empty tensor, rope.cpp:11-12).
class RoPELayer : public Layer {
 public:
  explicit RoPELayer(base::DeviceType device_type, int32_t dim, int32_t kv_dim, int32_t head_size);
  base::Status check() const override;
  base::Status forward() override;

 private:
  int32_t dim_ = 0;
  int32_t kv_dim_ = 0;
  int32_t head_size_ = 0;
};
}  // namespace op
#endif
