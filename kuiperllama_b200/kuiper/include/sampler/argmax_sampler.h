// Token samplers.  The decode path is greedy: ArgmaxSampler returns the index of the largest logit,
// the lowest index on ties, computed on the device the logits live on (argmax_kernel_cu).
// sampler/sampler.h forwards here.
#ifndef KLLM_KUIPER_SAMPLER_ARGMAX_SAMPLER_H_
#define KLLM_KUIPER_SAMPLER_ARGMAX_SAMPLER_H_
#include <cstddef>

#include "base/base.h"

namespace sampler {
class Sampler {
 public:
  explicit Sampler(base::DeviceType device_type) : device_type_(device_type) {}
  virtual ~Sampler() = default;
  // logits: `size` floats on this sampler's device; blocks until the id is known
  virtual size_t sample(const float* logits, size_t size, void* stream = nullptr) = 0;

 protected:
  base::DeviceType device_type_;
};

class ArgmaxSampler final : public Sampler {
 public:
  using Sampler::Sampler;
  size_t sample(const float* logits, size_t size, void* stream) override;
};
}  // namespace sampler
#endif  // KLLM_KUIPER_SAMPLER_ARGMAX_SAMPLER_H_
