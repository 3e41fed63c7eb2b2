#ifndef KLLM_KUIPER_SAMPLER_ARGMAX_SAMPLER_H_
#define KLLM_KUIPER_SAMPLER_ARGMAX_SAMPLER_H_
#include <base/base.h>

#include "sampler.h"
namespace sampler {
// Greedy: index of the maximum logit, lowest index on ties (reference argmax_sampler.cpp:5-13).
class ArgmaxSampler : public Sampler {
 public:
  explicit ArgmaxSampler(base::DeviceType device_type) : Sampler(device_type) {}
  size_t sample(const float* logits, size_t size, void* stream) override;
};
}  // namespace sampler
#endif
