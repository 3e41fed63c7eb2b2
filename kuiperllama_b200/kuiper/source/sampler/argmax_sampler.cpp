#include "sampler/argmax_sampler.h"

#include "../op/kernels/kernels_interface.h"
namespace sampler {
size_t ArgmaxSampler::sample(const float* logits, size_t size, void* stream) {
  CHECK(device_type_ == base::DeviceType::kDeviceCUDA)
      << "ArgmaxSampler: this library has no CPU backend";
  return kernel::argmax_kernel_cu(logits, size, stream);
}
}  // namespace sampler
