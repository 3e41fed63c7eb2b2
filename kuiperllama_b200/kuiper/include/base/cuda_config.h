// kernel::CudaConfig: the one CUDA stream all layers of a model enqueue on.  The config owns the
// stream (created by the model in init(), destroyed with the last reference to the config).
#ifndef KLLM_KUIPER_BASE_CUDA_CONFIG_H_
#define KLLM_KUIPER_BASE_CUDA_CONFIG_H_
#include <cuda_runtime_api.h>

namespace kernel {
struct CudaConfig {
  CudaConfig() = default;
  CudaConfig(const CudaConfig&) = delete;
  CudaConfig& operator=(const CudaConfig&) = delete;
  ~CudaConfig() {
    if (stream != nullptr) cudaStreamDestroy(stream);
  }
  cudaStream_t stream = nullptr;
};
}  // namespace kernel
#endif  // KLLM_KUIPER_BASE_CUDA_CONFIG_H_
