// kernel::CudaConfig -- the stream every layer of a model shares
// (reference kuiper/include/base/cuda_config.h:6-13: the config OWNS its stream).
#ifndef KLLM_KUIPER_BASE_CUDA_CONFIG_H_
#define KLLM_KUIPER_BASE_CUDA_CONFIG_H_
#include <cuda_runtime_api.h>
namespace kernel {
struct CudaConfig {
  cudaStream_t stream = nullptr;
  ~CudaConfig() {
    if (stream != nullptr) cudaStreamDestroy(stream);
  }
};
}  // namespace kernel
#endif  // KLLM_KUIPER_BASE_CUDA_CONFIG_H_
