// The kernel registry -- the primary plug point of the KuiperLLama API.  Typedefs and getters are
// kept verbatim-compatible with kuiper/source/op/kernels/kernels_interface.h:6-68 (same names,
// argument order and meaning; outputs are passed `const` and mutated, as there).  For
// kDeviceCUDA every getter returns an adapter (kernels_interfaces.cpp) that unwraps the tensors
// and calls the sm_100a C-ABI of include/kllm_b200.h.  There is NO CPU backend in this library:
// asking for kDeviceCPU is a fatal error that says so (the CPU restatement lives under oracle/
// and is test infrastructure only).
#ifndef KLLM_KUIPER_KERNELS_INTERFACE_H_
#define KLLM_KUIPER_KERNELS_INTERFACE_H_
#include <base/cuda_config.h>

#include "tensor/tensor.h"
namespace kernel {
typedef void (*AddKernel)(const tensor::Tensor& input1, const tensor::Tensor& input2,
                          const tensor::Tensor& output, void* stream);

typedef void (*MatmulKernel)(const tensor::Tensor& input, const tensor::Tensor& weight,
                             const tensor::Tensor& output, float scale, const CudaConfig* config);

typedef void (*MatmulKernelQuant)(const tensor::Tensor& input, const tensor::Tensor& weight,
                                  const tensor::Tensor& output, int32_t group_size,
                                  const tensor::Tensor& scale, const CudaConfig* config);

typedef void (*EmbeddingKernel)(const tensor::Tensor& input, const tensor::Tensor& weight,
                                const tensor::Tensor& output, int32_t vocab_size, void* stream);

typedef void (*SwigluKernel)(const tensor::Tensor& input1, const tensor::Tensor& input2,
                             const tensor::Tensor& output, void* stream);

typedef void (*MHAKernel)(int32_t pos, int32_t head_num, int32_t layer_index, int32_t seq_len,
                          int32_t kv_dim, int32_t kv_mul, int32_t head_size,
                          const tensor::Tensor& mha_out, const tensor::Tensor& query_tensor,
                          const tensor::Tensor& score_tensor,
                          const tensor::Tensor& key_cache_tensor,
                          const tensor::Tensor& value_cache_tensor, base::DeviceType device_type,
                          CudaConfig*);

typedef void (*RMSNormKernel)(const tensor::Tensor& input, const tensor::Tensor& weight,
                              const tensor::Tensor& output, void* stream);

typedef void (*RoPEKernel)(int32_t dim, int32_t kv_dim, int32_t head_size,
                           const tensor::Tensor& input_q, const tensor::Tensor& input_k,
                           const tensor::Tensor& input_pos, const tensor::Tensor& sin_cache,
                           const tensor::Tensor& cos_cache, void* stream);

typedef void (*ScaleKernel)(float scale, const tensor::Tensor& input, void* stream);

typedef void (*SoftmaxInplaceKernel)(const tensor::Tensor& input, void* stream);

typedef void (*ScaleSumKernel)(const tensor::Tensor& value, const tensor::Tensor& scale,
                               const tensor::Tensor& output, int t, int size, int stride,
                               void* stream);

AddKernel get_add_kernel(base::DeviceType device_type);
EmbeddingKernel get_emb_kernel(base::DeviceType device_type);
MatmulKernel get_matmul_kernel(base::DeviceType device_type);
MatmulKernelQuant get_matmul_kernel_quant8(base::DeviceType device_type);
MHAKernel get_mha_kernel(base::DeviceType device_type);
RMSNormKernel get_rmsnorm_kernel(base::DeviceType device_type);
RoPEKernel get_rope_kernel(base::DeviceType device_type);
ScaleKernel get_scale_kernel(base::DeviceType device_type);
SoftmaxInplaceKernel get_softmax_kernel(base::DeviceType device_type);
SwigluKernel get_swiglu_kernel(base::DeviceType device_type, void* stream = nullptr);
ScaleSumKernel get_scale_sum_kernel(base::DeviceType device_type);

// Non-registry entry points the model code uses (reference rope_kernel.cu:138-151,
// argmax_kernel.cu:73-87).
void sin_cos_cache_calc_cu(int head_size, int max_seq_len, const tensor::Tensor& sin_cache,
                           const tensor::Tensor& cos_cache, cudaStream_t stream);
size_t argmax_kernel_cu(const float* input_ptr, size_t size, void* stream);

// RoPE pairing / constants this library was built for: KLLM_FLAVOUR_* of kllm_b200.h, chosen by
// the same compile definitions as the reference (LLAMA3_SUPPORT / QWEN2_SUPPORT / none).
int build_flavour();
}  // namespace kernel
#endif  // KLLM_KUIPER_KERNELS_INTERFACE_H_
