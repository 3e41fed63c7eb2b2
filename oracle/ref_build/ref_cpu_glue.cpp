// Link-time stand-ins for the seven reference CPU kernels whose bodies are Armadillo
// expressions (kuiper/source/op/kernels/cpu/{matmul,rmsnorm,softmax,swiglu,add,scale,
// scale_sum}_kernel.cpp).  Armadillo/OpenBLAS are not installed here, so those translation
// units cannot be compiled; the registry (kernels_interfaces.cpp) still references their
// symbols.  Each stand-in keeps the reference signature and forwards to the C restatement in
// oracle/kuiper_oracle.c.  This lets the reference's OWN orchestration (llama3.cpp, cpu/
// mha_kernel.cpp, cpu/rope_kernel.cpp, cpu/emb_kernel.cpp -- all compiled unmodified) run on
// kDeviceCPU as a cross-check of ko_model_step.  TEST INFRASTRUCTURE ONLY.
#include "../source/op/kernels/cpu/add_kernel.h"
#include "../source/op/kernels/cpu/matmul_kernel.h"
#include "../source/op/kernels/cpu/rmsnorm_kernel.h"
#include "../source/op/kernels/cpu/scale_kernel.h"
#include "../source/op/kernels/cpu/scale_sum_kernel.h"
#include "../source/op/kernels/cpu/softmax_kernel.h"
#include "../source/op/kernels/cpu/swiglu_kernel.h"
#include "../source/op/kernels/kernels_interface.h"
#include "kuiper_oracle.h"

namespace kernel {
void add_kernel_cpu(const tensor::Tensor& a, const tensor::Tensor& b, const tensor::Tensor& out,
                    void*) {
  ko_add(a.ptr<float>(), b.ptr<float>(), const_cast<float*>(out.ptr<float>()),
         static_cast<int>(a.size()));
}

void matmul_kernel_cpu(const tensor::Tensor& input, const tensor::Tensor& weight,
                       const tensor::Tensor& output, float scale, const CudaConfig*) {
  ko_matmul_f32(input.ptr<float>(), weight.ptr<float>(), const_cast<float*>(output.ptr<float>()),
                weight.get_dim(1), weight.get_dim(0), scale);
}

void rmsnorm_kernel_cpu(const tensor::Tensor& input, const tensor::Tensor& weight,
                        const tensor::Tensor& output, void*) {
#ifdef QWEN2_SUPPORT
  const float eps = 1e-6f;
#else
  const float eps = 1e-5f;
#endif
  ko_rmsnorm(input.ptr<float>(), weight.ptr<float>(), const_cast<float*>(output.ptr<float>()),
             static_cast<int>(input.size()), eps);
}

void scale_inplace_cpu(float scale, const tensor::Tensor& t, void*) {
  float* p = const_cast<float*>(t.ptr<float>());
  for (size_t i = 0; i < t.size(); ++i) p[i] = p[i] * scale;
}

void scale_sum_kernel_cpu(const tensor::Tensor& value, const tensor::Tensor& scale,
                          const tensor::Tensor& output, int pos, int size, int stride, void*) {
  ko_scale_sum(value.ptr<float>(), scale.ptr<float>(), const_cast<float*>(output.ptr<float>()),
               pos, size, stride);
}

void softmax_inplace_cpu(const tensor::Tensor& input, void*) {
  ko_softmax_inplace(const_cast<float*>(input.ptr<float>()), static_cast<int>(input.size()));
}

void softmax_inplace_cpu(const float* input_ptr, size_t size) {
  ko_softmax_inplace(const_cast<float*>(input_ptr), static_cast<int>(size));
}

void swiglu_kernel_cpu(const tensor::Tensor& in1, const tensor::Tensor& in2,
                       const tensor::Tensor& out, void*) {
  // cpu/swiglu_kernel.cpp:21 also overwrites input1 with silu(input1).
  float* x1 = const_cast<float*>(in1.ptr<float>());
  const int n = static_cast<int>(in1.size());
  ko_swiglu(x1, in2.ptr<float>(), const_cast<float*>(out.ptr<float>()), n);
}
}  // namespace kernel

// CPU whole-model runner over the reference's own llama3.cpp orchestration.
#include "model/llama3.h"
extern "C" {
void* kref_cpu_model_create(const char* checkpoint_path) {
  auto* m = new model::LLama2Model(base::TokenizerType::kEncodeSpe, "stub-tokenizer",
                                   checkpoint_path, false);
  auto st = m->init(base::DeviceType::kDeviceCPU);
  if (!st) {
    fprintf(stderr, "kref_cpu_model_create: %s\n", st.get_err_msg().c_str());
    delete m;
    return nullptr;
  }
  return m;
}
void kref_cpu_model_destroy(void* h) { delete static_cast<model::LLama2Model*>(h); }
int kref_cpu_model_step(void* h, int token, int pos, float* logits_out, int vocab) {
  const model::LLama2Model& m = *static_cast<model::LLama2Model*>(h);
  tensor::Tensor pos_tensor = m.get_buffer(model::ModelBufferType::kInputPos);
  pos_tensor.index<int32_t>(0) = pos;
  std::vector<int32_t> tokens{token};
  const auto& emb = m.embedding(tokens);
  tensor::Tensor input = m.fill_input(pos_tensor, emb, false);
  int next = -1;
  m.predict(input, pos_tensor, false, next);
  if (logits_out) {
    const tensor::Tensor& logits = m.get_buffer(model::ModelBufferType::kForwardOutput);
    memcpy(logits_out, logits.ptr<float>(), sizeof(float) * vocab);
  }
  return next;
}
}
