// model::Qwen2Model (reference kuiper/include/model/qwen2.h): the Llama decoder plus q/k/v
// biases, read from the export_qwen2.py layout.  Build the library with -DQWEN2_SUPPORT=ON for
// the Qwen2 arithmetic flavour (half-split RoPE, theta 1e6, eps 1e-6), exactly as the reference.
#ifndef KLLM_KUIPER_MODEL_QWEN2_H_
#define KLLM_KUIPER_MODEL_QWEN2_H_
#include "llama3.h"
namespace model {
using Qwen2Layers = LLama2Layers;

class Qwen2Model : public LLama2Model {
 public:
  explicit Qwen2Model(base::TokenizerType tokenizer_type, std::string token_path, std::string model_path,
                      bool is_quant_model)
      : LLama2Model(tokenizer_type, std::move(token_path), std::move(model_path), is_quant_model,
                    /*qkv_bias=*/!is_quant_model) {}
};
}  // namespace model
#endif
