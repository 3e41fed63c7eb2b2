// model::LLama2Model (reference kuiper/include/model/llama3.h:11-76): Llama-2/3 style decoder.
//
// Same public API; the per-token path is different.  predict() on an embedding row produced by
// embedding()/fill_input() -- the only way demo/main.cpp drives the model -- runs the fused,
// device-resident decoder of libkllm_b200 (one persistent sm_100a kernel per token, weights
// streamed by TMA).  forward() keeps the reference's layer-by-layer orchestration over the
// op registry (same arithmetic, more launches) for callers that hand in their own activations.
#ifndef KLLM_KUIPER_MODEL_LLAMA3_H_
#define KLLM_KUIPER_MODEL_LLAMA3_H_
#include <base/cuda_config.h>

#include "model.h"
#include "op/add.h"
#include "op/embedding.h"
#include "op/rope.h"
#include "op/swiglu.h"

struct kllm_decoder;  // include/kllm_b200.h

namespace model {

struct LLama2Layers {
  std::shared_ptr<op::Layer> add_layer_;
  std::shared_ptr<op::Layer> rope_layer_;
  std::shared_ptr<op::Layer> swiglu_layer_;
  std::shared_ptr<op::Layer> mha_layer_;

  std::vector<std::shared_ptr<op::Layer>> wq_layers_;
  std::vector<std::shared_ptr<op::Layer>> wk_layers_;
  std::vector<std::shared_ptr<op::Layer>> wv_layers_;
  std::vector<std::shared_ptr<op::Layer>> wo_layers_;

  std::vector<std::shared_ptr<op::Layer>> w1_layers_;
  std::vector<std::shared_ptr<op::Layer>> w2_layers_;
  std::vector<std::shared_ptr<op::Layer>> rmsnorm_layers_;  // [0,L) attention, [L,2L) ffn, [2L] final
  std::vector<std::shared_ptr<op::Layer>> w3_layers_;
  std::shared_ptr<op::Layer> cls_layer_;

  std::shared_ptr<op::Layer> embedding_layer_;

  void to_cuda(std::shared_ptr<kernel::CudaConfig> config);
};

class LLama2Model : public Model {
 public:
  explicit LLama2Model(base::TokenizerType tokenizer_type, std::string token_path, std::string model_path,
                       bool is_quant_model);
  ~LLama2Model() override;

  base::Status init(base::DeviceType device_type) override;
  base::Status predict(const tensor::Tensor& input, const tensor::Tensor& pos_tensor, bool is_prompt,
                       int& next) const override;
  base::Status forward(const tensor::Tensor& input, const tensor::Tensor& pos_tensor, int& next) const override;
  op::EmbeddingOutput embedding(const std::vector<int>& tokens) const override;
  tensor::Tensor& get_buffer(ModelBufferType buffer_idx) override;
  const tensor::Tensor& get_buffer(ModelBufferType buffer_idx) const override;

  // "persistent" / "graph" (the fused decoder's engine) -- diagnostic
  const char* decoder_engine() const;

 protected:
  // qkv_bias: checkpoint carries a bias after each of wq/wk/wv per layer (Qwen2 layout)
  LLama2Model(base::TokenizerType tokenizer_type, std::string token_path, std::string model_path,
              bool is_quant_model, bool qkv_bias);

 private:
  void init_mem() override;
  base::Status create_layers() override;
  void create_param_layers() override;
  void create_nonparam_layers() override;
  void create_param_quant_layers() override;
  base::Status create_decoder();
  void ensure_lazy_buffer(ModelBufferType buffer_idx) const;

  void attention_mha(int32_t layer_idx, const tensor::Tensor& pos_tensor) const;
  void attention_rms(int32_t layer_idx, const tensor::Tensor& input) const;
  void feed_forward(int32_t layer_idx, const tensor::Tensor& input) const;
  void attention_qkv(int32_t layer_idx, const tensor::Tensor& pos_tensor) const;
  void cls_logits(const tensor::Tensor& input) const;
  int32_t post_processing(const tensor::Tensor& pos, bool is_prompt) const override;

 private:
  bool qkv_bias_ = false;
  std::shared_ptr<kernel::CudaConfig> cuda_config_;
  std::unique_ptr<LLama2Layers> llama_layers_;
  kllm_decoder* decoder_ = nullptr;
  // tokens of the most recent embedding() call, to map an input row back to its token id
  mutable std::vector<int32_t> last_tokens_;
  mutable const float* last_embeddings_ = nullptr;
  mutable bool logits_in_decoder_ = false;
};
}  // namespace model
#endif
