#ifndef KLLM_KUIPER_OP_EMBEDDING_H_
#define KLLM_KUIPER_OP_EMBEDDING_H_
#include <utility>

#include "layer.h"
namespace op {
struct EmbeddingOutput {
  tensor::Tensor input_tokens;      // CPU int32 [n]
  tensor::Tensor input_embeddings;  // [n, dim] on the model's device
  tensor::Tensor input_token_num;   // only its size() = n is used
  explicit EmbeddingOutput(tensor::Tensor input_tokens, tensor::Tensor input_embeddings,
                           tensor::Tensor input_token_num)
      : input_tokens(std::move(input_tokens)),
        input_embeddings(std::move(input_embeddings)),
        input_token_num(std::move(input_token_num)) {}
};

// Row gather from the [vocab, dim] table.  Inputs: token ids (CPU int32), a tensor whose size()
// is the token count; output [n, dim] (reference embedding.cpp:18-25, llama3.cpp:589-594).
class EmbeddingLayer : public LayerParam {
 public:
  explicit EmbeddingLayer(base::DeviceType device_type, int32_t dim, int32_t seq_len, int32_t vocab_size);
  base::Status check() const override;
  base::Status forward() override;

 private:
  int32_t dim_ = 0;
  int32_t seq_len_ = 0;
  int32_t vocab_size_ = 0;
};
}  // namespace op
#endif
