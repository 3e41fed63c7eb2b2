// The concrete operators of the decode path, one class per entry of the kernel registry
// (source/op/kernels/kernels_interface.h).  Class names, constructor arguments and slot meaning
// follow the reference's op/*.h so model code written against it keeps working; the per-operator
// headers (op/add.h, op/matmul.h, ...) forward here.
//
//                      inputs (slot order)                          output        weights
//   VecAddLayer        a, b                                         a + b         -
//   SwiGLULayer        gate = W1 h, up = W3 h                       silu(gate)*up -
//   RmsNormLayer       x                                            norm(x) * w   w[dim]
//   RoPELayer          q, k, pos (CPU int32), sin table, cos table  (q, k rotated in place)
//   MultiHeadAttention q, score workspace, key cache, value cache   attention     -
//   MatmulLayer        x[dim1]                                      W x (+ bias)  W[dim0, dim1] fp32 | int8+scales
//   EmbeddingLayer     token ids (CPU int32), n as a tensor size    rows of E     E[vocab, dim]
//
// Every forward() = check() then ONE registry kernel on the layer's stream; nothing synchronises.
#ifndef KLLM_KUIPER_OP_DECODER_LAYERS_H_
#define KLLM_KUIPER_OP_DECODER_LAYERS_H_
#include <utility>
#include <vector>

#include "layer.h"

namespace op {
// Adds hide the base class's slot-binding forward(in..., out) overloads; bring them back.
#define KLLM_LAYER_COMMON          \
  using Layer::forward;            \
  base::Status check() const override; \
  base::Status forward() override;

class VecAddLayer : public Layer {
 public:
  explicit VecAddLayer(base::DeviceType device_type);
  KLLM_LAYER_COMMON
};

class SwiGLULayer : public Layer {
 public:
  SwiGLULayer(base::DeviceType device_type, int32_t hidden_dim);
  KLLM_LAYER_COMMON
 private:
  int32_t hidden_dim_ = 0;
};

class RmsNormLayer : public LayerParam {
 public:
  RmsNormLayer(base::DeviceType device_type, int32_t dim);
  KLLM_LAYER_COMMON
 private:
  int32_t dim_ = 0;
};

class RoPELayer : public Layer {
 public:
  RoPELayer(base::DeviceType device_type, int32_t dim, int32_t kv_dim, int32_t head_size);
  KLLM_LAYER_COMMON
 private:
  int32_t dim_ = 0, kv_dim_ = 0, head_size_ = 0;
};

// Single-position attention over the KV cache of one layer: scores, softmax, weighted values for
// every query head (grouped-query attention: kv_mul query heads share one kv head).
class MultiHeadAttention : public Layer {
 public:
  MultiHeadAttention(base::DeviceType device_type, int32_t layer_index, int32_t kv_mul, int32_t kv_dim,
                     int32_t seq_len, int32_t head_num, int32_t head_size);
  KLLM_LAYER_COMMON
  void set_pos(int32_t pos) { pos_ = pos; }
  void set_layer_idx(int32_t layer_idx) { layer_index_ = layer_idx; }
 private:
  int32_t layer_index_ = 0, pos_ = 0;
  int32_t kv_mul_ = 0, kv_dim_ = 0, seq_len_ = 0, head_num_ = 0, head_size_ = 0;
};

// out[dim0] = W[dim0, dim1] . in[dim1] (+ bias[dim0] for Qwen2's q / k / v).
class MatmulLayer : public LayerParam {
 public:
  MatmulLayer(base::DeviceType device_type, int32_t dim0, int32_t dim1, bool is_quant_layer = false,
              bool has_bias = false);
  KLLM_LAYER_COMMON
  bool has_bias() const { return has_bias_; }
  // wraps `dims` floats at bias_ptr (a view, like set_weight)
  base::Status set_bias(int32_t idx, int32_t& dims, const void* bias_ptr, base::DeviceType device_type);
  tensor::Tensor& get_bias(int32_t idx);
  const tensor::Tensor& get_bias(int32_t idx) const;
  void to_cuda() override;
 private:
  int32_t dim0_ = 0, dim1_ = 0;
  bool has_bias_ = false;
  std::vector<tensor::Tensor> bias_;
};

// What Model::embedding() hands to fill_input(): the ids, their embedding rows, and the count.
struct EmbeddingOutput {
  tensor::Tensor input_tokens;      // CPU int32 [n]
  tensor::Tensor input_embeddings;  // [n, dim] on the model's device
  tensor::Tensor input_token_num;   // only its size() == n matters
  EmbeddingOutput(tensor::Tensor tokens, tensor::Tensor embeddings, tensor::Tensor token_num)
      : input_tokens(std::move(tokens)),
        input_embeddings(std::move(embeddings)),
        input_token_num(std::move(token_num)) {}
};

class EmbeddingLayer : public LayerParam {
 public:
  EmbeddingLayer(base::DeviceType device_type, int32_t dim, int32_t seq_len, int32_t vocab_size);
  KLLM_LAYER_COMMON
 private:
  int32_t dim_ = 0, seq_len_ = 0, vocab_size_ = 0;
};
#undef KLLM_LAYER_COMMON
}  // namespace op
#endif  // KLLM_KUIPER_OP_DECODER_LAYERS_H_
