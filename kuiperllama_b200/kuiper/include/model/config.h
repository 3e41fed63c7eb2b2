#ifndef KLLM_KUIPER_MODEL_CONFIG_H_
#define KLLM_KUIPER_MODEL_CONFIG_H_
#include <cstdint>
#include <ostream>
namespace model {
// The 7 x int32 header at the start of every checkpoint (reference model/config.h:5-13;
// export.py:91-92).  A negative vocab_size flags a separate classifier matrix.
struct ModelConfig {
  int32_t dim = 0;
  int32_t hidden_dim = 0;
  int32_t layer_num = 0;
  int32_t head_num = 0;
  int32_t kv_head_num = 0;
  int32_t vocab_size = 0;
  int32_t seq_len = 0;
};

// Header + derived quantities (model.cpp:125-151).
struct TransformerConfig {
  int32_t kv_dim_ = 0;
  int32_t kv_mul_ = 0;
  int32_t head_size_ = 0;
  int32_t vocab_size_ = 0;
  int32_t dim_ = 0;
  int32_t hidden_dim_ = 0;
  int32_t layer_num_ = 0;
  int32_t head_num_ = 0;
  int32_t kv_head_num_ = 0;
  int32_t seq_len_ = 0;
  bool is_shared_weight_ = false;

  friend std::ostream& operator<<(std::ostream& os, const TransformerConfig& c) {
    return os << "\nkv_dim: " << c.kv_dim_ << "\nkv_mul_: " << c.kv_mul_ << "\nhead_size: " << c.head_size_
              << "\nvocab_size_: " << c.vocab_size_ << "\ndim: " << c.dim_ << "\nhidden_dim_: " << c.hidden_dim_
              << "\nlayer_num: " << c.layer_num_ << "\nhead_num_: " << c.head_num_
              << "\nkv_head_num: " << c.kv_head_num_ << "\nseq_len_: " << c.seq_len_
              << "\nis_shared_weight: " << c.is_shared_weight_;
  }
};
}  // namespace model
#endif
