// A KuiperLLama checkpoint on disk and in memory (what the reference declares in model/config.h and
// model/raw_model_data.h; both headers forward here).
//
//   file  = header | payload
//   header: 7 x int32 {dim, hidden_dim, layer_num, head_num, kv_head_num, vocab_size, seq_len}
//           (tools/export.py); vocab_size > 0 says "the classifier shares the embedding table",
//           < 0 "a separate classifier follows"; int8 files (export.py --version 3) add an eighth
//           int32, the quantisation group size
//   payload: fp32 tensors, or int8 blocks each followed by its fp32 group scales
//
// RawModelData is the mmap of that file; weight(offset) addresses the payload in elements of the
// file's weight type (floats for fp32 files, bytes for int8 files).
#ifndef KLLM_KUIPER_MODEL_CHECKPOINT_FILE_H_
#define KLLM_KUIPER_MODEL_CHECKPOINT_FILE_H_
#include <cstddef>
#include <cstdint>
#include <ostream>

namespace model {
struct ModelConfig {  // the header, field for field
  int32_t dim = 0, hidden_dim = 0, layer_num = 0, head_num = 0, kv_head_num = 0, vocab_size = 0, seq_len = 0;
};
static_assert(sizeof(ModelConfig) == 28, "checkpoint header is 7 x int32");

// header + what follows from it
struct TransformerConfig {
  int32_t dim_ = 0, hidden_dim_ = 0, layer_num_ = 0, head_num_ = 0, kv_head_num_ = 0, seq_len_ = 0;
  int32_t vocab_size_ = 0;         // |header vocab_size|
  bool is_shared_weight_ = false;  // header vocab_size > 0
  int32_t head_size_ = 0;          // dim / head_num
  int32_t kv_dim_ = 0;             // dim * kv_head_num / head_num
  int32_t kv_mul_ = 0;             // head_num / kv_head_num: query heads per kv head
};
inline std::ostream& operator<<(std::ostream& os, const TransformerConfig& c) {
  return os << "\nkv_dim: " << c.kv_dim_ << "\nkv_mul_: " << c.kv_mul_ << "\nhead_size: " << c.head_size_
            << "\nvocab_size_: " << c.vocab_size_ << "\ndim: " << c.dim_ << "\nhidden_dim_: " << c.hidden_dim_
            << "\nlayer_num: " << c.layer_num_ << "\nhead_num_: " << c.head_num_ << "\nkv_head_num: " << c.kv_head_num_
            << "\nseq_len_: " << c.seq_len_ << "\nis_shared_weight: " << c.is_shared_weight_;
}

struct RawModelData {
  RawModelData() = default;
  RawModelData(const RawModelData&) = delete;
  RawModelData& operator=(const RawModelData&) = delete;
  virtual ~RawModelData();  // unmaps and closes
  virtual const void* weight(size_t offset) const = 0;

  int32_t fd = -1;
  size_t file_size = 0;
  void* data = nullptr;         // the whole mapping
  void* weight_data = nullptr;  // first payload byte
};
struct RawModelDataFp32 final : RawModelData {
  const void* weight(size_t offset) const override { return static_cast<const float*>(weight_data) + offset; }
};
struct RawModelDataInt8 final : RawModelData {
  const void* weight(size_t offset) const override { return static_cast<const int8_t*>(weight_data) + offset; }
};
}  // namespace model
#endif  // KLLM_KUIPER_MODEL_CHECKPOINT_FILE_H_
