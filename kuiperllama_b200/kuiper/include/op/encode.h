// Tokenizer "layers" (reference op/encode.h).  Tokenisation is not on the decode hot path and its
// third-party stack (sentencepiece, re2, abseil, nlohmann_json) is not vendored.
//   SpeEncodeLayer: SentencePiece BPE models (Llama-2 / TinyLlama `tokenizer.model`) are read by
//     the library's own implementation (op/spm_bpe.h), or by libsentencepiece when built with
//     -DKLLM_WITH_SENTENCEPIECE.  The path "<none>" (or an empty one) selects a deterministic
//     id-level stand-in for synthetic checkpoints (ids in, "<id>" text out); any other path that
//     cannot be loaded is fatal, as in the reference (encode.cpp:24-34).
//   BpeEncodeLayer / QwenEncodeLayer: byte-level BPE `tokenizer.json` files (Llama-3, Qwen2) are
//     read by the library's own implementation (op/byte_bpe.h); "<none>" selects the stand-in.
#ifndef KLLM_KUIPER_OP_ENCODE_H_
#define KLLM_KUIPER_OP_ENCODE_H_
#include <memory>
#include <string>
#include <vector>

#include "layer.h"
#ifdef KLLM_WITH_SENTENCEPIECE
#include <sentencepiece_processor.h>
#else
#include "spm_bpe.h"
#endif
#include "byte_bpe.h"
namespace op {
class EncodeLayerBase : public Layer {
 public:
  explicit EncodeLayerBase(std::string token_model_path, bool has_bos, bool has_eos)
      : Layer(base::DeviceType::kDeviceCPU, LayerType::kLayerEncode, "Encode"),
        has_bos_(has_bos),
        has_eos_(has_eos),
        token_model_path_(std::move(token_model_path)) {}
  virtual std::vector<int32_t> encode(const std::string& sentence) const = 0;
  virtual std::string decode(int32_t token_id) const = 0;
  virtual std::string decode(const std::vector<int32_t>& token_ids) const = 0;
  virtual bool is_sentence_ending(int32_t token_id) const = 0;
  virtual int32_t vocab_size() const = 0;

 protected:
  bool has_bos_ = true;
  bool has_eos_ = false;
  std::string token_model_path_;
};

class SpeEncodeLayer : public EncodeLayerBase {
 public:
  explicit SpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos);
  std::vector<int32_t> encode(const std::string& sentence) const override;
  std::string decode(int32_t token_id) const override;
  std::string decode(const std::vector<int32_t>& token_ids) const override;
  bool is_sentence_ending(int32_t token_id) const override;
  int32_t vocab_size() const override;

 private:
#ifdef KLLM_WITH_SENTENCEPIECE
  std::unique_ptr<sentencepiece::SentencePieceProcessor> spe;
#else
  std::unique_ptr<SpmBpeModel> spm_;  // null: the id-level stand-in
#endif
  int32_t stub_vocab_ = 32000;
};

// Byte-level BPE front ends (Llama-3 / Qwen2 tokenizer.json; reference encode.cpp:62-180).
class BpeEncodeLayer : public EncodeLayerBase {
 public:
  // Llama-3 special tokens: <|begin_of_text|>, <|end_of_text|>, <|eot_id|>
  explicit BpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos);
  std::vector<int32_t> encode(const std::string& sentence) const override;
  std::string decode(int32_t token_id) const override;
  std::string decode(const std::vector<int32_t>& token_ids) const override;
  bool is_sentence_ending(int32_t token_id) const override;
  int32_t vocab_size() const override;

 protected:
  // shared by the two families: load the file, look the three special tokens up
  BpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos, const char* bos, const char* eos,
                 const char* stop2, int32_t stub_vocab);
  std::unique_ptr<ByteBpeModel> bpe_;  // null: the id-level stand-in
  int32_t bos_id_ = -1;
  int32_t eos_id_ = -1;
  int32_t stop_token1_ = -1, stop_token2_ = -1;
  int32_t num_token_ = 0;
};

// Qwen2 special tokens: <|im_start|>, <|im_end|>, <|endoftext|>
class QwenEncodeLayer : public BpeEncodeLayer {
 public:
  explicit QwenEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos);
};
}  // namespace op
#endif
