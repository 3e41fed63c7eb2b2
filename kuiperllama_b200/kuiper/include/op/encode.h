// Tokenizer front ends ("encode layers") of the kuiper:: API.  The class names the model code and
// the reference's callers use -- EncodeLayerBase, SpeEncodeLayer, BpeEncodeLayer, QwenEncodeLayer --
// are thin shells here: each picks a TokenizerBackend and the special tokens of its family, and
// one shared implementation does the framing (BOS / EOS) and the stop-token test.
//
//   SpeEncodeLayer    SentencePiece BPE models, the `tokenizer.model` of Llama-2 / TinyLlama
//                     (own reader: op/spm_bpe.h; libsentencepiece with -DKLLM_WITH_SENTENCEPIECE)
//   BpeEncodeLayer    byte-level BPE `tokenizer.json` of Llama-3 (own reader: op/byte_bpe.h);
//                     <|begin_of_text|> / <|end_of_text|>, also stops on <|eot_id|>
//   QwenEncodeLayer   the same for Qwen2: <|im_start|> / <|im_end|>, also stops on <|endoftext|>
//
// The path "<none>" (or an empty one) selects a deterministic id-level stand-in for synthetic
// checkpoints: one id per byte in, "<id>" text out, never a sentence end.  Any other path that
// cannot be loaded is fatal, as in the reference.  Tokenisation is host-side, off the hot path.
#ifndef KLLM_KUIPER_OP_ENCODE_H_
#define KLLM_KUIPER_OP_ENCODE_H_
#include <memory>
#include <string>
#include <vector>

#include "layer.h"

namespace op {
// What a tokenizer implementation provides; ids are the model's vocabulary ids.
class TokenizerBackend {
 public:
  virtual ~TokenizerBackend() = default;
  virtual std::vector<int32_t> encode(const std::string& text) const = 0;  // no BOS / EOS
  virtual std::string decode(const std::vector<int32_t>& ids) const = 0;
  virtual int32_t vocab_size() const = 0;
  virtual int32_t id_of(const std::string& special_token) const = 0;  // -1 if absent
};

class EncodeLayerBase : public Layer {
 public:
  // has_bos / has_eos: whether encode() frames the ids with the family's BOS / EOS token
  EncodeLayerBase(std::string token_model_path, bool has_bos, bool has_eos);
  ~EncodeLayerBase() override;

  virtual std::vector<int32_t> encode(const std::string& sentence) const;
  virtual std::string decode(const std::vector<int32_t>& token_ids) const;
  virtual std::string decode(int32_t token_id) const;
  virtual bool is_sentence_ending(int32_t token_id) const;  // generation stops on these ids
  virtual int32_t vocab_size() const;

 protected:
  // called by the family constructors once they know which backend and which specials
  void adopt(std::unique_ptr<TokenizerBackend> backend, int32_t bos_id, int32_t eos_id, int32_t extra_stop_id,
             bool stops_generation);
  static bool is_stand_in_path(const std::string& path) { return path.empty() || path == "<none>"; }

  bool has_bos_ = true, has_eos_ = false;
  std::string token_model_path_;
  std::unique_ptr<TokenizerBackend> backend_;
  int32_t bos_id_ = -1, eos_id_ = -1, stop_token2_ = -1;
  bool stops_generation_ = false;
};

class SpeEncodeLayer : public EncodeLayerBase {
 public:
  SpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos);
};

class BpeEncodeLayer : public EncodeLayerBase {
 public:
  BpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos);  // Llama-3 specials

 protected:
  BpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos, const char* bos, const char* eos,
                 const char* extra_stop, int32_t stand_in_vocab);
};

class QwenEncodeLayer : public BpeEncodeLayer {
 public:
  QwenEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos);
};
}  // namespace op
#endif  // KLLM_KUIPER_OP_ENCODE_H_
