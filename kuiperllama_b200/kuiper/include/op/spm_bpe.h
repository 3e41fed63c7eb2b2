// A self-contained reader for SentencePiece BPE models (the `tokenizer.model` of Llama-2 /
// TinyLlama checkpoints) -- what the reference gets from libsentencepiece through
// SentencePieceProcessor::{Load, EncodeAsIds, DecodeIds, bos_id, eos_id, GetPieceSize}
// (kuiper/source/op/encode.cpp:13-60).  Own implementation of the published format and
// algorithm, no third-party code:
//   * the model file is a protobuf `ModelProto` (pieces with score and type, TrainerSpec,
//     NormalizerSpec); the wire format is parsed directly;
//   * normalisation: the "identity" rule Llama tokenizers use (no character map), dummy prefix,
//     whitespace escaping to U+2581, optional whitespace squeezing;
//   * BPE: repeatedly merge the adjacent pair whose concatenation is the best-scoring vocabulary
//     piece (ties: leftmost), then map unknown characters to <0xXX> byte pieces (byte_fallback);
//   * decoding: pieces back to text, byte pieces reassembled, control pieces dropped.
// Unigram models and models that carry a compiled NFKC character map are refused with an error.
#ifndef KLLM_KUIPER_OP_SPM_BPE_H_
#define KLLM_KUIPER_OP_SPM_BPE_H_
#include <cstdint>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace op {
class SpmBpeModel {
 public:
  // Empty string on success, else the reason the file cannot be used.
  std::string load(const std::string& path);
  std::string load_from_bytes(std::string_view blob);

  std::vector<int32_t> encode(std::string_view text) const;
  std::string decode(const std::vector<int32_t>& ids) const;

  int32_t piece_size() const { return static_cast<int32_t>(pieces_.size()); }
  int32_t bos_id() const { return bos_id_; }
  int32_t eos_id() const { return eos_id_; }
  int32_t unk_id() const { return unk_id_; }
  const std::string& id_to_piece(int32_t id) const { return pieces_.at(id).text; }

 private:
  enum PieceType { kNormal = 1, kUnknown = 2, kControl = 3, kUserDefined = 4, kUnused = 5, kByte = 6 };
  struct Piece {
    std::string text;
    float score = 0.f;
    int type = kNormal;
  };
  std::string normalize(std::string_view text) const;

  std::vector<Piece> pieces_;
  std::unordered_map<std::string_view, int32_t> merge_vocab_;  // NORMAL / USER_DEFINED / UNUSED pieces
  int32_t byte_piece_[256];
  int32_t unk_id_ = 0, bos_id_ = 1, eos_id_ = 2;
  bool byte_fallback_ = false;
  bool add_dummy_prefix_ = true, remove_extra_whitespaces_ = true, escape_whitespaces_ = true;
  bool has_user_defined_ = false;
  std::string unk_surface_ = " \xE2\x81\x87 ";
};
}  // namespace op
#endif
