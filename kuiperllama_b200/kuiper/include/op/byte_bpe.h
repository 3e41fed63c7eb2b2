// Byte-level BPE tokenizers in the Hugging Face `tokenizer.json` format (Llama-3, Qwen2): what the
// reference builds from nlohmann_json + re2 + a vendored cpp-tiktoken (kuiper/source/op/encode.cpp:
// 62-180).  Own implementation of the published algorithm, no third-party code:
//   * a small JSON reader for tokenizer.json (model.vocab, model.merges, added_tokens, the split
//     pattern of the pre-tokenizer);
//   * the GPT-4 / Llama-3 / Qwen2 split pattern as a hand-written scanner over code points
//     (contractions | letters | digits (1 or 1-3) | punctuation | newlines | blanks), with \p{L} /
//     \p{N} from generated Unicode tables;
//   * GPT-2 byte <-> printable-character mapping, merges applied in rank order inside each piece;
//   * added (special) tokens matched in the raw text first.
// It follows the tokenizers these models ship with (validated against the `tokenizers` package in
// tests/test_tokenizer.py), not the reference's " " -> "Ġ" pre-replacement.
#ifndef KLLM_KUIPER_OP_BYTE_BPE_H_
#define KLLM_KUIPER_OP_BYTE_BPE_H_
#include <cstdint>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

namespace op {
class ByteBpeModel {
 public:
  // Empty string on success, else why the file cannot be used.
  std::string load(const std::string& path);
  std::string load_from_json(std::string_view json);

  std::vector<int32_t> encode(std::string_view text) const;
  std::string decode(const std::vector<int32_t>& ids) const;
  int32_t vocab_size() const { return static_cast<int32_t>(id_to_token_.size()); }
  // id of an added / ordinary token given its literal content, -1 if absent
  int32_t token_to_id(const std::string& content) const;

  // exposed for tests: the pre-tokenizer split of `text` (byte offsets of piece starts)
  std::vector<std::string> split(std::string_view text) const;

 private:
  void encode_piece(std::string_view piece, std::vector<int32_t>& out) const;

  std::unordered_map<std::string, int32_t> vocab_;        // byte-level (printable) spelling -> id
  std::unordered_map<std::string, int32_t> merge_rank_;   // "left right" -> rank
  std::vector<std::pair<std::string, int32_t>> added_;    // literal content -> id, longest first
  std::vector<std::string> id_to_token_;                  // printable spelling, or literal for added
  std::vector<bool> id_is_added_;
  int max_digits_ = 1;          // \p{N} (Qwen2) or \p{N}{1,3} (Llama-3)
  bool ignore_merges_ = false;  // whole piece in the vocabulary -> that id (tiktoken semantics)
};
}  // namespace op
#endif
