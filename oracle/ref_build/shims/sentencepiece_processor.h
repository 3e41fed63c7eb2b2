// Stub of the sentencepiece C++ API, just wide enough for the reference's
// SpeEncodeLayer (kuiper/source/op/encode.cpp:13-56) to construct.  The GPU-oracle
// build only needs Model::init to succeed on synthetic checkpoints; token text is
// never produced.  Test infrastructure only (oracle/_ref), never linked into the product.
#pragma once
#include <string>
#include <vector>
namespace sentencepiece {
namespace util {
enum class StatusCode : int { kOk = 0, kNotFound = 5 };
class Status {
 public:
  Status() = default;
  explicit Status(StatusCode c) : code_(c) {}
  StatusCode code() const { return code_; }
  bool ok() const { return code_ == StatusCode::kOk; }

 private:
  StatusCode code_ = StatusCode::kOk;
};
}  // namespace util

class SentencePieceProcessor {
 public:
  util::Status Load(const std::string&) { return util::Status(); }
  std::vector<int> EncodeAsIds(const std::string& text) const {
    return std::vector<int>(text.empty() ? 0 : 1, 1);
  }
  std::string DecodeIds(const std::vector<int>& ids) const {
    std::string out;
    for (int id : ids) out += "<" + std::to_string(id) + ">";
    return out;
  }
  int bos_id() const { return 1; }
  int eos_id() const { return -12345; }  // never matches: decode runs the full length
  int GetPieceSize() const { return 32000; }
};
}  // namespace sentencepiece
