// Tokenizer front ends (reference kuiper/source/op/encode.cpp).  See op/encode.h for scope.
#include "op/encode.h"

#include <fstream>
#include <sstream>

namespace op {
namespace {
// "<12><7>" style text for ids: lossless and obviously synthetic
std::string ids_to_text(const std::vector<int32_t>& ids) {
  std::ostringstream os;
  for (int32_t id : ids) os << '<' << id << '>';
  return os.str();
}
// stand-in encoding: BOS (1) then one id per byte, offset past the control ids
std::vector<int32_t> bytes_to_ids(const std::string& s, bool bos, bool eos, int32_t vocab) {
  std::vector<int32_t> ids;
  if (bos) ids.push_back(1);
  for (unsigned char c : s) ids.push_back(3 + static_cast<int32_t>(c) % (vocab > 259 ? 256 : 1));
  if (eos) ids.push_back(2);
  return ids;
}
}  // namespace

SpeEncodeLayer::SpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos)
    : EncodeLayerBase(std::move(token_model_path), has_bos, has_eos) {
#ifdef KLLM_WITH_SENTENCEPIECE
  spe = std::make_unique<sentencepiece::SentencePieceProcessor>();
  auto rc = spe->Load(token_model_path_);
  if (!rc.ok()) {
    LOG(FATAL) << "The token model path is not valid, please check the path and type of token model.";
  }
#else
  if (token_model_path_.empty() || token_model_path_ == "<none>") {
    LOG(INFO) << "no tokenizer model given: using the id-level stand-in tokenizer";
    return;
  }
  spm_ = std::make_unique<SpmBpeModel>();
  const std::string err = spm_->load(token_model_path_);
  if (!err.empty()) {
    LOG(FATAL) << "The token model path is not valid, please check the path and type of token model: "
               << token_model_path_ << ": " << err;
  }
#endif
}

std::vector<int32_t> SpeEncodeLayer::encode(const std::string& sentence) const {
#ifdef KLLM_WITH_SENTENCEPIECE
  std::vector<int32_t> ids = spe->EncodeAsIds(sentence);
  if (has_bos_) ids.insert(ids.begin(), spe->bos_id());
  if (has_eos_) ids.push_back(spe->eos_id());
  return ids;
#else
  if (!spm_) return bytes_to_ids(sentence, has_bos_, has_eos_, stub_vocab_);
  std::vector<int32_t> ids = spm_->encode(sentence);
  if (has_bos_) ids.insert(ids.begin(), spm_->bos_id());
  if (has_eos_) ids.push_back(spm_->eos_id());
  return ids;
#endif
}

std::string SpeEncodeLayer::decode(int32_t token_id) const { return decode(std::vector<int32_t>{token_id}); }

std::string SpeEncodeLayer::decode(const std::vector<int32_t>& token_ids) const {
#ifdef KLLM_WITH_SENTENCEPIECE
  return spe->DecodeIds(token_ids);
#else
  return spm_ ? spm_->decode(token_ids) : ids_to_text(token_ids);
#endif
}

bool SpeEncodeLayer::is_sentence_ending(int32_t token_id) const {
#ifdef KLLM_WITH_SENTENCEPIECE
  return token_id == spe->eos_id();
#else
  // stand-in (synthetic checkpoints): always decode the requested number of steps
  return spm_ ? token_id == spm_->eos_id() : false;
#endif
}

int32_t SpeEncodeLayer::vocab_size() const {
#ifdef KLLM_WITH_SENTENCEPIECE
  return spe->GetPieceSize();
#else
  return spm_ ? spm_->piece_size() : stub_vocab_;
#endif
}

BpeEncodeLayer::BpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos, const char* bos,
                               const char* eos, const char* stop2, int32_t stub_vocab)
    : EncodeLayerBase(std::move(token_model_path), has_bos, has_eos) {
  if (token_model_path_.empty() || token_model_path_ == "<none>") {
    bos_id_ = 1, eos_id_ = 2;
    num_token_ = stub_vocab;  // overwritten by the checkpoint header (model.cpp)
    LOG(INFO) << "no tokenizer.json given: using the id-level stand-in tokenizer";
    return;
  }
  bpe_ = std::make_unique<ByteBpeModel>();
  const std::string err = bpe_->load(token_model_path_);
  if (!err.empty()) {
    LOG(FATAL) << "The token model path is not valid, please check the path and type of token model: "
               << token_model_path_ << ": " << err;
  }
  bos_id_ = bpe_->token_to_id(bos);
  eos_id_ = bpe_->token_to_id(eos);
  stop_token1_ = eos_id_;
  stop_token2_ = bpe_->token_to_id(stop2);
  num_token_ = bpe_->vocab_size();
}

BpeEncodeLayer::BpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos)
    : BpeEncodeLayer(std::move(token_model_path), has_bos, has_eos, "<|begin_of_text|>", "<|end_of_text|>",
                     "<|eot_id|>", 128256) {}

std::vector<int32_t> BpeEncodeLayer::encode(const std::string& sentence) const {
  if (!bpe_) return bytes_to_ids(sentence, has_bos_, has_eos_, num_token_);
  std::vector<int32_t> ids = bpe_->encode(sentence);
  if (has_bos_ && bos_id_ >= 0) ids.insert(ids.begin(), bos_id_);
  if (has_eos_ && eos_id_ >= 0) ids.push_back(eos_id_);
  return ids;
}
std::string BpeEncodeLayer::decode(int32_t token_id) const { return decode(std::vector<int32_t>{token_id}); }
std::string BpeEncodeLayer::decode(const std::vector<int32_t>& token_ids) const {
  return bpe_ ? bpe_->decode(token_ids) : ids_to_text(token_ids);
}
bool BpeEncodeLayer::is_sentence_ending(int32_t token_id) const {
  // stand-in (synthetic checkpoints): always decode the requested number of steps
  return bpe_ != nullptr && token_id >= 0 && (token_id == stop_token1_ || token_id == stop_token2_);
}
int32_t BpeEncodeLayer::vocab_size() const { return num_token_; }

QwenEncodeLayer::QwenEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos)
    : BpeEncodeLayer(std::move(token_model_path), has_bos, has_eos, "<|im_start|>", "<|im_end|>", "<|endoftext|>",
                     151936) {}
}  // namespace op
