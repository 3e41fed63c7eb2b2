// Encode layers: one shared implementation over a TokenizerBackend (see op/encode.h).
#include "op/encode.h"

#include <sstream>
#include <utility>

#include "op/byte_bpe.h"
#ifdef KLLM_WITH_SENTENCEPIECE
#include <sentencepiece_processor.h>
#else
#include "op/spm_bpe.h"
#endif

namespace op {
namespace {
[[noreturn]] void refuse(const std::string& path, const std::string& why) {
  LOG(FATAL) << "The token model path is not valid, please check the path and type of token model: " << path
             << ": " << why;
  std::abort();
}

// Synthetic checkpoints have no tokenizer file: BOS-free, one id per byte (offset past the control
// ids) in, "<12><7>" out -- lossless and obviously synthetic.
class StandInBackend final : public TokenizerBackend {
 public:
  explicit StandInBackend(int32_t vocab) : vocab_(vocab) {}
  std::vector<int32_t> encode(const std::string& text) const override {
    std::vector<int32_t> ids;
    for (unsigned char c : text) ids.push_back(3 + static_cast<int32_t>(c) % (vocab_ > 259 ? 256 : 1));
    return ids;
  }
  std::string decode(const std::vector<int32_t>& ids) const override {
    std::ostringstream os;
    for (int32_t id : ids) os << '<' << id << '>';
    return os.str();
  }
  int32_t vocab_size() const override { return vocab_; }
  int32_t id_of(const std::string&) const override { return -1; }

 private:
  int32_t vocab_;
};

#ifdef KLLM_WITH_SENTENCEPIECE
class SentencePieceBackend final : public TokenizerBackend {
 public:
  explicit SentencePieceBackend(const std::string& path) {
    if (!spe_.Load(path).ok()) refuse(path, "libsentencepiece cannot load it");
  }
  std::vector<int32_t> encode(const std::string& text) const override { return spe_.EncodeAsIds(text); }
  std::string decode(const std::vector<int32_t>& ids) const override { return spe_.DecodeIds(ids); }
  int32_t vocab_size() const override { return spe_.GetPieceSize(); }
  int32_t id_of(const std::string& tok) const override { return tok == "<s>" ? spe_.bos_id() : tok == "</s>" ? spe_.eos_id() : -1; }

 private:
  sentencepiece::SentencePieceProcessor spe_;
};
#else
class SpmBackend final : public TokenizerBackend {
 public:
  explicit SpmBackend(const std::string& path) {
    const std::string err = model_.load(path);
    if (!err.empty()) refuse(path, err);
  }
  std::vector<int32_t> encode(const std::string& text) const override { return model_.encode(text); }
  std::string decode(const std::vector<int32_t>& ids) const override { return model_.decode(ids); }
  int32_t vocab_size() const override { return model_.piece_size(); }
  int32_t id_of(const std::string& tok) const override {
    return tok == "<s>" ? model_.bos_id() : tok == "</s>" ? model_.eos_id() : -1;
  }

 private:
  SpmBpeModel model_;
};
#endif

class ByteBpeBackend final : public TokenizerBackend {
 public:
  explicit ByteBpeBackend(const std::string& path) {
    const std::string err = model_.load(path);
    if (!err.empty()) refuse(path, err);
  }
  std::vector<int32_t> encode(const std::string& text) const override { return model_.encode(text); }
  std::string decode(const std::vector<int32_t>& ids) const override { return model_.decode(ids); }
  int32_t vocab_size() const override { return model_.vocab_size(); }
  int32_t id_of(const std::string& tok) const override { return model_.token_to_id(tok); }

 private:
  ByteBpeModel model_;
};
}  // namespace

// ---- the shared implementation ------------------------------------------------------------------------------
EncodeLayerBase::EncodeLayerBase(std::string token_model_path, bool has_bos, bool has_eos)
    : Layer(base::DeviceType::kDeviceCPU, LayerType::kLayerEncode, "Encode"),
      has_bos_(has_bos),
      has_eos_(has_eos),
      token_model_path_(std::move(token_model_path)) {}
EncodeLayerBase::~EncodeLayerBase() = default;

void EncodeLayerBase::adopt(std::unique_ptr<TokenizerBackend> backend, int32_t bos_id, int32_t eos_id,
                            int32_t extra_stop_id, bool stops_generation) {
  backend_ = std::move(backend);
  bos_id_ = bos_id, eos_id_ = eos_id, stop_token2_ = extra_stop_id;
  stops_generation_ = stops_generation;
}

std::vector<int32_t> EncodeLayerBase::encode(const std::string& sentence) const {
  CHECK(backend_ != nullptr);
  std::vector<int32_t> ids = backend_->encode(sentence);
  if (has_bos_ && bos_id_ >= 0) ids.insert(ids.begin(), bos_id_);
  if (has_eos_ && eos_id_ >= 0) ids.push_back(eos_id_);
  return ids;
}
std::string EncodeLayerBase::decode(const std::vector<int32_t>& token_ids) const {
  CHECK(backend_ != nullptr);
  return backend_->decode(token_ids);
}
std::string EncodeLayerBase::decode(int32_t token_id) const { return decode(std::vector<int32_t>{token_id}); }
bool EncodeLayerBase::is_sentence_ending(int32_t token_id) const {
  return stops_generation_ && token_id >= 0 && (token_id == eos_id_ || token_id == stop_token2_);
}
int32_t EncodeLayerBase::vocab_size() const {
  CHECK(backend_ != nullptr);
  return backend_->vocab_size();
}

// ---- the families -----------------------------------------------------------------------------------------------
SpeEncodeLayer::SpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos)
    : EncodeLayerBase(std::move(token_model_path), has_bos, has_eos) {
  if (is_stand_in_path(token_model_path_)) {
    LOG(INFO) << "no tokenizer model given: using the id-level stand-in tokenizer";
    adopt(std::make_unique<StandInBackend>(32000), /*bos=*/1, /*eos=*/2, -1, /*stops_generation=*/false);
    return;
  }
#ifdef KLLM_WITH_SENTENCEPIECE
  auto backend = std::make_unique<SentencePieceBackend>(token_model_path_);
#else
  auto backend = std::make_unique<SpmBackend>(token_model_path_);
#endif
  const int32_t bos = backend->id_of("<s>"), eos = backend->id_of("</s>");
  adopt(std::move(backend), bos, eos, -1, true);
}

BpeEncodeLayer::BpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos, const char* bos,
                               const char* eos, const char* extra_stop, int32_t stand_in_vocab)
    : EncodeLayerBase(std::move(token_model_path), has_bos, has_eos) {
  if (is_stand_in_path(token_model_path_)) {
    LOG(INFO) << "no tokenizer.json given: using the id-level stand-in tokenizer";
    // the vocabulary size is overwritten by the checkpoint header anyway (model.cpp)
    adopt(std::make_unique<StandInBackend>(stand_in_vocab), 1, 2, -1, false);
    return;
  }
  auto backend = std::make_unique<ByteBpeBackend>(token_model_path_);
  const int32_t b = backend->id_of(bos), e = backend->id_of(eos), x = backend->id_of(extra_stop);
  adopt(std::move(backend), b, e, x, true);
}

BpeEncodeLayer::BpeEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos)
    : BpeEncodeLayer(std::move(token_model_path), has_bos, has_eos, "<|begin_of_text|>", "<|end_of_text|>",
                     "<|eot_id|>", 128256) {}

QwenEncodeLayer::QwenEncodeLayer(std::string token_model_path, bool has_bos, bool has_eos)
    : BpeEncodeLayer(std::move(token_model_path), has_bos, has_eos, "<|im_start|>", "<|im_end|>", "<|endoftext|>",
                     151936) {}
}  // namespace op
