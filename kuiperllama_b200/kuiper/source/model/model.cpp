// model::Model: the checkpoint mapping, the named buffers and the tokenizer plumbing shared by the
// model families (see model/model.h).
#include "model/model.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>

namespace model {
RawModelData::~RawModelData() {
  if (data != nullptr && data != MAP_FAILED) munmap(data, file_size);
  if (fd != -1) close(fd);
}

Model::Model(base::TokenizerType tokenizer_type, base::ModelType model_type, std::string token_path,
             std::string model_path, bool is_quant_model)
    : is_quant_model_(is_quant_model),
      token_path_(std::move(token_path)),
      model_path_(std::move(model_path)),
      model_type_(model_type),
      tokenizer_type_(tokenizer_type) {}

// ---- named buffers -------------------------------------------------------------------------------------
base::Status Model::insert_buffer(ModelBufferType buffer_idx, const tensor::Tensor& tensor) {
  if (tensor.is_empty()) return base::error::InvalidArgument("The tensor is empty for inserting buffer.");
  if (!buffers_.emplace(buffer_idx, tensor).second)
    return base::error::KeyHasExits("buffer " + std::to_string(int(buffer_idx)) + " is registered already");
  return base::error::Success();
}

const tensor::Tensor& Model::get_buffer(ModelBufferType buffer_idx) const {
  const auto it = buffers_.find(buffer_idx);
  CHECK(it != buffers_.end()) << "model buffer " << int(buffer_idx) << " does not exist";
  return it->second;
}
tensor::Tensor& Model::get_buffer(ModelBufferType buffer_idx) {
  return const_cast<tensor::Tensor&>(std::as_const(*this).get_buffer(buffer_idx));
}

std::pair<tensor::Tensor, tensor::Tensor> Model::slice_kv_cache(int32_t layer_idx, int32_t token_pos) const {
  const int32_t width = config_->kv_dim_;
  const int64_t first = (static_cast<int64_t>(layer_idx) * config_->seq_len_ + token_pos) * width;
  auto row_of = [&](ModelBufferType cache) {
    auto* p = const_cast<float*>(get_buffer(cache).ptr<float>(first));
    tensor::Tensor row(base::DataType::kDataTypeFp32, width, false, nullptr, p);
    row.set_device_type(device_type_);
    return row;
  };
  return {row_of(ModelBufferType::kKeyCache), row_of(ModelBufferType::kValueCache)};
}

tensor::Tensor Model::fill_input(const tensor::Tensor& pos_tensor, const op::EmbeddingOutput& embedding_output,
                                 bool is_prompt) const {
  // a prompt's embeddings hold one row per prompt position; afterwards there is only the new token's
  const int64_t row = is_prompt ? pos_tensor.index<int32_t>(0) : 0;
  const int32_t dim = config_->dim_;
  auto* p = const_cast<float*>(embedding_output.input_embeddings.ptr<float>(row * dim));
  tensor::Tensor input(base::DataType::kDataTypeFp32, dim);
  CHECK(input.assign(std::make_shared<base::Buffer>(dim * sizeof(float), nullptr, p, true)));
  input.set_device_type(device_type_);
  return input;
}

// ---- text: straight to the encode layer -------------------------------------------------------------------
std::vector<int32_t> Model::encode(const std::string& sentence) const {
  CHECK(encode_layer_ != nullptr);
  return encode_layer_->encode(sentence);
}
std::string Model::decode(int32_t token_idx) const { return decode(std::vector<int32_t>{token_idx}); }
std::string Model::decode(std::vector<int32_t> token_idxs) const {
  CHECK(encode_layer_ != nullptr);
  return encode_layer_->decode(token_idxs);
}
bool Model::is_sentence_ending(int32_t token_idx) const {
  CHECK(encode_layer_ != nullptr);
  return encode_layer_->is_sentence_ending(token_idx);
}

// ---- loading ------------------------------------------------------------------------------------------------
base::Status Model::gen_model_from_file() {
  config_ = std::make_unique<TransformerConfig>();
  struct Stage {
    const char* what;
    base::Status (Model::*run)();
  };
  const Stage stages[] = {{"creating the tokenizer", &Model::create_encode_layer},
                          {"reading the checkpoint", &Model::read_model_file},
                          {"creating the layers", &Model::create_layers}};
  for (const Stage& stage : stages) {
    const base::Status st = (this->*stage.run)();
    if (!st) {
      LOG(ERROR) << "model " << model_path_ << ": " << stage.what << " failed: " << st.get_err_msg();
      return st;
    }
  }
  return base::error::Success();
}

base::Status Model::create_encode_layer() {
  switch (tokenizer_type_) {
    case base::TokenizerType::kEncodeSpe:
      encode_layer_ = std::make_unique<op::SpeEncodeLayer>(token_path_, /*has_bos=*/true, /*has_eos=*/false);
      break;
    default:  // byte-level BPE; which family is a build-time choice, as in the reference
#if defined(QWEN2_SUPPORT)
      encode_layer_ = std::make_unique<op::QwenEncodeLayer>(token_path_, /*has_bos=*/false, /*has_eos=*/false);
#else
      encode_layer_ = std::make_unique<op::BpeEncodeLayer>(token_path_, /*has_bos=*/true, /*has_eos=*/false);
#endif
  }
  // provisional: the checkpoint header has the last word on the vocabulary size
  config_->vocab_size_ = encode_layer_->vocab_size();
  if (config_->vocab_size_ <= 0) return base::error::InternalError("The tokenizer reports an empty vocabulary.");
  return base::error::Success();
}

base::Status Model::read_model_file() {
  namespace err = base::error;
  if (model_path_.empty()) return err::PathNotValid("The checkpoint path is empty.");
  const int fd = open(model_path_.c_str(), O_RDONLY);
  if (fd == -1) return err::PathNotValid("Cannot open the checkpoint " + model_path_ + " (does it exist?)");

  // header: 7 int32 (+ the group size for int8 files), then the payload
  const size_t header_bytes = sizeof(ModelConfig) + (is_quant_model_ ? sizeof(int32_t) : 0);
  struct stat info {};
  void* map = MAP_FAILED;
  if (fstat(fd, &info) == 0 && static_cast<size_t>(info.st_size) >= header_bytes)
    map = mmap(nullptr, info.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (map == MAP_FAILED || map == nullptr) {
    close(fd);
    return err::ModelParseError("Cannot map the checkpoint " + model_path_ + " (shorter than its header, or mmap failed).");
  }
  if (is_quant_model_) raw_model_data_ = std::make_shared<RawModelDataInt8>();
  else raw_model_data_ = std::make_shared<RawModelDataFp32>();
  RawModelData& raw = *raw_model_data_;
  raw.fd = fd;
  raw.file_size = static_cast<size_t>(info.st_size);
  raw.data = map;
  raw.weight_data = static_cast<char*>(map) + header_bytes;

  ModelConfig header{};
  std::memcpy(&header, map, sizeof(header));
  if (is_quant_model_) std::memcpy(&group_size_, static_cast<char*>(map) + sizeof(header), sizeof(int32_t));
  const base::Status st = generate_model_infos(header);
  if (!st) return st;
  LOG(INFO) << "checkpoint " << model_path_ << ": " << raw.file_size << " bytes, "
            << (is_quant_model_ ? "int8 group-quantised" : "fp32") << "; tokenizer " << token_path_;
  LOG(INFO) << "\nThe model info: " << *config_;
  return err::Success();
}

base::Status Model::generate_model_infos(const ModelConfig& h) const {
  const bool plausible = h.dim > 0 && h.hidden_dim > 0 && h.layer_num > 0 && h.seq_len > 0 && h.vocab_size != 0 &&
                         h.head_num > 0 && h.kv_head_num > 0 && h.dim % h.head_num == 0 &&
                         h.head_num % h.kv_head_num == 0;
  if (!plausible) return base::error::ModelParseError("The checkpoint header holds an impossible configuration.");
  TransformerConfig& c = *config_;
  c.dim_ = h.dim, c.hidden_dim_ = h.hidden_dim, c.layer_num_ = h.layer_num, c.seq_len_ = h.seq_len;
  c.head_num_ = h.head_num, c.kv_head_num_ = h.kv_head_num;
  c.head_size_ = h.dim / h.head_num;
  c.kv_mul_ = h.head_num / h.kv_head_num;
  c.kv_dim_ = c.head_size_ * h.kv_head_num;
  c.is_shared_weight_ = h.vocab_size > 0;  // the sign is a flag: negative = a separate classifier follows
  c.vocab_size_ = std::abs(h.vocab_size);
  return base::error::Success();
}
}  // namespace model
