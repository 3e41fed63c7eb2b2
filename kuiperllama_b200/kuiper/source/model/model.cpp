// model::Model: checkpoint file, named buffers, tokenizer (reference kuiper/source/model/model.cpp).
#include "model/model.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>

namespace model {
Model::Model(base::TokenizerType tokenizer_type, base::ModelType model_type, std::string token_path,
             std::string model_path, bool is_quant_model)
    : is_quant_model_(is_quant_model),
      token_path_(std::move(token_path)),
      model_path_(std::move(model_path)),
      model_type_(model_type),
      tokenizer_type_(tokenizer_type) {}

base::ModelType Model::model_type() const { return model_type_; }
const std::string& Model::token_path() const { return token_path_; }
const std::string& Model::model_path() const { return model_path_; }

base::Status Model::insert_buffer(ModelBufferType buffer_idx, const tensor::Tensor& tensor) {
  if (buffers_.count(buffer_idx) > 0)
    return base::error::KeyHasExits(std::to_string(int(buffer_idx)) + " has exits in the buffers");
  if (tensor.is_empty()) return base::error::InvalidArgument("The tensor is empty for inserting buffer.");
  buffers_.insert({buffer_idx, tensor});
  return base::error::Success();
}

tensor::Tensor& Model::get_buffer(ModelBufferType buffer_idx) {
  CHECK_GT(buffers_.count(buffer_idx), 0u) << int(buffer_idx);
  return buffers_.at(buffer_idx);
}
const tensor::Tensor& Model::get_buffer(ModelBufferType buffer_idx) const {
  CHECK_GT(buffers_.count(buffer_idx), 0u) << int(buffer_idx);
  return buffers_.at(buffer_idx);
}

base::Status Model::read_model_file() {
  using namespace base;
  if (model_path_.empty()) return error::PathNotValid("Failed to open the weight file, the model path is empty!");
  const int32_t fd = open(model_path_.c_str(), O_RDONLY);
  if (fd == -1)
    return error::PathNotValid("Failed to open the weight file " + model_path_ + " may be the path does not exist!");
  struct stat st {};
  if (fstat(fd, &st) == -1) {
    close(fd);
    return error::ModelParseError("Failed to retrieve the file size information from the model file.");
  }
  const size_t header = sizeof(ModelConfig) + (is_quant_model_ ? sizeof(int32_t) : 0);
  if (static_cast<size_t>(st.st_size) < header) {
    close(fd);
    return error::ModelParseError("Failed to retrieve the configuration information from the model file.");
  }
  void* map = mmap(nullptr, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (map == MAP_FAILED || map == nullptr) {
    close(fd);
    return error::ModelParseError("Failed to map the weight file " + model_path_ + " into memory.");
  }
  if (is_quant_model_) {
    raw_model_data_ = std::make_shared<RawModelDataInt8>();
  } else {
    raw_model_data_ = std::make_shared<RawModelDataFp32>();
  }
  raw_model_data_->fd = fd;
  raw_model_data_->file_size = st.st_size;
  raw_model_data_->data = map;
  raw_model_data_->weight_data = static_cast<int8_t*>(map) + header;

  ModelConfig config{};
  std::memcpy(&config, map, sizeof(ModelConfig));
  if (is_quant_model_) std::memcpy(&group_size_, static_cast<int8_t*>(map) + sizeof(ModelConfig), sizeof(int32_t));
  base::Status gen = generate_model_infos(config);
  if (!gen) return gen;

  LOG(INFO) << "The model path: " << model_path_ << " (" << raw_model_data_->file_size << " byte, "
            << (is_quant_model_ ? "int8 group-quantised" : "fp32") << ")";
  LOG(INFO) << "The tokenizer path: " << token_path_;
  LOG(INFO) << "\nThe model info: " << *config_;
  return error::Success();
}

base::Status Model::generate_model_infos(const ModelConfig& config) const {
  if (config.dim <= 0 || config.head_num <= 0 || config.kv_head_num <= 0 || config.layer_num <= 0 ||
      config.hidden_dim <= 0 || config.seq_len <= 0 || config.vocab_size == 0 ||
      config.dim % config.head_num != 0 || config.head_num % config.kv_head_num != 0)
    return base::error::ModelParseError("The checkpoint header holds an impossible configuration.");
  config_->dim_ = config.dim;
  config_->hidden_dim_ = config.hidden_dim;
  config_->layer_num_ = config.layer_num;
  config_->head_num_ = config.head_num;
  config_->kv_head_num_ = config.kv_head_num;
  config_->seq_len_ = config.seq_len;
  config_->kv_dim_ = (config.dim * config.kv_head_num) / config.head_num;
  config_->kv_mul_ = config.head_num / config.kv_head_num;
  config_->head_size_ = config.dim / config.head_num;
  config_->is_shared_weight_ = config.vocab_size > 0;  // negative = separate classifier
  config_->vocab_size_ = std::abs(config.vocab_size);
  return base::error::Success();
}

base::Status Model::create_encode_layer() {
  using namespace base;
  if (tokenizer_type_ == TokenizerType::kEncodeSpe) {
    encode_layer_ = std::make_unique<op::SpeEncodeLayer>(token_path_, true, false);
  } else {
#if defined(QWEN2_SUPPORT)
    encode_layer_ = std::make_unique<op::QwenEncodeLayer>(token_path_, false, false);
#else
    encode_layer_ = std::make_unique<op::BpeEncodeLayer>(token_path_, true, false);
#endif
  }
  if (!encode_layer_) return error::InternalError("Create the encode layer failed.");
  config_->vocab_size_ = encode_layer_->vocab_size();
  if (config_->vocab_size_ <= 0) return error::InternalError("The vocab size param read error from the model file!");
  return error::Success();
}

base::Status Model::gen_model_from_file() {
  config_ = std::make_unique<TransformerConfig>();
  base::Status st = create_encode_layer();
  if (!st) {
    LOG(ERROR) << "Create the encode layer failed! " << st.get_err_msg();
    return st;
  }
  st = read_model_file();
  if (!st) {
    LOG(ERROR) << "Read model file " << model_path_ << " failed! " << st.get_err_msg();
    return st;
  }
  st = create_layers();
  if (!st) LOG(ERROR) << "Create layers for the model file " << model_path_ << " failed! " << st.get_err_msg();
  return st;
}

std::vector<int32_t> Model::encode(const std::string& sentence) const {
  CHECK(encode_layer_ != nullptr);
  return encode_layer_->encode(sentence);
}
bool Model::is_sentence_ending(int32_t token_idx) const {
  CHECK(encode_layer_ != nullptr);
  return encode_layer_->is_sentence_ending(token_idx);
}
std::string Model::decode(int32_t token_idx) const {
  CHECK(encode_layer_ != nullptr);
  return encode_layer_->decode(token_idx);
}
std::string Model::decode(std::vector<int32_t> token_idxs) const {
  CHECK(encode_layer_ != nullptr);
  return encode_layer_->decode(token_idxs);
}

std::pair<tensor::Tensor, tensor::Tensor> Model::slice_kv_cache(int32_t layer_idx, int32_t token_pos) const {
  const int64_t row = (static_cast<int64_t>(layer_idx) * config_->seq_len_ + token_pos) * config_->kv_dim_;
  auto view = [&](ModelBufferType which) {
    float* p = const_cast<float*>(get_buffer(which).ptr<float>(row));
    tensor::Tensor t(base::DataType::kDataTypeFp32, config_->kv_dim_, false, nullptr, p);
    t.set_device_type(device_type_);
    return t;
  };
  return {view(ModelBufferType::kKeyCache), view(ModelBufferType::kValueCache)};
}

tensor::Tensor Model::fill_input(const tensor::Tensor& pos_tensor, const op::EmbeddingOutput& embedding_output,
                                 bool is_prompt) const {
  const int32_t pos = pos_tensor.index<int32_t>(0);
  const int32_t row = is_prompt ? pos : 0;
  float* p = const_cast<float*>(embedding_output.input_embeddings.ptr<float>(static_cast<int64_t>(row) * config_->dim_));
  auto view = std::make_shared<base::Buffer>(config_->dim_ * sizeof(float), nullptr, p, true);
  tensor::Tensor input(base::DataType::kDataTypeFp32, config_->dim_);
  input.assign(view);
  input.set_device_type(device_type_);
  return input;
}
}  // namespace model
