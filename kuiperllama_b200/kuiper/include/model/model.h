// model::Model -- what every decoder model of the kuiper:: API is to its caller:
//
//   text     encode() / decode() / is_sentence_ending()           (the tokenizer front end)
//   tokens   embedding(ids) -> rows; fill_input(pos, rows, is_prompt) -> the row for this position
//   step     predict(row, pos, is_prompt, next): one position; greedy id in `next` unless is_prompt
//            forward(row, pos, next): the same without the sampling step
//   state    get_buffer(ModelBufferType): the named tensors (kInputPos is how the demo passes the
//            position; kForwardOutput holds the logits), slice_kv_cache(layer, pos)
//
// and to its subclass: the loading pipeline init() drives --
//   gen_model_from_file(): create_encode_layer() -> read_model_file() (mmap + header) -> create_layers()
// Method names and signatures follow reference kuiper/include/model/model.h so demo/main.cpp and
// subclasses written against it compile unchanged.
#ifndef KLLM_KUIPER_MODEL_MODEL_H_
#define KLLM_KUIPER_MODEL_MODEL_H_
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "checkpoint_file.h"
#include "op/decoder_layers.h"
#include "op/encode.h"
#include "sampler/argmax_sampler.h"
#include "tensor/tensor.h"

namespace model {
class Model {
 public:
  Model(base::TokenizerType tokenizer_type, base::ModelType model_type, std::string token_path,
        std::string model_path, bool is_quant_model);
  virtual ~Model() = default;

  base::ModelType model_type() const { return model_type_; }
  const std::string& token_path() const { return token_path_; }
  const std::string& model_path() const { return model_path_; }

  // ---- life cycle and the per-position step (implemented by the model family) --------------------
  virtual base::Status init(base::DeviceType device_type) = 0;
  virtual base::Status predict(const tensor::Tensor& input, const tensor::Tensor& pos_tensor, bool is_prompt,
                               int& next) const = 0;
  virtual base::Status forward(const tensor::Tensor& input, const tensor::Tensor& pos_tensor, int& next) const = 0;
  virtual op::EmbeddingOutput embedding(const std::vector<int>& tokens) const = 0;

  // ---- text -----------------------------------------------------------------------------------------
  virtual std::vector<int32_t> encode(const std::string& sentence) const;
  virtual std::string decode(int32_t token_idx) const;
  virtual std::string decode(std::vector<int32_t> token_idxs) const;
  virtual bool is_sentence_ending(int32_t token_idx) const;

  // ---- state ----------------------------------------------------------------------------------------
  virtual tensor::Tensor& get_buffer(ModelBufferType buffer_idx);
  virtual const tensor::Tensor& get_buffer(ModelBufferType buffer_idx) const;
  // (key, value) views of cache row [layer_idx, token_pos, :]
  virtual std::pair<tensor::Tensor, tensor::Tensor> slice_kv_cache(int32_t layer_idx, int32_t token_pos) const;
  // the embedding row that is the model input at this position: row `pos` of a prompt's embeddings,
  // row 0 otherwise (a view; the layer-by-layer path updates it in place)
  virtual tensor::Tensor fill_input(const tensor::Tensor& pos_tensor, const op::EmbeddingOutput& embedding_output,
                                    bool is_prompt) const;

 protected:
  friend struct ModelInspector;  // tools/kuiper_selftest.cpp: drives the loading pipeline without a GPU
  // loading pipeline
  virtual base::Status gen_model_from_file();
  virtual base::Status create_encode_layer();
  virtual base::Status read_model_file();
  virtual base::Status generate_model_infos(const ModelConfig& config) const;
  virtual base::Status insert_buffer(ModelBufferType buffer_idx, const tensor::Tensor& tensor);
  virtual int32_t post_processing(const tensor::Tensor& pos, bool is_prompt) const = 0;

 private:
  // hooks of the model family, called from the pipeline above
  virtual void init_mem() = 0;
  virtual base::Status create_layers() = 0;
  virtual void create_param_layers() = 0;
  virtual void create_nonparam_layers() = 0;
  virtual void create_param_quant_layers() = 0;

 protected:
  bool is_quant_model_ = false;
  int32_t group_size_ = 1;  // int8 files: from the header
  std::string token_path_, model_path_;
  base::DeviceType device_type_ = base::DeviceType::kDeviceUnknown;
  base::ModelType model_type_ = base::ModelType::kModelTypeUnknown;
  base::TokenizerType tokenizer_type_ = base::TokenizerType::kEncodeUnknown;
  std::unique_ptr<TransformerConfig> config_;
  std::shared_ptr<RawModelData> raw_model_data_;
  std::unique_ptr<op::EncodeLayerBase> encode_layer_;
  std::unique_ptr<sampler::Sampler> sampler_;
  mutable std::map<ModelBufferType, tensor::Tensor> buffers_;
};
}  // namespace model
#endif  // KLLM_KUIPER_MODEL_MODEL_H_
