// kuiper_decode: drive model::LLama2Model / Qwen2Model with raw token ids, the way demo/main.cpp
// drives it with text (embedding -> fill_input -> predict per position).
//
//   kuiper_decode <checkpoint> <llama|qwen> <fp32|int8> <n_steps> <id0> [id1 ...]
//                 [--layers] [--copy-at K] [--logits out.f32]
//
// ids are the prompt; after the prompt the model free-runs greedily until n_steps positions
// have been processed.  Prints the id chosen at every position (-1 for prompt steps before the
// last prompt token) on one line.  --layers uses Model::forward (layer-by-layer op registry path)
// instead of predict's fused decoder.  --copy-at K hands predict() a COPY of the embedding row at
// position K (so that step cannot be recognised and runs layer by layer in the middle of a sequence
// the fused decoder started).  --logits writes the last position's logits as raw fp32.
#include <base/base.h>
#include <cuda_runtime_api.h>
#include <glog/logging.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "model/llama3.h"
#include "model/qwen2.h"

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s <checkpoint> <llama|qwen> <fp32|int8> <n_steps> <id0> [id1 ...] "
                         "[--layers] [--copy-at K] [--logits out.f32]\n", argv[0]);
    return 2;
  }
  const std::string checkpoint = argv[1], family = argv[2], prec = argv[3];
  const int n_steps = std::atoi(argv[4]);
  std::vector<int> prompt;
  bool layers = false;
  int copy_at = -1;
  std::string logits_path;
  for (int i = 5; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--layers")) layers = true;
    else if (!std::strcmp(argv[i], "--copy-at") && i + 1 < argc) copy_at = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--logits") && i + 1 < argc) logits_path = argv[++i];
    else prompt.push_back(std::atoi(argv[i]));
  }
  if (prompt.empty() || n_steps <= 0) return 2;
  const bool quant = prec == "int8";

  std::unique_ptr<model::LLama2Model> m;
  if (family == "qwen") {
    m = std::make_unique<model::Qwen2Model>(base::TokenizerType::kEncodeBpe, "<none>", checkpoint, quant);
  } else {
    m = std::make_unique<model::LLama2Model>(base::TokenizerType::kEncodeSpe, "<none>", checkpoint, quant);
  }
  base::Status st = m->init(base::DeviceType::kDeviceCUDA);
  if (!st) {
    std::fprintf(stderr, "init failed: %s\n", st.get_err_msg().c_str());
    return 1;
  }
  std::fprintf(stderr, "engine: %s%s\n", m->decoder_engine(), layers ? " (unused: --layers)" : "");

  tensor::Tensor pos_tensor = m->get_buffer(model::ModelBufferType::kInputPos);
  const int32_t prompt_len = static_cast<int32_t>(prompt.size());
  auto prompt_embedding = m->embedding(prompt);
  std::vector<float> host_logits;
  int next = -1;
  std::vector<int> chosen;
  auto run = [&](const tensor::Tensor& input, bool is_prompt) {
    if (!layers) {
      if (pos_tensor.index<int32_t>(0) == copy_at) {
        // the row is a view into the model's embedding buffer: copy it into storage of its own
        tensor::Tensor copy(base::DataType::kDataTypeFp32, static_cast<int32_t>(input.size()), true,
                            base::CUDADeviceAllocatorFactory::get_instance());
        CHECK(cudaMemcpy(copy.ptr<float>(), input.ptr<float>(), input.byte_size(), cudaMemcpyDeviceToDevice) ==
              cudaSuccess);
        STATUS_CHECK(m->predict(copy, pos_tensor, is_prompt, next));
      } else {
        STATUS_CHECK(m->predict(input, pos_tensor, is_prompt, next));
      }
      return;
    }
    STATUS_CHECK(m->forward(input, pos_tensor, next));
    next = -1;
    if (!is_prompt) {  // greedy argmax, lowest index on ties (argmax_sampler.cpp)
      tensor::Tensor lg = m->get_buffer(model::ModelBufferType::kForwardOutput).clone();
      lg.to_cpu();
      const float* p = lg.ptr<float>();
      size_t best = 0;
      for (size_t i = 1; i < lg.size(); ++i)
        if (p[i] > p[best]) best = i;
      next = static_cast<int>(best);
    }
  };
  for (int32_t pos = 0; pos < n_steps; ++pos) {
    pos_tensor.index<int32_t>(0) = pos;
    if (pos < prompt_len - 1) {
      run(m->fill_input(pos_tensor, prompt_embedding, true), true);
    } else if (pos == prompt_len - 1) {
      // the last prompt token already samples (demo/main.cpp:19-24 treats it as a prompt row but
      // is_prompt=false from pos == prompt_len-1 on)
      tensor::Tensor row = m->fill_input(pos_tensor, prompt_embedding, true);
      run(row, false);
    } else {
      auto emb = m->embedding({next});
      run(m->fill_input(pos_tensor, emb, false), false);
    }
    chosen.push_back(next);
  }
  for (size_t i = 0; i < chosen.size(); ++i) std::printf("%s%d", i ? " " : "", chosen[i]);
  std::printf("\n");
  if (!logits_path.empty() && m->tensor_parallel().rank == 0) {  // under kuiper_tp_launch every rank holds the same logits
    tensor::Tensor lg = m->get_buffer(model::ModelBufferType::kForwardOutput).clone();
    lg.to_cpu();
    FILE* f = std::fopen(logits_path.c_str(), "wb");
    if (!f) return 1;
    std::fwrite(lg.ptr<float>(), sizeof(float), lg.size(), f);
    std::fclose(f);
  }
  return 0;
}
