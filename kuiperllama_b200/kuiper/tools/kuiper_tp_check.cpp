// kuiper_tp_check: the tensor-parallel plumbing of the C++ host side, checked WITHOUT a GPU.
//
//   kuiper_tp_check shard <dim> <hidden> <layers> <heads> <kv_heads> <vocab> <group|0> <world> <rank>
//       prints the shard of model/tensor_parallel.h (tests compare it with kuiperllama_b200/tensor_parallel.py)
//   kuiper_tp_check rendezvous <world> <rank> <port>
//       one rank of a TCP rendezvous on 127.0.0.1: all_gather of a rank-stamped 64-byte blob, a barrier,
//       a second all_gather; exit code 0 when every byte is where it belongs
//   kuiper_tp_check load <checkpoint> <llama|qwen> <fp32|int8> <world> <rank>
//       runs the model's loading pipeline (mmap, header, layers) for one rank -- no init(), nothing
//       uploaded -- and prints dims + FNV-1a hashes of every layer matrix (and int8 scales, biases) this
//       rank would upload: the load-time sharding against shard_weights() of the Python side
#include <base/base.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "model/llama3.h"
#include "model/qwen2.h"
#include "model/tensor_parallel.h"
#include "op/decoder_layers.h"

namespace model {
struct ModelInspector {
  static base::Status load(LLama2Model& m) { return m.gen_model_from_file(); }
  static const TransformerConfig& config(const LLama2Model& m) { return *m.config_; }
  static const LLama2Layers& layers(const LLama2Model& m) { return *m.llama_layers_; }
};
}  // namespace model

namespace {
uint64_t fnv1a(const void* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  const unsigned char* c = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) h = (h ^ c[i]) * 1099511628211ull;
  return h;
}

int do_shard(char** a) {
  model::TransformerConfig c;
  c.dim_ = std::atoi(a[0]), c.hidden_dim_ = std::atoi(a[1]), c.layer_num_ = std::atoi(a[2]);
  c.head_num_ = std::atoi(a[3]), c.kv_head_num_ = std::atoi(a[4]), c.vocab_size_ = std::atoi(a[5]);
  const int group = std::atoi(a[6]), world = std::atoi(a[7]), rank = std::atoi(a[8]);
  c.head_size_ = c.dim_ / c.head_num_, c.kv_mul_ = c.head_num_ / c.kv_head_num_;
  c.kv_dim_ = c.head_size_ * c.kv_head_num_;
  model::TpShard s;
  const base::Status st = model::tp_shard(c, group, world, rank, &s);
  if (!st) {
    std::printf("error %s\n", st.get_err_msg().c_str());
    return 3;
  }
  std::printf("q %d %d k %d %d f %d %d heads %d kv_heads %d hidden %d comm_words %d\n", s.q0, s.q1, s.k0, s.k1, s.f0,
              s.f1, s.head_num, s.kv_head_num, s.hidden_dim, model::tp_comm_words(c, world));
  return 0;
}

int do_rendezvous(char** a) {
  model::TpConfig cfg;
  cfg.world = std::atoi(a[0]), cfg.rank = std::atoi(a[1]), cfg.port = std::atoi(a[2]);
  model::TpRendezvous r;
  base::Status st = r.open(cfg, 30);
  if (!st) {
    std::fprintf(stderr, "open: %s\n", st.get_err_msg().c_str());
    return 1;
  }
  for (int round = 0; round < 2; ++round) {
    unsigned char mine[64];
    for (int i = 0; i < 64; ++i) mine[i] = static_cast<unsigned char>(cfg.rank * 64 + i + round);
    std::vector<unsigned char> all(64u * cfg.world, 0);
    st = r.all_gather(mine, sizeof(mine), all.data());
    if (!st) {
      std::fprintf(stderr, "all_gather: %s\n", st.get_err_msg().c_str());
      return 1;
    }
    for (int q = 0; q < cfg.world; ++q)
      for (int i = 0; i < 64; ++i)
        if (all[q * 64 + i] != static_cast<unsigned char>(q * 64 + i + round)) {
          std::fprintf(stderr, "rank %d: byte %d of rank %d is wrong in round %d\n", cfg.rank, i, q, round);
          return 1;
        }
    st = r.barrier();
    if (!st) return 1;
  }
  std::printf("rank %d of %d: ok\n", cfg.rank, cfg.world);
  return 0;
}

int do_load(char** a) {
  const std::string path = a[0], family = a[1], prec = a[2];
  model::TpConfig cfg;
  cfg.world = std::atoi(a[3]), cfg.rank = std::atoi(a[4]);
  const bool quant = prec == "int8";
  std::unique_ptr<model::LLama2Model> m;
  if (family == "qwen")
    m = std::make_unique<model::Qwen2Model>(base::TokenizerType::kEncodeBpe, "<none>", path, quant);
  else
    m = std::make_unique<model::LLama2Model>(base::TokenizerType::kEncodeSpe, "<none>", path, quant);
  m->set_tensor_parallel(cfg);
  using I = model::ModelInspector;
  const base::Status st = I::load(*m);
  if (!st) {
    std::printf("error %s\n", st.get_err_msg().c_str());
    return 3;
  }
  const auto& ly = I::layers(*m);
  auto dump = [&](const char* name, const std::vector<std::shared_ptr<op::Layer>>& group) {
    for (size_t l = 0; l < group.size(); ++l) {
      auto mm = std::static_pointer_cast<op::MatmulLayer>(group[l]);
      const tensor::Tensor& w = mm->get_weight(0);
      std::printf("%s %zu dims %d %d w %016llx", name, l, w.get_dim(0), w.get_dim(1),
                  static_cast<unsigned long long>(fnv1a(w.ptr<int8_t>(), w.byte_size())));
      if (quant) {
        const tensor::Tensor& s = mm->get_scales();
        std::printf(" s %zu %016llx", s.size(), static_cast<unsigned long long>(fnv1a(s.ptr<float>(), s.byte_size())));
      }
      if (family == "qwen" && !quant && (name[1] == 'q' || name[1] == 'k' || name[1] == 'v')) {
        const tensor::Tensor& b = mm->get_bias(0);
        std::printf(" b %zu %016llx", b.size(), static_cast<unsigned long long>(fnv1a(b.ptr<float>(), b.byte_size())));
      }
      std::printf("\n");
    }
  };
  dump("wq", ly.wq_layers_), dump("wk", ly.wk_layers_), dump("wv", ly.wv_layers_), dump("wo", ly.wo_layers_);
  dump("w1", ly.w1_layers_), dump("w2", ly.w2_layers_), dump("w3", ly.w3_layers_);
  const auto& sh = m->tensor_parallel_shard();
  std::printf("local heads %d kv_heads %d hidden %d\n", sh.head_num, sh.kv_head_num, sh.hidden_dim);
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc >= 11 && !std::strcmp(argv[1], "shard")) return do_shard(argv + 2);
  if (argc >= 5 && !std::strcmp(argv[1], "rendezvous")) return do_rendezvous(argv + 2);
  if (argc >= 7 && !std::strcmp(argv[1], "load")) return do_load(argv + 2);
  std::fprintf(stderr, "usage: %s shard|rendezvous|load ... (see the header of tools/kuiper_tp_check.cpp)\n", argv[0]);
  return 2;
}
