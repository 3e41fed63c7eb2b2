// Tensor parallelism for the C++ host side (not in the reference, which pins
// device 0: kuiper/source/model/llama3.cpp:118).  One PROCESS per GPU, no torch, no MPI:
//
//   TpConfig       who this process is (world, rank, CUDA device) and where the ranks meet
//                  (a TCP port on 127.0.0.1).  from_env() reads KUIPER_TP_WORLD / KUIPER_TP_RANK /
//                  KUIPER_TP_DEVICE / KUIPER_TP_ADDR / KUIPER_TP_PORT, which is what
//                  tools/kuiper_tp_launch sets for each child, so an UNCHANGED demo/main.cpp runs
//                  sharded when started through the launcher.
//   TpShard        the slice of every weight matrix a rank owns -- the same rules as
//                  kuiperllama_b200/tensor_parallel.py (heads / kv heads / FFN rows; int8 FFN shards
//                  in units of 256 columns) so both host sides build identical shards.
//   TpRendezvous   rank 0 listens, the others connect; all_gather() of fixed-size blobs (the 64-byte
//                  CUDA-IPC handles of kllm_comm) and barrier().  Control plane only: the data plane is
//                  the tagged peer-memory exchange inside the persistent kernel (csrc/megakernel.cu).
//
// LLama2Model::init() cuts the shards straight out of the mmap'd checkpoint at load time: the
// row slices of wq / wk / wv / w1 / w3 are contiguous spans of the file and are uploaded as such; the
// column slices of wo / w2 (and their int8 group scales) are packed once into host staging buffers,
// uploaded through the pinned double-buffered uploader and freed.
#ifndef KLLM_KUIPER_MODEL_TENSOR_PARALLEL_H_
#define KLLM_KUIPER_MODEL_TENSOR_PARALLEL_H_
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "base/base.h"
#include "model/checkpoint_file.h"

namespace model {
struct TpConfig {
  int world = 1;
  int rank = 0;
  int device = -1;  // CUDA device ordinal; -1 = rank (0 on a single GPU)
  std::string addr = "127.0.0.1";
  int port = 29641;

  bool on() const { return world > 1; }
  int cuda_device() const { return device >= 0 ? device : (on() ? rank : 0); }
  static TpConfig from_env();
};

struct TpShard {
  int32_t q0 = 0, q1 = 0;  // rows of wq == columns of wo
  int32_t k0 = 0, k1 = 0;  // rows of wk / wv
  int32_t f0 = 0, f1 = 0;  // rows of w1 / w3 == columns of w2
  int32_t head_num = 0, kv_head_num = 0, hidden_dim = 0;  // LOCAL counts (kllm_decoder_desc under tp)
};
// int8 FFN shards are cut in units of this many columns: whole quantisation groups (64) and 16-byte
// rows of group scales per shard -- what the persistent engine's TMA ring can stage
constexpr int32_t kTpInt8FfnUnit = 256;

// The slice rank `rank` of `world` owns, or an error Status saying why the model does not split.
base::Status tp_shard(const TransformerConfig& c, int32_t group_size, int world, int rank, TpShard* out);
// floats a rank's exchange area must carry: dim for the residual exchange, vocab / world more lets the
// persistent engine shard the classifier by vocabulary (kllm_comm_create max_count)
int32_t tp_comm_words(const TransformerConfig& c, int world);

class TpRendezvous {
 public:
  TpRendezvous() = default;
  ~TpRendezvous();
  TpRendezvous(const TpRendezvous&) = delete;
  TpRendezvous& operator=(const TpRendezvous&) = delete;

  // rank 0 binds and accepts world - 1 peers; the others connect (retrying for timeout_s seconds)
  base::Status open(const TpConfig& cfg, int timeout_s = 120);
  // every rank contributes `bytes` bytes; all[r * bytes ...] = rank r's contribution, on every rank
  base::Status all_gather(const void* mine, size_t bytes, void* all);
  base::Status barrier();
  void close();
  bool is_open() const { return open_; }

 private:
  int world_ = 1, rank_ = 0;
  bool open_ = false;
  int listen_fd_ = -1;
  std::vector<int> peers_;  // rank 0: fd of rank r at [r]; others: [0] = rank 0
};
}  // namespace model
#endif  // KLLM_KUIPER_MODEL_TENSOR_PARALLEL_H_
