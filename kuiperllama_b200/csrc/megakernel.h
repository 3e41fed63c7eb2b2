// Host/device shared declarations of the persistent decode megakernel (megakernel.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace kllm {
namespace mega {

enum {
  kPhaseGemv = 0,
  kPhaseAttention = 1 /* scores of the split attention */,
  kPhaseAttnPV = 2 /* softmax + P.V of the split attention */,
  kPhaseAttnFused = 3 /* whole attention of one head in one CTA (attn_split == 1) */,
  kPhaseAttnFlash = 4 /* toleranced: split by timestep, online softmax, partials merged by CTA 0 of the head */,
  kPhaseGather = 5 /* tensor parallel, classifier sharded by vocabulary: collect every rank's logits, argmax partials */
};
constexpr int kProfStamps = 16;  // uint64 stamps per (CTA, phase) of kllm_decoder_profile

struct Seg {
  const void* w;        // fp32 or int8 [rows, in_dim]
  const float* scales;  // int8 only
  const float* bias;    // optional
  float* out;           // out[pos * pos_stride + row]
  long long pos_stride;
  int rows;
  int head_major;       // 1: out is the head-major value cache, index ((row/hs)*seq_len + pos)*hs + row%hs
  // optional local hand-off: row r is (also) published as a tagged 64-bit word at tag_out[r] for
  // the next phase to poll (out may then be null)
  unsigned long long* tag_out;
};

// One entry of the per-token schedule.  A GEMV phase's units (rows, or w1/w3 row pairs) are
// split evenly over the CTAs; each CTA streams its contiguous share through the stage ring.
struct Phase {
  int kind;
  int in_dim;
  int units;
  int n_seg;
  int swiglu;             // units are (w1 row, w3 row) pairs -> SiLU*gate epilogue
  int argmax;             // track (max, index) of the produced rows (classifier)
  int cls;                // classifier work: skipped for prompt positions (llama3.cpp:733-745 discards their logits)
  int x_from_emb;         // input vector is the embedding row of the current token
  int residual_from_emb;  // residual source is the embedding row (layer 0)
  int group_size, group_shift;
  int rows_per_stage;     // whole rows per ring stage (chunks_per_row == 1)
  int task_rows;          // rows handed to one consumer warp at a time (1, 2 or 4; picked per phase at init)
  int mma;                // int8 fast mode: the rows of a stage go through mma.sync m16n8k32 s8
  int team;               // int8 fast mode: a TEAM of consumer warps shares each ring stage, splitting the columns
  int row_pad;            // ... and are staged row_pad bytes apart more than their length (bank-conflict-free fragments)
  int chunks_per_row;     // > 1: a row spans this many stages (fp32 rows longer than a stage)
  int chunk_elems;
  int scale_off;          // byte offset of the scales region inside a stage (int8)
  int scale_row_bytes;
  int layer;              // attention: layer index
  float norm_eps;
  const float* x;         // input vector (global memory)
  const float* norm_w;    // optional RMSNorm weight applied to x
  const float* residual;  // optional residual vector
  // Tagged exchange (replaces the grid barrier AND, under tensor parallelism, the all-reduce):
  //   tp_out: each produced row is published as {value, tag} -- one 64-bit store -- into every
  //           rank's exchange area instead of seg[0].out; no residual add, no barrier after.
  //   tp_in : the input vector is x_old + (p_0 + p_1 + ... + p_{world-1}), p_r = rank r's tagged
  //           partials of exchange `exch`, polled until their tag is current; x_old is the CTA's
  //           own shared-memory copy of the residual stream (it starts as the embedding row of the
  //           token and is updated by every exchange), so the stream never round-trips through
  //           global memory between CTAs.
  int tp_in, tp_out;
  int exch;      // index (within the token) of the exchange tp_in consumes
  int exch_out;  // ... of the exchange tp_out publishes (a phase may do both: the sharded classifier)
  int barrier_after;      // 1: a grid barrier closes the phase
  int barrier_idx;        // barriers of this token passed once this phase is closed
  // Local tagged hand-offs (same words, one rank): the phase's input vector is polled from tag_in
  // (GEMV) or tq/tk/tv (attention: this head's query, its kv head's raw key and value rows);
  // outputs go to seg[].tag_out (GEMV) or ta (attention).  hand_in / hand_out number the
  // hand-offs of a token for the tags.
  const unsigned long long* tag_in;
  const unsigned long long* tq;
  const unsigned long long* tk;
  const unsigned long long* tv;
  unsigned long long* ta;
  int hand_in, hand_out;
  int hand_aux;           // attention P.V phase: the q|k|v hand-off (value row of the current position)
  Seg seg[3];
};

struct State {  // == StepState in decoder.cu
  int32_t token, pos, step, next;
};

struct Params {
  const Phase* phases;
  int n_phases, n_tokens;
  int attn_vsplit;      // V cache layout [L][kv_head][attn_vsplit][seq_len][head_size / attn_vsplit]
  int attn_parts;       // flash attention: threads per timestep in the scores pass
  int int8_fast;        // int8 weights: 1 = fixed-point activations on dp4a (toleranced), 0 = the reference's per-element order
  int n_cls_phases;     // trailing phases that make up the classifier (1, or 2 with the vocabulary-sharded form)
  int skip_cls_tokens;  // the first skip_cls_tokens positions of this launch are prompt tokens: no classifier pass
  int num_stages, stage_bytes, xbuf_bytes;
  int xres_bytes;  // shared-memory copy of the residual stream behind the input vector (tagged modes; else 0)
  int attn_tile;    // timesteps per K ring stage
  int attn_tile_v;  // timesteps per V ring stage (rows of head_size / attn_split floats)
  int attn_split;   // CTAs per query head: K tiles round-robin, P.V output dims split (bit-exact chains)
  unsigned long long* scores;  // [head][seq_len] tagged scaled scores: scores phase -> P.V phase
  int pf_stages;  // L2 prefetch run-ahead of the ring producer, in ring stages (0 = off)
  int group_size;
  int dim, vocab_size, head_num, head_size, kv_dim, kv_mul, seq_len, flavour;
  const float* tok_emb;
  const float* q;
  const float* k_raw;
  float* attn_out;
  float* score;
  // KV cache in the persistent engine's own layout (see megakernel.cu "KV layout"):
  //   K [L][kv_head][head_size/4][seq_len][4]    V [L][kv_head][attn_split][seq_len][head_size/attn_split]
  float* key_cache;
  const float* value_cache;
  const float* sin_cache;
  const float* cos_cache;
  State* state;
  int32_t* out_tokens;
  const int32_t* teacher;
  int max_steps;
  unsigned* barrier;
  unsigned barrier_base;
  int bars_per_token;
  // tagged exchange areas: tp_data[r] = rank r's area [2 slots][tp_world][tp_stride] of 64-bit
  // {tag:32 | fp32 bits:32}; tp_data[tp_rank] is local memory, the others NVLink peer mappings
  unsigned long long* tp_data[8];
  int tp_world, tp_rank, tp_stride, exch_per_token;
  unsigned tp_seq_base;
  unsigned hand_base;
  int hands_per_token;
  float* arg_val;
  int* arg_idx;
  // optional phase timeline of one token: prof[(cta * n_phases + phase) * 4 + k], k = phase
  // entered / input staged / last stage consumed / grid barrier passed (globaltimer ns)
  unsigned long long* prof;
  int prof_token;
};

}  // namespace mega

// Everything the engine needs to know about the model; all pointers are device pointers, the
// per-layer arrays are host arrays of device pointers (owned by the decoder).
struct MegaModel {
  int dim, hidden_dim, layer_num, head_num, kv_head_num, vocab_size, seq_len;
  int head_size, kv_dim, kv_mul, flavour, group_size;
  const float* tok_emb;
  const float* const* attn_norm;
  const float* const* ffn_norm;
  const float* final_norm;
  const void* const* wq; const void* const* wk; const void* const* wv; const void* const* wo;
  const void* const* w1; const void* const* w2; const void* const* w3;
  const void* wcls;
  const float* const* sq; const float* const* sk; const float* const* sv; const float* const* so;
  const float* const* s1; const float* const* s2; const float* const* s3;
  const float* scls;
  const float* const* bq; const float* const* bk; const float* const* bv;
  // activations / state owned by the decoder
  float* x; float* q; float* k_raw; float* attn_out; float* h; float* logits; float* score;
  float* key_cache; float* value_cache;
  const float* sin_cache; const float* cos_cache;
  void* state;
  int32_t* out_tokens;
  // tensor parallel (tp_world > 1): exchange areas of every rank (kllm_comm, CUDA IPC)
  int tp_world, tp_rank;
  unsigned long long* tp_data[8];
  int tp_stride;
  int numerics;  // kllm_decoder_desc::numerics
};

class MegaEngine {
 public:
  int init(const MegaModel& m, cudaStream_t stream);
  void destroy();
  // Run n_tokens consecutive positions starting from the device-resident state.
  int run(int n_tokens, const int32_t* teacher_dev, unsigned long long* prof_dev = nullptr,
          int prof_token = -1, int skip_cls_tokens = 0);
  int grid() const { return grid_; }
  bool ready() const { return ready_; }
  int stages() const { return stages_; }
  int stage_bytes() const { return stage_bytes_; }
  int phases() const { return n_phases_; }
  int attn_tile() const { return attn_tile_; }
  int attn_split() const { return attn_split_; }    // CTAs per query head
  int attn_vsplit() const { return attn_vsplit_; }  // slices of the V cache layout
  bool fast() const { return fast_ != 0; }
  int cls_rows() const { return cls_rows_; }  // classifier rows this rank streams per token
  int consumer_warps() const { return consumer_warps_; }
  bool int8_fast() const { return int8_fast_ != 0; }

 private:
  MegaModel model_{};
  cudaStream_t stream_ = nullptr;
  void* d_phases_ = nullptr;
  void* d_barrier_ = nullptr;
  void* d_arg_val_ = nullptr;
  void* d_arg_idx_ = nullptr;
  unsigned long long* d_tagged_ = nullptr;  // single-GPU exchange area (tp_world == 1)
  unsigned long long* d_handoff_ = nullptr;  // local tagged hand-off vectors (q | k | v | attn | h)
  bool tagged_ = false;
  int tagged_mode_ = 0;  // 0: grid barriers everywhere; 1: tagged residual exchange; 2: + tagged hand-offs
  int exch_per_token_ = 0, hands_per_token_ = 0;
  unsigned tp_seq_base_ = 0, hand_base_ = 0;
  int grid_ = 0, stages_ = 0, stage_bytes_ = 0, xbuf_bytes_ = 0, xres_bytes_ = 0, n_phases_ = 0, attn_tile_ = 0;
  int consumer_warps_ = 8, threads_ = 0;
  int attn_split_ = 1, attn_tile_v_ = 0;
  unsigned long long* d_scores_ = nullptr;  // tagged scores of the split attention
  int int8_fast_ = 0;
  int fast_ = 0;  // numerics: 0 = bit-exact with the reference, 1 = toleranced (free summation order)
  int attn_vsplit_ = 1, attn_parts_ = 1;
  int cls_rows_ = 0, n_cls_phases_ = 1;
  const void* kernel_ = nullptr;       // decode_megakernel<consumer warps, int8, false>
  const void* kernel_prof_ = nullptr;  // ... <.., true>: records the phase timeline stamps
  int n_barriers_per_token_ = 0;
  size_t smem_bytes_ = 0;
  unsigned barrier_base_ = 0;
  bool ready_ = false;
};

}  // namespace kllm
