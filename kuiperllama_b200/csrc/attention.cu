// Single-query (decode) attention over the fp32 KV cache: scores -> softmax -> weighted sum of
// values, one CTA per query head, GQA through head / kv_mul.
// Replaces multi_head_attention_kernel + softmax_gpu
// (kuiper/source/op/kernels/cuda/mha_kernel.cu:7-130).
//
// Arithmetic follows the reference kernel operation for operation so the output is
// bit-identical: each score is one left-to-right FFMA chain over head_size, softmax sums are
// taken by 256 strided lanes folded with the cub block-reduce tree, and out[i] is a single
// FFMA chain over t = 0..pos.  What changes is the memory side: q is staged in shared memory,
// value rows are streamed through a double-buffered shared-memory tile by all 256 threads
// (coalesced 128-bit loads) while the head_size chain-owning threads consume them, so the
// serial chain runs at shared-memory latency instead of L2 latency.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "../../include/kllm_b200.h"
#include "kllm_device.cuh"
#include "kllm_host.h"

namespace kllm {

constexpr int kMhaThreads = 256;
constexpr int kVTile = 32;  // timesteps per staged value tile

// cub::BlockReduce<float,256>::Sum (warp-reductions algorithm): per-warp shuffle tree, then
// thread 0 adds the 8 warp aggregates left to right.  Returns the total in thread 0 only.
__device__ __forceinline__ float block256_sum(float v, float* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_tree_sum(v);
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  float total = 0.f;
  if (threadIdx.x == 0) {
    total = s_warp[0];
#pragma unroll
    for (int w = 1; w < kMhaThreads / 32; ++w) total = __fadd_rn(total, s_warp[w]);
  }
  return total;
}

__device__ __forceinline__ float block256_max(float v, float* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, off));
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  float m = s_warp[0];
#pragma unroll
  for (int w = 1; w < kMhaThreads / 32; ++w) m = fmaxf(m, s_warp[w]);
  return m;  // every thread
}

__global__ void __launch_bounds__(kMhaThreads)
mha_decode_kernel(PosArg pos_arg, int seq_len, const float* __restrict__ query, float* score_ptr,
                  float* output, const float* __restrict__ key_cache,
                  const float* __restrict__ value_cache, int kv_dim, int kv_mul, int head_size,
                  long long layer_offset) {
  extern __shared__ __align__(16) float smem[];
  float* q_s = smem;                           // [head_size]
  float* v_s = smem + head_size;               // [2][kVTile][head_size]
  __shared__ float s_warp[kMhaThreads / 32];
  __shared__ float s_bcast;

  const int head = blockIdx.x;
  const int tid = threadIdx.x;
  const int pos = pos_arg.get();
  const float scale = 1.f / sqrtf(static_cast<float>(head_size));
  const float* query_head = query + static_cast<size_t>(head) * head_size;
  float* score_head = score_ptr + static_cast<size_t>(head) * seq_len;
  const int head_offset = (head / kv_mul) * head_size;
  const float* kbase = key_cache + layer_offset + head_offset;
  const float* vbase = value_cache + layer_offset + head_offset;

  for (int i = tid; i < head_size; i += kMhaThreads) q_s[i] = query_head[i];
  __syncthreads();

  // ---- scores: mha_kernel.cu:61-91 --------------------------------------------------
  const float4* q4 = reinterpret_cast<const float4*>(q_s);
  for (int t = tid; t <= pos; t += kMhaThreads) {
    const float4* k4 = reinterpret_cast<const float4*>(kbase + static_cast<size_t>(t) * kv_dim);
    float score = 0.0f;
#pragma unroll 4
    for (int i = 0; i < (head_size >> 2); ++i) {
      const float4 kv = k4[i];
      const float4 qv = q4[i];
      score = __fmaf_rn(kv.x, qv.x, score);
      score = __fmaf_rn(kv.y, qv.y, score);
      score = __fmaf_rn(kv.z, qv.z, score);
      score = __fmaf_rn(kv.w, qv.w, score);
    }
    score_head[t] = __fmul_rn(score, scale);
  }
  __syncthreads();

  // ---- softmax: mha_kernel.cu:7-45 ---------------------------------------------------
  const int size = pos + 1;
  float max_val = tid < size ? score_head[tid] : -FLT_MAX;
  for (int i = tid + kMhaThreads; i < size; i += kMhaThreads) max_val = fmaxf(max_val, score_head[i]);
  max_val = block256_max(max_val, s_warp);
  __syncthreads();

  float sum = 0.0f;
  for (int i = tid; i < size; i += kMhaThreads) {
    const float e = expf(score_head[i] - max_val);
    score_head[i] = e;
    sum += e;
  }
  sum = block256_sum(sum, s_warp);
  if (tid == 0) s_bcast = sum;
  __syncthreads();
  sum = s_bcast;
  for (int i = tid; i < size; i += kMhaThreads) score_head[i] = score_head[i] / sum;
  __syncthreads();

  // ---- weighted value sum: mha_kernel.cu:97-109 ----------------------------------------
  // All threads stage value tiles; threads < head_size own one output chain each.
  const int vec_per_row = head_size >> 2;
  const int n_tiles = (size + kVTile - 1) / kVTile;
  auto stage = [&](int tile, int buf) {
    const int t0 = tile * kVTile;
    float4* dst = reinterpret_cast<float4*>(v_s + static_cast<size_t>(buf) * kVTile * head_size);
    for (int e = tid; e < kVTile * vec_per_row; e += kMhaThreads) {
      const int tt = e / vec_per_row, c = e % vec_per_row;
      if (t0 + tt <= pos)
        dst[e] = *reinterpret_cast<const float4*>(vbase + static_cast<size_t>(t0 + tt) * kv_dim + 4 * c);
    }
  };
  float value = 0.0f;
  stage(0, 0);
  __syncthreads();
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < n_tiles) stage(tile + 1, buf ^ 1);
    if (tid < head_size) {
      const float* vt = v_s + static_cast<size_t>(buf) * kVTile * head_size + tid;
      const int t0 = tile * kVTile;
      const int cnt = min(kVTile, size - t0);
#pragma unroll 8
      for (int tt = 0; tt < cnt; ++tt)
        value = __fmaf_rn(score_head[t0 + tt], vt[tt * head_size], value);
    }
    __syncthreads();
  }
  if (tid < head_size) output[static_cast<size_t>(head) * head_size + tid] = value;
}

int launch_mha(PosArg pos, int head_num, int layer_index, int seq_len, int kv_dim, int kv_mul,
               int head_size, float* mha_out, const float* query, float* score,
               const float* key_cache, const float* value_cache, cudaStream_t stream) {
  if (!mha_out || !query || !score || !key_cache || !value_cache) return KLLM_E_INVALID;
  if (head_num <= 0 || kv_mul <= 0 || head_size <= 0 || layer_index < 0 || seq_len <= 0)
    return KLLM_E_INVALID;
  if ((head_size & 3) != 0 || (kv_dim & 3) != 0 || head_size > kMhaThreads)
    return KLLM_E_UNSUPPORTED;
  const long long layer_offset = static_cast<long long>(layer_index) * seq_len * kv_dim;
  const size_t smem = sizeof(float) * (head_size + 2 * kVTile * head_size);
  mha_decode_kernel<<<head_num, kMhaThreads, smem, stream>>>(pos, seq_len, query, score, mha_out,
                                                            key_cache, value_cache, kv_dim,
                                                            kv_mul, head_size, layer_offset);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace kllm

extern "C" int kllm_mha_decode_f32(int pos, int head_num, int layer_index, int seq_len, int kv_dim,
                                   int kv_mul, int head_size, float* mha_out, const float* query,
                                   float* score, const float* key_cache, const float* value_cache,
                                   void* stream) {
  if (pos < 0 || pos >= seq_len) return KLLM_E_INVALID;
  return kllm::launch_mha(kllm::PosArg{nullptr, pos}, head_num, layer_index, seq_len, kv_dim,
                          kv_mul, head_size, mha_out, query, score, key_cache, value_cache,
                          static_cast<cudaStream_t>(stream));
}
