// Persistent decode megakernel for sm_100a: the whole per-token forward of LLama2Model /
// Qwen2Model (kuiper/source/model/llama3.cpp:147-167, 600-745) in ONE cooperative launch that
// can run any number of consecutive positions.
//
// Why: at batch 1 the path is a 4-26 GB/token weight stream; with one launch per op (even 6
// fused launches per layer in a CUDA graph) every kernel boundary drains the HBM pipeline
// (profiles/r01a: 15-42 % DRAM utilisation per kernel).  Here one CTA per SM stays resident
// and a dedicated producer warp streams that CTA's share of EVERY weight matrix -- and of the
// KV cache -- in schedule order through a ring of shared-memory stages with TMA bulk copies
// (cp.async.bulk -> UBLKCP) signalled on mbarriers.  Weights and past KV rows never depend on
// the current token's activations, so the producer runs ahead across phase boundaries, grid
// barriers and the attention phase: HBM keeps streaming while consumers wait for each other.
//
// Consumers: 8 warps.  Every ring stage is OWNED by one warp (round robin), which waits for it,
// processes all its rows in groups of up to 4 that share the input-vector loads (the consumers
// are bound by the 128 B/cycle shared-memory pipe, so traffic per weight byte matters), and
// releases it.  Schedule per layer (5 grid barriers):  QKV(+bias) | attention(+RoPE) |
// Wo+residual | W1,W3->SiLU*gate | W2+residual ; then classifier + greedy argmax.
//
// Arithmetic is the same as the per-op kernels (gemv.cu / attention.cu / elementwise.cu): every
// dot product, reduction tree, softmax sum and value chain reproduces the reference CUDA
// kernels' floating-point order, so logits stay bit-identical to the reference's CUDA path.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/kllm_b200.h"
#include "kllm_device.cuh"
#include "kllm_host.h"
#include "megakernel.h"
#include "megakernel_device.cuh"

namespace kllm {
namespace mega {

// ---- attention phase (mha_kernel.cu:47-110 + rope_kernel.cu), two steps inside one schedule phase
//
// Per-SM TMA ingest is ~47 GB/s, so a head-per-CTA mapping (8 query heads of a GQA group each
// re-reading the same 2 x pos x 256 B of KV on 32 SMs) is ingest-bound; instead the KV bytes are
// read once and spread over every SM:
//   step A (all CTAs): item (kv group g, K tile j) -> scores of ALL query heads of the group
//       over the tile's timesteps (one FFMA chain per (head, t), reference order), written to the
//       global [head][seq_len] score buffer; item (g, n_kt) = the current position: rotate the
//       new key, store it, score it.  Each finished item bumps flags[g] (release).
//   step B (few CTAs): item (g, value slab, head chunk): wait for flags[g] (acquire), one warp
//       per head does the reference softmax (256 virtual lanes folded in cub order) into shared
//       memory, then walks its FFMA chain over the slab's V tiles; all warps share the tiles.
// No grid barrier between A and B: only the CTAs that own a B item wait, on their group's flag.
struct AttnPlan {
  int n_kt, n_vt, items_a, hpc, nchunks, items_b;
};
__device__ __forceinline__ AttnPlan attn_plan(const Params& P, int pos) {
  AttnPlan a;
  const AttnGeom g = attn_geom(P);
  const int kv_heads = P.kv_dim / P.head_size;
  a.n_kt = ceil_div(pos, g.tk);
  a.n_vt = ceil_div(pos, g.tv);
  a.items_a = kv_heads * (a.n_kt + 1);
  const int cap = P.xbuf_bytes >> 2;
  const int need = (pos + 1 + 3) & ~3;
  int hpc = cap / need;  // heads whose probabilities fit the workspace
  if (hpc > P.kv_mul) hpc = P.kv_mul;
  if (hpc > kNW) hpc = kNW;
  if (hpc < 1) hpc = 1;  // probabilities live in the global score buffer instead
  a.hpc = hpc;
  a.nchunks = ceil_div(P.kv_mul, hpc);
  a.items_b = kv_heads * g.slabs * a.nchunks;
  return a;
}

__device__ void attention_phase(const Params& P, const Phase& ph, int cta, int G, int pos, float* ws,
                                unsigned char* stages, uint64_t* full_bar, uint64_t* empty_bar,
                                Ring& ring, unsigned& flag_base, unsigned long long* stamp) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int hs = P.head_size, seq_len = P.seq_len, S = P.num_stages, kv_mul = P.kv_mul;
  const AttnGeom g = attn_geom(P);
  const AttnPlan pl = attn_plan(P, pos);
  const int kv_heads = P.kv_dim / hs;
  const float scale = 1.f / sqrtf(static_cast<float>(hs));
  float* q_s = ws;                 // [kv_mul][hs] rotated queries of one group
  float* k_s = ws + kv_mul * hs;   // [hs] rotated key of the current position

  // ================= step A: scores =================
  for (int item = cta; item < pl.items_a; item += G) {
    const int grp = item / (pl.n_kt + 1);
    const int j = item % (pl.n_kt + 1);
    const size_t head_block = (static_cast<size_t>(ph.layer) * kv_heads + grp) * seq_len * hs;
    float* kcache = P.key_cache + head_block;
    consumer_sync();  // previous item done with q_s / k_s
    // RoPE of the group's queries (and, for the current-position item, of the new key row):
    // rope_kernel.cu as compiled -- x' = fma(cos, x0, -(sin*x1)), y' = fma(sin, x0, cos*x1)
    const int half = hs >> 1;
    for (int idx = tid; idx < (kv_mul + 1) * half; idx += kConsumerThreads) {
      const int hh = idx / half, pr = idx % half;
      if (hh == kv_mul && j != pl.n_kt) continue;
      int i0, i1;
      if (P.flavour == KLLM_FLAVOUR_LLAMA2) {
        i0 = 2 * pr, i1 = 2 * pr + 1;
      } else {
        i0 = pr, i1 = pr + half;
      }
      const float fci = P.sin_cache[static_cast<size_t>(pos) * hs + 2 * pr];
      const float fcr = P.cos_cache[static_cast<size_t>(pos) * hs + 2 * pr];
      if (hh < kv_mul) {
        const float* qg = P.q + static_cast<size_t>(grp * kv_mul + hh) * hs;
        const float x0 = __ldcg(qg + i0), x1 = __ldcg(qg + i1);
        q_s[hh * hs + i0] = __fmaf_rn(fcr, x0, -__fmul_rn(fci, x1));
        q_s[hh * hs + i1] = __fmaf_rn(fci, x0, __fmul_rn(fcr, x1));
      } else {
        const float* kg = P.k_raw + grp * hs;
        const float x0 = __ldcg(kg + i0), x1 = __ldcg(kg + i1);
        const float r0 = __fmaf_rn(fcr, x0, -__fmul_rn(fci, x1));
        const float r1 = __fmaf_rn(fci, x0, __fmul_rn(fcr, x1));
        k_s[i0] = r0;
        k_s[i1] = r1;
        kcache[(static_cast<size_t>(i0 >> 2) * seq_len + pos) * 4 + (i0 & 3)] = r0;
        kcache[(static_cast<size_t>(i1 >> 2) * seq_len + pos) * 4 + (i1 & 3)] = r1;
      }
    }
    consumer_sync();
    if (j < pl.n_kt) {
      // K tile: warp w takes timesteps [w*tpw, (w+1)*tpw) of the tile, lane = timestep; the
      // timestep's key row sits in registers while the group's heads are walked.
      const int t0 = j * g.tk;
      const int nt = min(g.tk, pos - t0);
      const int tpw = g.tk / kNW;
      const int tl = warp * tpw + lane;
      mbar_wait(&full_bar[ring.slot], ring.parity);
      if (lane < tpw && tl < nt) {
        const float4* tile = reinterpret_cast<const float4*>(stages + static_cast<size_t>(ring.slot) * P.stage_bytes);
        for (int hh = 0; hh < kv_mul; ++hh) {
          const float4* q4 = reinterpret_cast<const float4*>(q_s + hh * hs);
          float sc = 0.0f;
#pragma unroll 4
          for (int c = 0; c < (hs >> 2); ++c) {
            const float4 kv = tile[c * g.tk + tl];
            const float4 qv = q4[c];
            sc = __fmaf_rn(kv.x, qv.x, sc);
            sc = __fmaf_rn(kv.y, qv.y, sc);
            sc = __fmaf_rn(kv.z, qv.z, sc);
            sc = __fmaf_rn(kv.w, qv.w, sc);
          }
          P.score[static_cast<size_t>(grp * kv_mul + hh) * seq_len + t0 + tl] = __fmul_rn(sc, scale);
        }
      }
      consumer_sync();
      if (tid == 0) mbar_arrive(&empty_bar[ring.slot]);
      ring.advance(S);
    } else {
      // current position: q . rotated new key, one thread per head
      if (tid < kv_mul) {
        const float4* q4 = reinterpret_cast<const float4*>(q_s + tid * hs);
        const float4* k4 = reinterpret_cast<const float4*>(k_s);
        float sc = 0.0f;
        for (int c = 0; c < (hs >> 2); ++c) {
          const float4 kv = k4[c];
          const float4 qv = q4[c];
          sc = __fmaf_rn(kv.x, qv.x, sc);
          sc = __fmaf_rn(kv.y, qv.y, sc);
          sc = __fmaf_rn(kv.z, qv.z, sc);
          sc = __fmaf_rn(kv.w, qv.w, sc);
        }
        P.score[static_cast<size_t>(grp * kv_mul + tid) * seq_len + pos] = __fmul_rn(sc, scale);
      }
      consumer_sync();
    }
    if (tid == 0) red_release_add(P.attn_flags + grp, 1u);  // orders the CTA's score stores (bar.sync above)
  }
  if (stamp) stamp[1] = stamp[7] = global_ns();
  const unsigned flag_need = flag_base + static_cast<unsigned>(pl.n_kt + 1);
  flag_base = flag_need;

  // ================= step B: softmax + weighted values =================
  const int need = (pos + 1 + 3) & ~3;
  const bool p_in_smem = static_cast<long long>(pl.hpc) * need <= (P.xbuf_bytes >> 2);
  for (int item = cta; item < pl.items_b; item += G) {
    const int grp = item / (g.slabs * pl.nchunks);
    const int sl = (item / pl.nchunks) % g.slabs;
    const int ck = item % pl.nchunks;
    const int h_local = ck * pl.hpc + warp;
    const bool active = warp < pl.hpc && h_local < kv_mul;
    const int head = grp * kv_mul + h_local;
    const size_t head_block = (static_cast<size_t>(ph.layer) * kv_heads + grp) * seq_len * hs;
    const float* vslab = P.value_cache + head_block + static_cast<size_t>(sl) * seq_len * g.sw;
    if (tid == 0) {
      while (static_cast<int>(ld_acquire_u32(P.attn_flags + grp) - flag_need) < 0) {
      }
    }
    consumer_sync();  // also: previous item done with the workspace
    float* prob = nullptr;
    float v_pos = 0.f;
    if (active) {
      float* srow = P.score + static_cast<size_t>(head) * seq_len;
      prob = p_in_smem ? (ws + static_cast<size_t>(warp) * need) : srow;
      if (lane < g.sw) v_pos = __ldcg(vslab + static_cast<size_t>(pos) * g.sw + lane);
      // ---- softmax, mha_kernel.cu:7-45: 256 strided lanes, cub BlockReduce order; this warp
      // carries virtual lane (l + 32 w) in acc[w].
      const int size = pos + 1;
      float mx = -FLT_MAX;
      for (int i = lane; i < size; i += 32) {
        const float v = __ldcg(srow + i);
        prob[i] = v;
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, off));
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int base = 0; base < size; base += 256) {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const int i = base + 32 * w + lane;
          if (i < size) {
            const float e = expf(prob[i] - mx);
            prob[i] = e;
            acc[w] += e;
          }
        }
      }
      float total = warp_tree_sum(acc[0]);
#pragma unroll
      for (int w = 1; w < 8; ++w) total = __fadd_rn(total, warp_tree_sum(acc[w]));
      total = __shfl_sync(kFull, total, 0);
      for (int i = lane; i < size; i += 32) prob[i] = prob[i] / total;
      __syncwarp();
    }
    if (stamp) stamp[7] = global_ns();
    // ---- weighted value sum, mha_kernel.cu:97-109: one FFMA chain per output element; the
    // CTA's warps (one head each) share the slab's V tiles.
    float value = 0.0f;
    for (int jv = 0; jv < pl.n_vt; ++jv) {
      const int t0 = jv * g.tv;
      const int nt = min(g.tv, pos - t0);
      mbar_wait(&full_bar[ring.slot], ring.parity);
      if (active && lane < g.sw) {
        const float* vt = reinterpret_cast<const float*>(stages + static_cast<size_t>(ring.slot) * P.stage_bytes) + lane;
        const float* pr = prob + t0;
        int tt = 0;
        if ((reinterpret_cast<uintptr_t>(pr) & 15) == 0) {
          const float4* pr4 = reinterpret_cast<const float4*>(pr);
#pragma unroll 2
          for (; tt + 4 <= nt; tt += 4) {
            const float4 p4 = pr4[tt >> 2];
            value = __fmaf_rn(p4.x, vt[(tt + 0) * g.sw], value);
            value = __fmaf_rn(p4.y, vt[(tt + 1) * g.sw], value);
            value = __fmaf_rn(p4.z, vt[(tt + 2) * g.sw], value);
            value = __fmaf_rn(p4.w, vt[(tt + 3) * g.sw], value);
          }
        }
        for (; tt < nt; ++tt) value = __fmaf_rn(pr[tt], vt[tt * g.sw], value);
      }
      consumer_sync();
      if (tid == 0) mbar_arrive(&empty_bar[ring.slot]);
      ring.advance(S);
    }
    if (active && lane < g.sw) {
      value = __fmaf_rn(prob[pos], v_pos, value);
      P.attn_out[static_cast<size_t>(head) * hs + sl * g.sw + lane] = value;
    }
  }
}

// ---- the kernel ---------------------------------------------------------------------------------
template <bool kInt8>
__global__ void __launch_bounds__(kThreads, 1) decode_megakernel(const Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ float s_warp[kNW];
  __shared__ float s_bcast;
  __shared__ float s_argv[kNW];
  __shared__ int s_argi[kNW];
  // The ring leaves only a few KB of L1, so everything the inner loops touch lives in shared
  // memory or registers: the consumer's and the producer's current schedule entries are copied
  // here (they are usually in different phases).
  __shared__ Phase s_phase_cons;
  __shared__ Phase s_phase_prod;

  float* xs = reinterpret_cast<float*>(smem);
  unsigned char* stages = smem + P.xbuf_bytes;
  const int S = P.num_stages;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const bool is_producer = warp == kNW;
  const int cta = blockIdx.x;
  const int G = gridDim.x;
  constexpr int wbytes = kInt8 ? 1 : 4;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);   // producer's expect_tx arrival (+ transaction bytes)
      mbar_init(&empty_bar[s], 1);  // the owning consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  Ring ring{0, 0u, 0};

  // =============================== producer warp ===============================================
  if (is_producer) {
    const uint64_t policy = policy_evict_first();   // weights: streamed once per token
    const uint64_t policy_kv = policy_evict_last();  // KV tiles: re-read every token, keep in L2
    int ppos = P.state->pos;
    for (int tok = 0; tok < P.n_tokens; ++tok, ++ppos) {
      for (int pi = 0; pi < P.n_phases; ++pi) {
        {
          const uint32_t* src = reinterpret_cast<const uint32_t*>(P.phases + pi);
          uint32_t* dst = reinterpret_cast<uint32_t*>(&s_phase_prod);
          __syncwarp();
          for (int i = lane; i < static_cast<int>(sizeof(Phase) / 4); i += 32) dst[i] = __ldg(src + i);
          __syncwarp();
        }
        const Phase& ph = s_phase_prod;
        unsigned long long* pstamp = (P.prof != nullptr && tok == P.prof_token && lane == 0)
                                         ? P.prof + (static_cast<size_t>(cta) * P.n_phases + pi) * 8
                                         : nullptr;
        if (pstamp) pstamp[4] = global_ns();
        if (ph.kind == kPhaseAttention) {
          const AttnPlan pl = attn_plan(P, ppos);
          if (ppos > 0) {
            // rows t < pos: final since the previous token.  Order the async-proxy reads after the
            // grid barrier that closed the previous token's attention phase.
            if (tok > 0 && lane == 0) {
              const unsigned need = P.barrier_base + static_cast<unsigned>((tok - 1) * P.n_phases + pi + 1) *
                                                         static_cast<unsigned>(G);
              while (static_cast<int>(ld_acquire_u32(P.barrier) - need) < 0) {
              }
              asm volatile("fence.proxy.async;" ::: "memory");
            }
            __syncwarp();
            const int hs = P.head_size;
            const AttnGeom g = attn_geom(P);
            const int kv_heads = P.kv_dim / hs;
            for (int item = cta; item < pl.items_a; item += G) {
              const int grp = item / (pl.n_kt + 1), j = item % (pl.n_kt + 1);
              if (j == pl.n_kt) continue;
              const float* kbase = P.key_cache + (static_cast<size_t>(ph.layer) * kv_heads + grp) * P.seq_len * hs;
              const int t0 = j * g.tk;
              const int nt = min(g.tk, ppos - t0);
              mbar_wait(&empty_bar[ring.slot], ring.parity ^ 1u);
              unsigned char* dst = stages + static_cast<size_t>(ring.slot) * P.stage_bytes;
              if (lane == 0) mbar_expect_tx(&full_bar[ring.slot], static_cast<uint32_t>(nt) * hs * 4);
              __syncwarp();
              if (lane < (hs >> 2))
                bulk_g2s(dst + static_cast<size_t>(lane) * g.tk * 16,
                         kbase + (static_cast<size_t>(lane) * P.seq_len + t0) * 4,
                         static_cast<uint32_t>(nt) * 16, &full_bar[ring.slot], policy_kv);
              ring.advance(S);
            }
            for (int item = cta; item < pl.items_b; item += G) {
              const int grp = item / (g.slabs * pl.nchunks);
              const int sl = (item / pl.nchunks) % g.slabs;
              const float* vslab = P.value_cache + (static_cast<size_t>(ph.layer) * kv_heads + grp) * P.seq_len * hs +
                                   static_cast<size_t>(sl) * P.seq_len * g.sw;
              for (int jv = 0; jv < pl.n_vt; ++jv) {
                const int t0 = jv * g.tv;
                const int nt = min(g.tv, ppos - t0);
                mbar_wait(&empty_bar[ring.slot], ring.parity ^ 1u);
                if (lane == 0) {
                  mbar_expect_tx(&full_bar[ring.slot], static_cast<uint32_t>(nt) * g.sw * 4);
                  bulk_g2s(stages + static_cast<size_t>(ring.slot) * P.stage_bytes,
                           vslab + static_cast<size_t>(t0) * g.sw, static_cast<uint32_t>(nt) * g.sw * 4,
                           &full_bar[ring.slot], policy_kv);
                }
                __syncwarp();
                ring.advance(S);
              }
            }
          }
          if (pstamp) pstamp[5] = global_ns();
          continue;
        }
        const int u0 = static_cast<int>(static_cast<long long>(cta) * ph.units / G);
        const int u1 = static_cast<int>(static_cast<long long>(cta + 1) * ph.units / G);
        const int rpu = ph.swiglu ? 2 : 1;
        const int row_bytes = ph.in_dim * wbytes;
        if (ph.chunks_per_row == 1) {
          const int ups = ph.rows_per_stage / rpu;
          for (int u = u0; u < u1; u += ups) {
            const int n = min(ups, u1 - u);
            const int nrows = n * rpu;
            mbar_wait(&empty_bar[ring.slot], ring.parity ^ 1u);
            unsigned char* dst = stages + static_cast<size_t>(ring.slot) * P.stage_bytes;
            if (lane == 0)
              mbar_expect_tx(&full_bar[ring.slot],
                             static_cast<uint32_t>(nrows) * (row_bytes + ph.scale_row_bytes));
            __syncwarp();
            for (int i = lane; i < nrows; i += 32) {
              const RowRef rr = resolve_row(ph, u + i / rpu, i % rpu);
              const long long e = static_cast<long long>(rr.row) * ph.in_dim;
              const unsigned char* src = static_cast<const unsigned char*>(ph.seg[rr.seg].w) + e * wbytes;
              bulk_g2s(dst + static_cast<size_t>(i) * row_bytes, src, row_bytes, &full_bar[ring.slot], policy);
              if (kInt8) {
                const long long g0 = ph.group_shift >= 0 ? (e >> ph.group_shift) : (e / ph.group_size);
                bulk_g2s(dst + ph.scale_off + static_cast<size_t>(i) * ph.scale_row_bytes,
                         ph.seg[rr.seg].scales + g0, ph.scale_row_bytes, &full_bar[ring.slot], policy);
              }
            }
            ring.advance(S);
          }
        } else {
          for (int u = u0; u < u1; ++u) {
            const RowRef rr = resolve_row(ph, u, 0);
            const unsigned char* src = static_cast<const unsigned char*>(ph.seg[rr.seg].w) +
                                       static_cast<long long>(rr.row) * row_bytes;
            for (int c = 0; c < ph.chunks_per_row; ++c) {
              const int e0 = c * ph.chunk_elems;
              const int ne = min(ph.chunk_elems, ph.in_dim - e0);
              mbar_wait(&empty_bar[ring.slot], ring.parity ^ 1u);
              if (lane == 0) {
                mbar_expect_tx(&full_bar[ring.slot], static_cast<uint32_t>(ne) * wbytes);
                bulk_g2s(stages + static_cast<size_t>(ring.slot) * P.stage_bytes,
                         src + static_cast<size_t>(e0) * wbytes, static_cast<uint32_t>(ne) * wbytes,
                         &full_bar[ring.slot], policy);
              }
              __syncwarp();
              ring.advance(S);
            }
          }
        }
        if (pstamp) pstamp[5] = global_ns();
      }
    }
    return;
  }

  // =============================== consumer warps ===============================================
  unsigned bar_target = P.barrier_base;
  unsigned flag_base = 0;  // attn_flags are zeroed before every launch
  int token = P.state->token;
  if (static_cast<unsigned>(token) >= static_cast<unsigned>(P.vocab_size)) token = 0;
  int pos = P.state->pos;
  int step = P.state->step;

  for (int tok = 0; tok < P.n_tokens; ++tok) {
    const float* emb_row = P.tok_emb + static_cast<size_t>(token) * P.dim;
    ArgBest best{0.f, -1};  // per lane; folded across the CTA after the classifier phase

    const bool prof_on = P.prof != nullptr && tok == P.prof_token && tid == 0;
    for (int pi = 0; pi < P.n_phases; ++pi) {
      {
        // (the grid barrier that ended the previous phase is the hazard fence for this copy)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.phases + pi);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_phase_cons);
        if (tid < static_cast<int>(sizeof(Phase) / 4)) dst[tid] = __ldg(src + tid);
        consumer_sync();
      }
      const Phase& ph = s_phase_cons;
      unsigned long long* stamp =
          prof_on ? P.prof + (static_cast<size_t>(cta) * P.n_phases + pi) * 8 : nullptr;
      if (stamp) stamp[0] = global_ns();

      if (ph.kind == kPhaseAttention) {
        if (stamp) stamp[1] = stamp[6] = stamp[7] = stamp[0];
        attention_phase(P, ph, cta, G, pos, xs, stages, full_bar, empty_bar, ring, flag_base, stamp);
        if (stamp) stamp[2] = global_ns();
        grid_barrier(P.barrier, bar_target, G);
        if (stamp) stamp[3] = global_ns();
        continue;
      }

      // ---- stage the input vector (and RMS-normalise it) --------------------------------------
      const int M = ph.in_dim;
      {
        const float* xg = ph.x_from_emb ? emb_row : ph.x;
        const float4* xg4 = reinterpret_cast<const float4*>(xg);
        float4* xs4w = reinterpret_cast<float4*>(xs);
        const int n4 = M >> 2;
        // up to kMaxNormRegs float4 of the (static) norm weight ride in registers while x arrives
        constexpr int kMaxNormRegs = 4;
        float4 nw[kMaxNormRegs];
        const float4* nw4 = reinterpret_cast<const float4*>(ph.norm_w);
        const bool norm_regs = ph.norm_w != nullptr && n4 <= kMaxNormRegs * kConsumerThreads;
        if (norm_regs) {
#pragma unroll
          for (int k = 0; k < kMaxNormRegs; ++k) {
            const int i = tid + k * kConsumerThreads;
            if (i < n4) nw[k] = __ldg(nw4 + i);
          }
        }
        for (int i = tid; i < n4; i += kConsumerThreads) xs4w[i] = __ldcg(xg4 + i);
        consumer_sync();
        if (ph.norm_w != nullptr) {
          if (warp == 0) {
            const float sc = rms_scale_smem(xs, M, ph.norm_eps, lane);
            if (lane == 0) s_bcast = sc;
          }
          consumer_sync();
          const float sc = s_bcast;
          // rmsnorm_kernel.cu:41-45: (scale * x) * w
          if (norm_regs) {
#pragma unroll
            for (int k = 0; k < kMaxNormRegs; ++k) {
              const int i = tid + k * kConsumerThreads;
              if (i < n4) {
                float4 v = xs4w[i];
                v.x = __fmul_rn(__fmul_rn(sc, v.x), nw[k].x);
                v.y = __fmul_rn(__fmul_rn(sc, v.y), nw[k].y);
                v.z = __fmul_rn(__fmul_rn(sc, v.z), nw[k].z);
                v.w = __fmul_rn(__fmul_rn(sc, v.w), nw[k].w);
                xs4w[i] = v;
              }
            }
          } else {
            for (int i = tid; i < M; i += kConsumerThreads)
              xs[i] = __fmul_rn(__fmul_rn(sc, xs[i]), ph.norm_w[i]);
          }
          consumer_sync();
        }
      }
      if (stamp) stamp[1] = global_ns();

      RowCtx cx;
      cx.xs4 = reinterpret_cast<const float4*>(xs);
      cx.residual = ph.residual_from_emb ? emb_row : ph.residual;
      cx.pos = pos;
      cx.head_size = P.head_size;
      cx.seq_len = P.seq_len;

      const int u0 = static_cast<int>(static_cast<long long>(cta) * ph.units / G);
      const int u1 = static_cast<int>(static_cast<long long>(cta + 1) * ph.units / G);
      const int rpu = ph.swiglu ? 2 : 1;

      if (ph.chunks_per_row == 1) {
        const int ups = ph.rows_per_stage / rpu;
        for (int u = u0; u < u1; u += ups) {
          mbar_wait(&full_bar[ring.slot], ring.parity);  // all warps track every fill (see attention)
          if (ring.count % kNW == warp) {  // this warp owns the stage
            const int n = min(ups, u1 - u);
            const unsigned char* sbase = stages + static_cast<size_t>(ring.slot) * P.stage_bytes;
            process_stage<kInt8>(ph, cx, sbase, u, n, lane, best);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[ring.slot]);
          }
          ring.advance(S);
        }
      } else {
        // rows longer than a stage (fp32 only): the owning warp carries its partial sums across
        // the row's consecutive stages; chunk boundaries are multiples of 128 packs so every
        // virtual thread still sees its packs in increasing order.
        for (int u = u0; u < u1; ++u) {
          const bool mine = ((u - u0) % kNW) == warp;
          float bias_v = 0.f, res_v = 0.f;
          RowRef rr{0, 0};
          if (mine && lane == 0) {
            rr = resolve_row(ph, u, 0);
            if (ph.seg[rr.seg].bias != nullptr) bias_v = __ldg(ph.seg[rr.seg].bias + rr.row);
            if (cx.residual != nullptr) res_v = __ldcg(cx.residual + rr.row);
          }
          float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
          for (int c = 0; c < ph.chunks_per_row; ++c) {
            mbar_wait(&full_bar[ring.slot], ring.parity);
            if (mine) {
              const int e0 = c * ph.chunk_elems;
              const int ne = min(ph.chunk_elems, M - e0);
              const float4* w[1] = {reinterpret_cast<const float4*>(
                  stages + static_cast<size_t>(ring.slot) * P.stage_bytes)};
              accum_f32<1>(w, cx.xs4 + (e0 >> 2), ne >> 2, lane, acc);
              __syncwarp();
              if (lane == 0) mbar_arrive(&empty_bar[ring.slot]);
            }
            ring.advance(S);
          }
          if (mine) {
            const float d0 = block128_sum_vt(acc[0]);
            if (lane == 0) {
              const Seg& sg = ph.seg[rr.seg];
              float v = d0;
              if (sg.bias != nullptr) v = __fadd_rn(v, bias_v);
              if (cx.residual != nullptr) v = __fadd_rn(res_v, v);
              if (sg.head_major) {
                sg.out[v_index(rr.row, pos, P.head_size, P.seq_len)] = v;
              } else {
                sg.out[static_cast<long long>(pos) * sg.pos_stride + rr.row] = v;
              }
              if (ph.argmax) arg_fold(best, v, rr.row);
            }
          }
        }
      }

      if (ph.argmax) {
        // per-CTA (max, lowest index) of the classifier rows this CTA produced
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
          arg_fold(best, __shfl_xor_sync(kFull, best.v, off), __shfl_xor_sync(kFull, best.i, off));
        if (lane == 0) {
          s_argv[warp] = best.v;
          s_argi[warp] = best.i;
        }
        consumer_sync();
        if (tid == 0) {
          ArgBest b{0.f, -1};
          for (int w = 0; w < kNW; ++w) arg_fold(b, s_argv[w], s_argi[w]);
          P.arg_val[cta] = b.v;
          P.arg_idx[cta] = b.i;
        }
      }
      if (stamp) stamp[2] = global_ns();
      grid_barrier(P.barrier, bar_target, G);
      if (stamp) stamp[3] = global_ns();
    }

    // ---- greedy id: every CTA folds the per-CTA partials identically (argmax_kernel.cu:49-71
    // semantics: maximum value, lowest index) -------------------------------------------------------
    ArgBest b{0.f, -1};
    for (int c = lane; c < G; c += 32) arg_fold(b, __ldcg(P.arg_val + c), __ldcg(P.arg_idx + c));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
      arg_fold(b, __shfl_xor_sync(kFull, b.v, off), __shfl_xor_sync(kFull, b.i, off));
    const int next = b.i < 0 ? 0 : b.i;
    if (cta == 0 && tid == 0) {
      if (P.out_tokens != nullptr && step < P.max_steps) P.out_tokens[step] = next;
    }
    token = (P.teacher != nullptr && step + 1 < P.max_steps) ? P.teacher[step + 1] : next;
    if (static_cast<unsigned>(token) >= static_cast<unsigned>(P.vocab_size)) token = 0;
    pos += 1;
    step += 1;
    if (cta == 0 && tid == 0 && tok == P.n_tokens - 1) {
      P.state->token = token;
      P.state->pos = pos;
      P.state->step = step;
      P.state->next = next;
    }
  }
}

}  // namespace mega

// ================================== host side ======================================================
using mega::Params;
using mega::Phase;

int MegaEngine::init(const MegaModel& m, cudaStream_t stream) {
  model_ = m;
  stream_ = stream;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return KLLM_E_NODEVICE;
  int sms = 0, coop = 0, max_smem = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (!coop) return KLLM_E_UNSUPPORTED;
  grid_ = sms;

  const int dim = m.dim, hid = m.hidden_dim, hs = m.head_size, kvd = m.kv_dim;
  const int q_rows = m.head_num * hs;
  const bool int8 = m.group_size > 0;
  const int wb = int8 ? 1 : 4;
  // shapes the ring handles: 16-byte rows, 128-byte aligned kv rows (L1-cached reads stay exact)
  if ((dim & 3) || (hid & 3) || (q_rows & 3) || (hs & 3)) return KLLM_E_UNSUPPORTED;
  // value slabs are 32 output dims wide (one warp each): head_size 32k, or a single narrow slab
  if (!((hs % 32 == 0 && hs / 32 <= mega::kNW) || hs < 32)) return KLLM_E_UNSUPPORTED;
  if (int8 && ((dim & 15) || (hid & 15) || (q_rows & 15) || (m.group_size & 3))) return KLLM_E_UNSUPPORTED;
  if ((hs * 4) % 16 != 0) return KLLM_E_UNSUPPORTED;
  if (int8) {
    const int dims[3] = {dim, hid, q_rows};
    for (int d : dims)
      if (d % m.group_size != 0 || ((d / m.group_size) * 4) % 16 != 0) return KLLM_E_UNSUPPORTED;
  }

  // ---- shared memory plan -----------------------------------------------------------------------
  const int max_in = std::max(std::max(dim, hid), q_rows);
  int xbuf = max_in * 4;
  const int attn_ws = (m.kv_mul + 1) * hs * 4;
  xbuf = std::max(xbuf, attn_ws);
  xbuf = (xbuf + 127) & ~127;
  const int budget = max_smem - xbuf - 2048;  // static smem + slack
  const int min_row = std::min(std::min(dim, hid), q_rows) * wb;
  (void)min_row;
  int stage_bytes = 48 * 1024;
  if (const char* e = getenv("KLLM_STAGE_BYTES")) stage_bytes = atoi(e);
  stage_bytes = (stage_bytes + 127) & ~127;
  int stages = budget / stage_bytes;
  if (stages > mega::kMaxStages) stages = mega::kMaxStages;
  if (const char* e = getenv("KLLM_STAGES")) stages = std::min(stages, atoi(e));
  if (stages < 2) return KLLM_E_UNSUPPORTED;
  stage_bytes_ = stage_bytes;
  stages_ = stages;
  // K tile: lanes carry up to 8 timesteps each; V tile: one slab of min(hs,32) dims
  attn_tk_ = std::min(stage_bytes / (hs * 4), 256) & ~31;
  attn_tv_ = std::min(stage_bytes / (std::min(hs, 32) * 4), 1024) & ~3;
  if (attn_tk_ < 32 || attn_tv_ < 4) return KLLM_E_UNSUPPORTED;
  xbuf_bytes_ = xbuf;
  smem_bytes_ = static_cast<size_t>(xbuf) + static_cast<size_t>(stages) * stage_bytes;

  // ---- phase table ---------------------------------------------------------------------------------
  std::vector<Phase> ph;
  auto plan = [&](Phase& p) -> int {
    const int row_bytes = p.in_dim * wb;
    p.group_size = m.group_size;
    p.group_shift = -1;
    if (int8 && (m.group_size & (m.group_size - 1)) == 0) {
      int s = 0;
      while ((1 << s) < m.group_size) ++s;
      p.group_shift = s;
    }
    p.scale_row_bytes = int8 ? (p.in_dim / m.group_size) * 4 : 0;
    const int rpu = p.swiglu ? 2 : 1;
    const int per_row = row_bytes + p.scale_row_bytes;
    if (per_row * rpu <= stage_bytes) {
      int rows = stage_bytes / per_row;
      rows -= rows % rpu;
      rows = std::min(rows, 32);  // one bulk copy per producer lane
      p.rows_per_stage = rows;
      p.chunks_per_row = 1;
      p.chunk_elems = p.in_dim;
      p.scale_off = ((rows * row_bytes) + 127) & ~127;
      if (p.scale_off + rows * p.scale_row_bytes > stage_bytes) {
        // shrink until weights + scales fit
        while (rows > rpu && (((rows * row_bytes + 127) & ~127) + rows * p.scale_row_bytes) > stage_bytes)
          rows -= rpu;
        p.rows_per_stage = rows;
        p.scale_off = ((rows * row_bytes) + 127) & ~127;
      }
    } else {
      if (int8 || p.swiglu) return KLLM_E_UNSUPPORTED;
      const int chunk_max = (stage_bytes / 4) & ~511;  // multiple of 128 packs
      p.chunks_per_row = (p.in_dim + chunk_max - 1) / chunk_max;
      int ce = (p.in_dim + p.chunks_per_row - 1) / p.chunks_per_row;
      ce = (ce + 511) & ~511;
      p.chunk_elems = ce;
      p.chunks_per_row = (p.in_dim + ce - 1) / ce;
      p.rows_per_stage = 1;
      p.scale_off = 0;
    }
    return 0;
  };

  const float eps = flavour_eps(m.flavour);
  for (int l = 0; l < m.layer_num; ++l) {
    const size_t layer_off = static_cast<size_t>(l) * m.seq_len * kvd;
    {  // attention_rms + q | k | v (+bias).  k goes to k_raw (rotated later), v into the cache.
      Phase p{};
      p.kind = mega::kPhaseGemv;
      p.in_dim = dim;
      p.n_seg = 3;
      p.x = m.x;
      p.x_from_emb = (l == 0);
      p.norm_w = m.attn_norm[l];
      p.norm_eps = eps;
      p.seg[0] = {m.wq[l], int8 ? m.sq[l] : nullptr, m.bq ? m.bq[l] : nullptr, m.q, 0, q_rows, 0};
      p.seg[1] = {m.wk[l], int8 ? m.sk[l] : nullptr, m.bk ? m.bk[l] : nullptr, m.k_raw, 0, kvd, 0};
      p.seg[2] = {m.wv[l], int8 ? m.sv[l] : nullptr, m.bv ? m.bv[l] : nullptr,
                  m.value_cache + layer_off, 0, kvd, 1};
      p.units = q_rows + 2 * kvd;
      if (int rc = plan(p)) return rc;
      ph.push_back(p);
    }
    {
      Phase p{};
      p.kind = mega::kPhaseAttention;
      p.layer = l;
      ph.push_back(p);
    }
    {  // wo + residual (llama3.cpp:672-684)
      Phase p{};
      p.kind = mega::kPhaseGemv;
      p.in_dim = q_rows;
      p.n_seg = 1;
      p.x = m.attn_out;
      p.seg[0] = {m.wo[l], int8 ? m.so[l] : nullptr, nullptr, m.x, 0, dim, 0};
      p.residual = m.x;
      p.residual_from_emb = (l == 0);
      p.units = dim;
      if (int rc = plan(p)) return rc;
      ph.push_back(p);
    }
    {  // ffn rmsnorm + w1 | w3 -> swiglu (llama3.cpp:686-708)
      Phase p{};
      p.kind = mega::kPhaseGemv;
      p.in_dim = dim;
      p.n_seg = 2;
      p.swiglu = 1;
      p.x = m.x;
      p.norm_w = m.ffn_norm[l];
      p.norm_eps = eps;
      p.seg[0] = {m.w1[l], int8 ? m.s1[l] : nullptr, nullptr, m.h, 0, hid, 0};
      p.seg[1] = {m.w3[l], int8 ? m.s3[l] : nullptr, nullptr, nullptr, 0, hid, 0};
      p.units = hid;
      if (int rc = plan(p)) return rc;
      ph.push_back(p);
    }
    {  // w2 + residual (llama3.cpp:711-719)
      Phase p{};
      p.kind = mega::kPhaseGemv;
      p.in_dim = hid;
      p.n_seg = 1;
      p.x = m.h;
      p.seg[0] = {m.w2[l], int8 ? m.s2[l] : nullptr, nullptr, m.x, 0, dim, 0};
      p.residual = m.x;
      p.units = dim;
      if (int rc = plan(p)) return rc;
      ph.push_back(p);
    }
  }
  {  // final rmsnorm + classifier (+ argmax partials)
    Phase p{};
    p.kind = mega::kPhaseGemv;
    p.in_dim = dim;
    p.n_seg = 1;
    p.x = m.x;
    p.norm_w = m.final_norm;
    p.norm_eps = eps;
    p.seg[0] = {m.wcls, int8 ? m.scls : nullptr, nullptr, m.logits, 0, m.vocab_size, 0};
    p.units = m.vocab_size;
    p.argmax = 1;
    if (int rc = plan(p)) return rc;
    ph.push_back(p);
  }
  n_phases_ = static_cast<int>(ph.size());
  n_barriers_per_token_ = n_phases_;

  if (cudaMalloc(&d_phases_, sizeof(Phase) * ph.size()) != cudaSuccess) return static_cast<int>(cudaErrorMemoryAllocation);
  cudaMemcpyAsync(d_phases_, ph.data(), sizeof(Phase) * ph.size(), cudaMemcpyHostToDevice, stream);
  if (cudaMalloc(&d_barrier_, 128) != cudaSuccess) return static_cast<int>(cudaErrorMemoryAllocation);
  cudaMemsetAsync(d_barrier_, 0, 128, stream);
  n_kv_heads_ = kvd / hs;
  if (cudaMalloc(&d_flags_, sizeof(unsigned) * n_kv_heads_) != cudaSuccess) return static_cast<int>(cudaErrorMemoryAllocation);
  if (cudaMalloc(&d_arg_val_, sizeof(float) * grid_) != cudaSuccess ||
      cudaMalloc(&d_arg_idx_, sizeof(int) * grid_) != cudaSuccess)
    return static_cast<int>(cudaErrorMemoryAllocation);
  cudaStreamSynchronize(stream);  // ph (host vector) must outlive the async copy

  kernel_ = int8 ? reinterpret_cast<const void*>(mega::decode_megakernel<true>)
                 : reinterpret_cast<const void*>(mega::decode_megakernel<false>);
  cudaError_t e = cudaFuncSetAttribute(kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_bytes_));
  if (e != cudaSuccess) return static_cast<int>(e);
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel_, mega::kThreads, smem_bytes_);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (occ < 1) return KLLM_E_UNSUPPORTED;
  barrier_base_ = 0;
  ready_ = true;
  return 0;
}

void MegaEngine::destroy() {
  if (d_phases_) cudaFree(d_phases_);
  if (d_barrier_) cudaFree(d_barrier_);
  if (d_flags_) cudaFree(d_flags_);
  d_flags_ = nullptr;
  if (d_arg_val_) cudaFree(d_arg_val_);
  if (d_arg_idx_) cudaFree(d_arg_idx_);
  d_phases_ = nullptr;
  d_barrier_ = nullptr;
  d_arg_val_ = nullptr;
  d_arg_idx_ = nullptr;
  ready_ = false;
}

int MegaEngine::run(int n_tokens, const int32_t* teacher_dev, unsigned long long* prof_dev,
                    int prof_token) {
  if (!ready_) return KLLM_E_STATE;
  Params P{};
  const MegaModel& m = model_;
  P.phases = static_cast<const Phase*>(d_phases_);
  P.n_phases = n_phases_;
  P.n_tokens = n_tokens;
  P.num_stages = stages_;
  P.stage_bytes = stage_bytes_;
  P.xbuf_bytes = xbuf_bytes_;
  P.attn_tk = attn_tk_;
  P.attn_tv = attn_tv_;
  P.group_size = m.group_size;
  P.dim = m.dim;
  P.vocab_size = m.vocab_size;
  P.head_num = m.head_num;
  P.head_size = m.head_size;
  P.kv_dim = m.kv_dim;
  P.kv_mul = m.kv_mul;
  P.seq_len = m.seq_len;
  P.flavour = m.flavour;
  P.tok_emb = m.tok_emb;
  P.q = m.q;
  P.k_raw = m.k_raw;
  P.attn_out = m.attn_out;
  P.score = m.score;
  P.key_cache = m.key_cache;
  P.value_cache = m.value_cache;
  P.sin_cache = m.sin_cache;
  P.cos_cache = m.cos_cache;
  P.state = static_cast<mega::State*>(m.state);
  P.out_tokens = m.out_tokens;
  P.teacher = teacher_dev;
  P.max_steps = m.seq_len;
  P.barrier = static_cast<unsigned*>(d_barrier_);
  P.barrier_base = barrier_base_;
  P.arg_val = static_cast<float*>(d_arg_val_);
  P.arg_idx = static_cast<int*>(d_arg_idx_);
  P.prof = prof_dev;
  P.prof_token = prof_token;
  P.attn_flags = static_cast<unsigned*>(d_flags_);
  cudaError_t me = cudaMemsetAsync(d_flags_, 0, sizeof(unsigned) * n_kv_heads_, stream_);
  if (me != cudaSuccess) return static_cast<int>(me);
  void* args[] = {&P};
  cudaError_t e = cudaLaunchCooperativeKernel(kernel_, dim3(grid_), dim3(mega::kThreads), args,
                                              smem_bytes_, stream_);
  if (e != cudaSuccess) return static_cast<int>(e);
  barrier_base_ += static_cast<unsigned>(n_tokens) * static_cast<unsigned>(n_barriers_per_token_) *
                   static_cast<unsigned>(grid_);
  count_launch();
  return 0;
}

}  // namespace kllm
