// Device side of the persistent decode megakernel (see megakernel.cu for the overview).
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "../../include/kllm_b200.h"
#include "kllm_device.cuh"
#include "megakernel.h"

namespace kllm {
namespace mega {

constexpr int kNW = 8;  // consumer warps
constexpr int kConsumerThreads = kNW * 32;
constexpr int kThreads = kConsumerThreads + 32;  // + one producer warp
constexpr int kMaxStages = 16;
constexpr int kMaxGroup = 4;  // rows a warp accumulates together (shares the x loads)

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Position in the stage ring.  Producer and every consumer warp walk the same stage sequence,
// so each keeps its own copy; `count` numbers the stages of this CTA since kernel start and
// decides which consumer warp owns a stage.
struct Ring {
  int slot;
  uint32_t parity;
  int count;
  __device__ __forceinline__ void advance(int stages) {
    ++count;
    if (++slot == stages) {
      slot = 0;
      parity ^= 1u;
    }
  }
};

// Grid barrier over the consumer threads of all CTAs.  Monotonic counter, wrap-safe compare.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, unsigned grid) {
  consumer_sync();
  target += grid;
  if (threadIdx.x == 0) {
    red_release_add(counter, 1u);
    // poll with relaxed loads (an acquire load invalidates L1 on every iteration), then one
    // acquire fence orders everything after the barrier
    while (static_cast<int>(ld_relaxed_u32(counter) - target) < 0) {
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  consumer_sync();
}

struct RowRef {
  int seg;
  int row;
};
__device__ __forceinline__ RowRef resolve_row(const Phase& ph, int unit, int sub) {
  if (ph.swiglu) return RowRef{sub, unit};
  int seg = 0, row = unit;
  if (ph.n_seg > 1 && row >= ph.seg[0].rows) {
    row -= ph.seg[0].rows;
    seg = 1;
    if (ph.n_seg > 2 && row >= ph.seg[1].rows) {
      row -= ph.seg[1].rows;
      seg = 2;
    }
  }
  return RowRef{seg, row};
}

// ---- exact-order accumulation from shared memory ----------------------------------------------
// fp32: virtual thread (lane + 32 j) owns packs base + 32 j + lane (matmul_kernel.cu:27-35).
// NR rows share every x load, so shared-memory traffic per weight byte is 1 + 1/NR -- the ring
// is consumed through the 128 B/cycle shared-memory pipe, which is what bounds the consumers.
template <int NR>
__device__ __forceinline__ void accum_f32(const float4* const (&w)[NR], const float4* x4,
                                          int n_packs, int lane, float (&acc)[NR][4]) {
  const int full = n_packs & ~127;
  for (int base = 0; base < full; base += 128) {
    float4 xv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xv[j] = x4[base + 32 * j + lane];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float4 wv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = w[r][base + 32 * j + lane];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[r][j] = __fadd_rn(dot4_ref(xv[j], wv[j]), acc[r][j]);
    }
  }
  if (full < n_packs) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = full + 32 * j + lane;
      if (idx < n_packs) {
        const float4 xv = x4[idx];
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r][j] = __fadd_rn(dot4_ref(xv, w[r][idx]), acc[r][j]);
      }
    }
  }
}

// int8: virtual thread (4 lane + e) owns elements 128 k + 4 lane + e (matmul_kernel.cu:70-74):
// acc = fma(x * scale, float(w), acc).  sc[r] = the row's staged scales (rows start on a group
// boundary, checked on the host).
template <int NR>
__device__ __forceinline__ void accum_w8(const uint32_t* const (&w)[NR], const float* const (&sc)[NR],
                                         const float4* x4, int M, int group_shift, int group_size,
                                         int lane, float (&acc)[NR][4]) {
  const int full_chunks = M >> 7;
  auto one = [&](int k) {
    const int i = (k << 7) + (lane << 2);
    const float4 xv = x4[i >> 2];
    const int g = group_shift >= 0 ? (i >> group_shift) : (i / group_size);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const uint32_t packed = w[r][i >> 2];
      const float s = sc[r][g];
      float wf[4];
      int8x4_to_float(packed, wf);
      acc[r][0] = __fmaf_rn(__fmul_rn(xv.x, s), wf[0], acc[r][0]);
      acc[r][1] = __fmaf_rn(__fmul_rn(xv.y, s), wf[1], acc[r][1]);
      acc[r][2] = __fmaf_rn(__fmul_rn(xv.z, s), wf[2], acc[r][2]);
      acc[r][3] = __fmaf_rn(__fmul_rn(xv.w, s), wf[3], acc[r][3]);
    }
  };
#pragma unroll 2
  for (int k = 0; k < full_chunks; ++k) one(k);
  if ((full_chunks << 7) + (lane << 2) < M) one(full_chunks);
}

// rmsnorm_kernel.cu:4-50 on x staged in shared memory (executed by warp 0).
__device__ __forceinline__ float rms_scale_smem(const float* xs, int n, float eps, int lane) {
  const int pack_num = n >> 2;
  const float4* xs4 = reinterpret_cast<const float4*>(xs);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int full = pack_num & ~127;
#pragma unroll 2
  for (int base = 0; base < full; base += 128) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = xs4[base + 32 * j + lane];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = acc[j];
      s = __fmaf_rn(v[j].x, v[j].x, s);
      s = __fmaf_rn(v[j].y, v[j].y, s);
      s = __fmaf_rn(v[j].z, v[j].z, s);
      s = __fmaf_rn(v[j].w, v[j].w, s);
      acc[j] = s;
    }
  }
  if (full < pack_num) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = full + 32 * j + lane;
      if (idx < pack_num) {
        const float4 v = xs4[idx];
        float s = acc[j];
        s = __fmaf_rn(v.x, v.x, s);
        s = __fmaf_rn(v.y, v.y, s);
        s = __fmaf_rn(v.z, v.z, s);
        s = __fmaf_rn(v.w, v.w, s);
        acc[j] = s;
      }
    }
  }
  float sum = block128_sum_vt(acc);
  sum = __shfl_sync(kFull, sum, 0);
  return rsqrtf(__fadd_rn(__fdiv_rn(sum, static_cast<float>(n)), eps));
}

struct ArgBest {
  float v;
  int i;
};
__device__ __forceinline__ void arg_fold(ArgBest& a, float ov, int oi) {
  if (oi >= 0 && (a.i < 0 || ov > a.v || (ov == a.v && oi < a.i))) {
    a.v = ov;
    a.i = oi;
  }
}

// Everything a row group needs besides the phase descriptor.
struct RowCtx {
  const float4* xs4;
  const float* residual;
  int pos;
  int head_size, seq_len;
};

// index into the slab-major value cache [kv_head][slab][seq_len][slab_width] for row (kv_dim index)
__device__ __forceinline__ size_t v_index(int row, int pos, int hs, int seq_len) {
  const int sw = hs < 32 ? hs : 32;
  const int slabs = hs / sw;
  const int kvh = row / hs, i = row % hs;
  return ((static_cast<size_t>(kvh) * slabs + i / sw) * seq_len + pos) * sw + (i % sw);
}

// One warp, NR rows of a stage at once: dot products in reference order, then bias / residual /
// SiLU*gate epilogue with lane r finishing row r.
template <int NR, bool kInt8>
__device__ __forceinline__ void process_rows(const Phase& ph, const RowCtx& cx,
                                             const unsigned char* sbase, int row0, int unit0,
                                             int lane, ArgBest& best) {
  const int M = ph.in_dim;
  const int row_bytes = M * (kInt8 ? 1 : 4);
  // addends first, so their L2 latency hides behind the accumulation
  float bias_v = 0.f, res_v = 0.f;
  RowRef rr{0, 0};
  if (!ph.swiglu && lane < NR) {
    rr = resolve_row(ph, unit0 + lane, 0);
    if (ph.seg[rr.seg].bias != nullptr) bias_v = __ldg(ph.seg[rr.seg].bias + rr.row);
    if (cx.residual != nullptr) res_v = __ldcg(cx.residual + rr.row);
  }
  float acc[NR][4];
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  float d[NR];
  if constexpr (!kInt8) {
    const float4* w[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r)
      w[r] = reinterpret_cast<const float4*>(sbase + static_cast<size_t>(row0 + r) * row_bytes);
    accum_f32<NR>(w, cx.xs4, M >> 2, lane, acc);
#pragma unroll
    for (int r = 0; r < NR; ++r) d[r] = block128_sum_vt(acc[r]);
  } else {
    const uint32_t* w[NR];
    const float* sc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      w[r] = reinterpret_cast<const uint32_t*>(sbase + static_cast<size_t>(row0 + r) * row_bytes);
      sc[r] = reinterpret_cast<const float*>(sbase + ph.scale_off +
                                             static_cast<size_t>(row0 + r) * ph.scale_row_bytes);
    }
    accum_w8<NR>(w, sc, cx.xs4, M, ph.group_shift, ph.group_size, lane, acc);
#pragma unroll
    for (int r = 0; r < NR; ++r) d[r] = block128_sum_quad(acc[r]);
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) d[r] = __shfl_sync(kFull, d[r], 0);

  if (ph.swiglu) {
    if constexpr (NR >= 2) {
#pragma unroll
      for (int u = 0; u < NR / 2; ++u) {
        const float o = swiglu_ref(d[2 * u], d[2 * u + 1]);  // swiglu_kernel.cu:16-21
        if (lane == u) ph.seg[0].out[unit0 + u] = o;
      }
    }
    return;
  }
  float v = d[0];
#pragma unroll
  for (int r = 1; r < NR; ++r)
    if (lane == r) v = d[r];
  if (lane < NR) {
    const Seg& sg = ph.seg[rr.seg];
    if (sg.bias != nullptr) v = __fadd_rn(v, bias_v);          // matmul.cpp:74-77: out + bias
    if (cx.residual != nullptr) v = __fadd_rn(res_v, v);       // llama3.cpp:683,719: x + out
    if (sg.head_major) {
      sg.out[v_index(rr.row, cx.pos, cx.head_size, cx.seq_len)] = v;
    } else {
      sg.out[static_cast<long long>(cx.pos) * sg.pos_stride + rr.row] = v;
    }
    if (ph.argmax) arg_fold(best, v, rr.row);
  }
}

template <bool kInt8>
__device__ __forceinline__ void process_stage(const Phase& ph, const RowCtx& cx,
                                              const unsigned char* sbase, int unit_first,
                                              int n_units, int lane, ArgBest& best) {
  const int rpu = ph.swiglu ? 2 : 1;
  int rows_left = n_units * rpu, row0 = 0, unit0 = unit_first;
  while (rows_left > 0) {
    int g = rows_left < kMaxGroup ? rows_left : kMaxGroup;
    if (rpu == 2) g &= ~1;
    switch (g) {
      case 4: process_rows<4, kInt8>(ph, cx, sbase, row0, unit0, lane, best); break;
      case 3: process_rows<3, kInt8>(ph, cx, sbase, row0, unit0, lane, best); break;
      case 2: process_rows<2, kInt8>(ph, cx, sbase, row0, unit0, lane, best); break;
      default: process_rows<1, kInt8>(ph, cx, sbase, row0, unit0, lane, best); break;
    }
    rows_left -= g;
    row0 += g;
    unit0 += g / rpu;
  }
}

// ---- attention geometry ------------------------------------------------------------------------
// KV layout (persistent engine only; kllm_decoder_read_kv converts back to [L][seq][kv_dim]):
//   K [L][kv_head][head_size/4][seq_len][4]  -- 16-byte chunk c of timestep t at ((c*seq_len)+t)*4
//       floats: a tile of T timesteps is hs/4 contiguous runs of T*16 bytes, and "lane t reads
//       chunk c" is a conflict-free 128-bit shared-memory access;
//   V [L][kv_head][slab][seq_len][sw], sw = min(hs,32)  -- one warp owns a slab of sw output
//       dims; a tile of T timesteps of a slab is one contiguous block and "lane i walks
//       column i" is conflict-free.
// Rows t < pos were written by earlier tokens, so -- like weights -- the producer streams them
// through the ring ahead of time; only row pos comes from registers / a direct load.
struct AttnGeom {
  int sw, slabs, tk, tv;
};
__device__ __forceinline__ AttnGeom attn_geom(const Params& P) {
  AttnGeom g;
  g.sw = P.head_size < 32 ? P.head_size : 32;
  g.slabs = P.head_size / g.sw;
  g.tk = P.attn_tk;
  g.tv = P.attn_tv;
  return g;
}
__device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace mega
}  // namespace kllm
