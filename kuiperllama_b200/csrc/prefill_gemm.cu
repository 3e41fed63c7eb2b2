// Batched GEMM for prompt prefill on the 5th-generation tensor cores (sm_100a):
//     out[T, N] = x[T, K] . w[N, K]^T        fp32 storage, TF32 multiply, fp32 accumulate in TMEM
//
// The reference feeds a prompt through one full single-token forward per position (demo/main.cpp:
// 18-23, llama3.cpp:147-167) -- T GEMVs that stream every weight T times.  With the T prompt rows
// as the N dimension of a tcgen05.mma the weights are streamed once per 256 tokens.
//
// This is an explicitly TOLERANCED kernel: TF32 keeps 10 mantissa bits of each operand, so results
// agree with the fp32 GEMV path to ~1e-3 relative, not bit for bit (tests/test_prefill_gpu.py states
// the bound).  The bit-exact decode path never calls it.
//
// Structure (one CTA per 128 weight rows x one block of BN tokens, 6 warps):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D tiles of w [128 x 32 fp32] and x [BN x 32 fp32],
//            128-byte swizzle, into a 4-stage shared-memory ring (SASS: UTMALDG)
//   warp 1   MMA issuer: one thread, 4 x tcgen05.mma.kind::tf32 (M 128, N BN, K 8) per stage, D in
//            TMEM; tcgen05.commit frees the stage / signals the epilogue (SASS: UTCHMMA / UTCBAR)
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns per warp (SASS: LDTM), transposed store
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/kllm_b200.h"
#include "kllm_host.h"

namespace kllm {
namespace tc {

constexpr int BM = 128;     // weight rows per CTA = MMA M
constexpr int BK = 32;      // fp32 elements per K block = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 8;   // tf32: 32 bytes of K per tcgen05.mma
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 4;  // 16 KB
constexpr int THREADS = 192;

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// Shared-memory matrix descriptor of a K-major tile with 128-byte swizzle (cute::UMMA::SmemDescriptor,
// cutlass include/cute/arch/mma_sm100_desc.hpp): start address >> 4 in bits [0,14), leading byte
// offset (unused with swizzle; 1) in [16,30), stride byte offset = 8 rows x 128 B = 1024 B >> 4 in
// [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10),
// both K-major (bits 15, 16 clear), N >> 3 in [17,23), M >> 4 in [24,29).
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                 float* __restrict__ out, int T, int N, int K) {
  constexpr int B_BYTES = BN * BK * 4;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;  // power of two >= 32
  extern __shared__ uint8_t raw_smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw_smem) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tiles = base;
  uint8_t* b_tiles = base + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_tiles + STAGES * B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BM;  // first weight row (output feature) of this CTA
  const int t0 = blockIdx.y * BN;  // first token
  const int kblocks = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_addr(&full[s]), 1);
      mbar_init(smem_addr(&empty[s]), 1);
    }
    mbar_init(smem_addr(tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one warp allocates the accumulator's TMEM columns and later frees them
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_addr(&empty[s]), ph ^ 1u);
        const uint32_t bar = smem_addr(&full[s]);
        mbar_expect_tx(bar, A_BYTES + B_BYTES);  // out-of-range rows / columns are zero-filled and still counted
        tma_load_2d(smem_addr(a_tiles + s * A_BYTES), &map_w, bar, kb * BK, n0);
        tma_load_2d(smem_addr(b_tiles + s * B_BYTES), &map_x, bar, kb * BK, t0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer ----
      const uint32_t idesc = umma_idesc_tf32(BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_addr(&full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_addr(a_tiles + s * A_BYTES);
        const uint32_t b_addr = smem_addr(b_tiles + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)  // 32 bytes of K per instruction: step inside the swizzle atom
          umma_tf32(tmem_d, umma_desc_sw128(a_addr + k * UMMA_K * 4), umma_desc_sw128(b_addr + k * UMMA_K * 4), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(smem_addr(&empty[s]));  // the stage may be refilled once these MMAs have read it
      }
      umma_commit(smem_addr(tmem_full));  // accumulator complete
    }
  } else {
    // ---- epilogue: TMEM lane = weight row, column = token; a warp may touch lanes 32 (warp % 4) .. +31 ----
    const int quarter = warp & 3;
    mbar_wait(smem_addr(tmem_full), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = n0 + quarter * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_d + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < N) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int t = t0 + c + j;
          if (t < T) out[static_cast<size_t>(t) * N + row] = __uint_as_float(r[j]);  // lanes = consecutive rows: coalesced
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

// cuTensorMapEncodeTiled comes from the driver (libcuda); it is looked up at run time so that the
// library links and loads on a machine without a driver (the build box).
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// 2-D fp32 tensor [rows, cols] (row-major, cols contiguous), box [box_rows x 32 columns], 128-byte swizzle
static int make_map(CUtensorMap* map, const float* ptr, int rows, int cols, int box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (fn == nullptr) return KLLM_E_NODEVICE;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 4};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : KLLM_E_INVALID;
}

template <int BN>
static int launch(const float* x, const float* w, float* out, int T, int K, int N, cudaStream_t stream) {
  CUtensorMap map_w, map_x;
  if (int rc = make_map(&map_w, w, N, K, BM)) return rc;
  if (int rc = make_map(&map_x, x, T, K, BN)) return rc;
  const size_t smem = 1024 + static_cast<size_t>(STAGES) * (A_BYTES + BN * BK * 4) + 256;
  static bool configured = false;
  if (!configured) {
    const cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const dim3 grid((N + BM - 1) / BM, (T + BN - 1) / BN);
  gemm_tf32_kernel<BN><<<grid, THREADS, smem, stream>>>(map_w, map_x, out, T, N, K);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace tc
}  // namespace kllm

extern "C" int kllm_gemm_tf32(const float* x, const float* w, float* out, int n_tokens, int in_dim, int out_dim,
                              void* stream) {
  if (!x || !w || !out || n_tokens <= 0 || in_dim <= 0 || out_dim <= 0) return KLLM_E_INVALID;
  // TMA needs 16-byte aligned bases and row pitches
  if ((in_dim & 3) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return KLLM_E_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  using namespace kllm::tc;
  if (n_tokens <= 32) return launch<32>(x, w, out, n_tokens, in_dim, out_dim, s);
  if (n_tokens <= 64) return launch<64>(x, w, out, n_tokens, in_dim, out_dim, s);
  if (n_tokens <= 128) return launch<128>(x, w, out, n_tokens, in_dim, out_dim, s);
  return launch<256>(x, w, out, n_tokens, in_dim, out_dim, s);
}
