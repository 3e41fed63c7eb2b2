// Tensor-parallel exchange for the decode path: the all-reduce after o_proj and after down_proj
// (the reference itself is single-GPU, llama3.cpp:118 pins device 0).
//
// One process per GPU.  Two transports behind one handle:
//   * KLLM_COMM_PEER  one-shot all-reduce over NVLink peer memory (CUDA IPC): every rank stores
//                     its partial vector straight into every peer's receive slot, publishes a
//                     sequence flag (st.release.sys), waits for the world's flags in its OWN
//                     memory (ld.acquire.sys), and sums the slots in rank order -- fused with the
//                     residual add that follows in the reference (llama3.cpp:683-684, 719).
//                     One 8-16 KiB vector per call: this is a latency problem, not a bandwidth
//                     one, so there is no ring and no reduce-scatter.  Rank-ordered summation
//                     makes every rank hold bit-identical x, run to run.
//   * KLLM_COMM_NCCL  ncclAllReduce on the decoder's stream (capturable into its CUDA graph),
//                     libnccl resolved at run time with dlopen so the single-GPU library has no
//                     NCCL dependency.
// Rendezvous (exchanging the NCCL id / the IPC handles) is the caller's job: torch.distributed
// in bench.py and the tests, any out-of-band channel elsewhere.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>

#include "kllm_host.h"

struct Id128 {
  char bytes[128];
};

namespace {
using namespace kllm;

constexpr int kMaxWorld = 8;
constexpr int kFlagStride = 32;  // uint32 per slot row (128 B)

// ---- NCCL through dlopen (ABI of nccl.h 2.x: ncclUniqueId = 128 bytes, ncclFloat32 = 7, ncclSum = 0)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /* ncclUniqueId by value */ Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  // a libnccl already mapped into the process (torch's) wins: same soname
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.lib, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
    dlclose(api.lib);
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

struct PeerTable {
  float* data[kMaxWorld];      // each rank's receive area: [2 slots][world][max_count]
  uint32_t* flags[kMaxWorld];  // each rank's flag area:    [2 slots][kFlagStride]
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// out[i] = (residual ? residual[i] : 0) + (p_0[i] + p_1[i] + ... + p_{world-1}[i]), p_r = rank r's
// `partial`.  One CTA: 8-16 KiB per call, the cost is the NVLink round trip, not the copy.
__global__ void __launch_bounds__(1024)
allreduce_oneshot_kernel(PeerTable peers, int rank, int world, const float* __restrict__ partial,
                         const float* residual, float* out, int count, int max_count, uint32_t* seq_ptr) {
  const uint32_t seq = *seq_ptr + 1;
  const int slot = seq & 1;
  const int n4 = count >> 2;
  const float4* src = reinterpret_cast<const float4*>(partial);
  // 1. push my partial into every rank's slot row `rank` (own copy last: it is the cheapest)
  for (int k = 1; k <= world; ++k) {
    const int p = (rank + k) % world;
    float4* dst = reinterpret_cast<float4*>(peers.data[p] + (static_cast<size_t>(slot) * world + rank) * max_count);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(peers.flags[threadIdx.x] + slot * kFlagStride + rank, seq);
    // 2. wait until every rank's partial for this call has landed here
    const uint32_t* mine = peers.flags[rank] + slot * kFlagStride + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) != seq) {
      if (clock64() - t0 > 120000000000LL) {  // ~1 min: a peer died; fail the launch instead of hanging
        printf("kllm tp: rank %d timed out waiting for rank %d (call %u)\n", rank, threadIdx.x, seq);
        __trap();
      }
    }
  }
  __syncthreads();
  // 3. rank-ordered sum (+ residual)
  const float* base = peers.data[rank] + static_cast<size_t>(slot) * world * max_count;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 acc = __ldcg(reinterpret_cast<const float4*>(base) + i);
    for (int r = 1; r < world; ++r) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(base + static_cast<size_t>(r) * max_count) + i);
      acc.x = __fadd_rn(acc.x, v.x), acc.y = __fadd_rn(acc.y, v.y);
      acc.z = __fadd_rn(acc.z, v.z), acc.w = __fadd_rn(acc.w, v.w);
    }
    if (residual != nullptr) {
      const float4 x = reinterpret_cast<const float4*>(residual)[i];
      acc.x = __fadd_rn(x.x, acc.x), acc.y = __fadd_rn(x.y, acc.y);
      acc.z = __fadd_rn(x.z, acc.z), acc.w = __fadd_rn(x.w, acc.w);
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
  if (threadIdx.x == 0) *seq_ptr = seq;
}

__global__ void residual_add_kernel(const float* residual, const float* sum, float* out, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = __fadd_rn(residual[i], sum[i]);
}
}  // namespace

struct kllm_comm {
  int world = 1, rank = 0, backend = KLLM_COMM_PEER, max_count = 0;
  // peer transport
  void* local = nullptr;  // IPC-exported allocation: flags then data
  size_t local_bytes = 0;
  void* remote[kMaxWorld] = {};
  PeerTable table{};
  uint32_t* seq = nullptr;
  bool connected = false;
  // nccl transport
  void* nccl = nullptr;
};

namespace {
size_t flag_bytes() { return 2 * kFlagStride * sizeof(uint32_t); }
size_t data_bytes(const kllm_comm* c) { return sizeof(float) * 2 * static_cast<size_t>(c->world) * c->max_count; }
// the persistent kernel's tagged exchange area follows the one-shot kernel's slots
unsigned long long* tagged_of(const kllm_comm* c, void* base) {
  return reinterpret_cast<unsigned long long*>(static_cast<char*>(base) + flag_bytes() + data_bytes(c));
}
void fill_table(kllm_comm* c, int r, void* base) {
  c->table.flags[r] = static_cast<uint32_t*>(base);
  c->table.data[r] = reinterpret_cast<float*>(static_cast<char*>(base) + flag_bytes());
}
}  // namespace

extern "C" {

int kllm_comm_unique_id(unsigned char* out128) {
  if (!out128) return KLLM_E_INVALID;
  NcclApi* api = nccl_api();
  if (!api) return KLLM_E_COMM;
  return api->GetUniqueId(out128) == 0 ? 0 : KLLM_E_COMM;
}

int kllm_comm_create(int world, int rank, int backend, int max_count, const unsigned char* nccl_id128,
                     kllm_comm** out) {
  if (!out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || max_count <= 0 || (max_count & 3))
    return KLLM_E_INVALID;
  if (backend != KLLM_COMM_PEER && backend != KLLM_COMM_NCCL) return KLLM_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return KLLM_E_NODEVICE;
  auto* c = new kllm_comm();
  c->world = world, c->rank = rank, c->backend = backend, c->max_count = max_count;
  if (backend == KLLM_COMM_NCCL) {
    NcclApi* api = nccl_api();
    if (!api || !nccl_id128) {
      delete c;
      return api ? KLLM_E_INVALID : KLLM_E_COMM;
    }
    Id128 id;
    std::memcpy(id.bytes, nccl_id128, sizeof(id.bytes));
    const int rc = api->CommInitRank(&c->nccl, world, id, rank);
    if (rc != 0) {
      std::fprintf(stderr, "kllm tp: ncclCommInitRank failed: %s\n",
                   api->GetErrorString ? api->GetErrorString(rc) : "?");
      delete c;
      return KLLM_E_COMM;
    }
    c->connected = true;
  } else {
    c->local_bytes = flag_bytes() + data_bytes(c) + sizeof(unsigned long long) * 2 * static_cast<size_t>(world) * max_count;
    if (cudaMalloc(&c->local, c->local_bytes) != cudaSuccess || cudaMalloc(&c->seq, sizeof(uint32_t)) != cudaSuccess) {
      kllm_comm_destroy(c);
      return static_cast<int>(cudaErrorMemoryAllocation);
    }
    cudaMemset(c->local, 0, c->local_bytes);
    cudaMemset(c->seq, 0, sizeof(uint32_t));
    fill_table(c, rank, c->local);
    c->connected = world == 1;
  }
  *out = c;
  return 0;
}

int kllm_comm_ipc_handle(kllm_comm* c, unsigned char* out64) {
  if (!c || !out64 || c->backend != KLLM_COMM_PEER) return KLLM_E_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, c->local);
  if (e != cudaSuccess) return static_cast<int>(e);
  std::memcpy(out64, &h, 64);
  return 0;
}

int kllm_comm_connect(kllm_comm* c, const unsigned char* handles) {
  if (!c || c->backend != KLLM_COMM_PEER) return KLLM_E_INVALID;
  if (c->world == 1) return 0;
  if (!handles) return KLLM_E_INVALID;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 64 * r, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(&c->remote[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      std::fprintf(stderr, "kllm tp: rank %d cannot map rank %d's buffer: %s\n", c->rank, r, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    fill_table(c, r, c->remote[r]);
  }
  c->connected = true;
  return 0;
}

int kllm_comm_allreduce_residual(kllm_comm* c, const float* partial, const float* residual, float* out,
                                 int count, void* stream) {
  if (!c || !partial || !out || count <= 0 || count > c->max_count || (count & 3)) return KLLM_E_INVALID;
  if (!c->connected) return KLLM_E_STATE;
  auto s = static_cast<cudaStream_t>(stream);
  if (c->backend == KLLM_COMM_PEER) {
    allreduce_oneshot_kernel<<<1, 1024, 0, s>>>(c->table, c->rank, c->world, partial, residual, out, count,
                                                c->max_count, c->seq);
    count_launch();
    return static_cast<int>(cudaGetLastError());
  }
  // NCCL: reduce in place in a scratch the caller owns (`partial` is the decoder's tp scratch)
  float* buf = const_cast<float*>(partial);
  const int rc = nccl_api()->AllReduce(buf, residual ? buf : out, static_cast<size_t>(count), /*ncclFloat32*/ 7,
                                       /*ncclSum*/ 0, c->nccl, s);
  if (rc != 0) return KLLM_E_COMM;
  count_launch();  // NCCL's kernel
  if (residual != nullptr) {
    residual_add_kernel<<<(count + 255) / 256, 256, 0, s>>>(residual, buf, out, count);
    count_launch();
    return static_cast<int>(cudaGetLastError());
  }
  return 0;
}

int kllm_comm_allreduce(void* comm, float* buf, int count, void* stream) {
  return kllm_comm_allreduce_residual(static_cast<kllm_comm*>(comm), buf, nullptr, buf, count, stream);
}

int kllm_comm_info(const kllm_comm* c, int* world, int* rank, int* backend) {
  if (!c) return KLLM_E_INVALID;
  if (world) *world = c->world;
  if (rank) *rank = c->rank;
  if (backend) *backend = c->backend;
  return 0;
}

}  // extern "C"

namespace kllm {
// Internal (decoder.cu): the tagged exchange areas of every rank for the persistent kernel.
int comm_tagged_areas(kllm_comm* c, unsigned long long** areas8, int* world, int* rank, int* stride) {
  if (!c || c->backend != KLLM_COMM_PEER || !c->connected) return KLLM_E_STATE;
  for (int r = 0; r < kMaxWorld; ++r) areas8[r] = nullptr;
  for (int r = 0; r < c->world; ++r) areas8[r] = tagged_of(c, r == c->rank ? c->local : c->remote[r]);
  *world = c->world, *rank = c->rank, *stride = c->max_count;
  return 0;
}
}  // namespace kllm

extern "C" {

void kllm_comm_destroy(kllm_comm* c) {
  if (!c) return;
  for (int r = 0; r < kMaxWorld; ++r)
    if (c->remote[r]) cudaIpcCloseMemHandle(c->remote[r]);
  if (c->local) cudaFree(c->local);
  if (c->seq) cudaFree(c->seq);
  if (c->nccl && nccl_api()) nccl_api()->CommDestroy(c->nccl);
  delete c;
}

}  // extern "C"
