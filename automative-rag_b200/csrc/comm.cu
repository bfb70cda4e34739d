// comm.cu — the ONE exchange step of each sharded stage as the engine's own kernels over
// NVLink / NVSwitch peer memory.  The reference is single-GPU (nothing to match, vectorstore.py / rerankers.py have no
// collective); correctness here means every rank ends with the bit-identical result the single-GPU path gives.
//
// One process per GPU.  Every rank owns a WIRE BLOCK in its own HBM:
//     [ flags: 2 parities x 16 sources, 128 B apart ][ wire: 2 parities x world slots x slot_bytes ]
// exported once with cudaIpcGetMemHandle and mapped by every peer (rs_comm_export / rs_comm_open; ranks living in one
// process map each other directly with cudaDeviceEnablePeerAccess).  A collective is ONE kernel per rank:
//   1. PUSH   all CTAs copy this rank's contribution into slot [parity][rank] of EVERY rank's wire block — plain
//             coalesced stores to peer-mapped addresses, travelling over NVLink through the switch;
//   2. SIGNAL the last CTA to finish pushing (atomic ticket) fences at system scope and writes the call's sequence
//             number into flag [parity][rank] of every rank (st.release.sys);
//   3. WAIT   every CTA polls its OWN rank's flags (local memory: ld.acquire.sys) until all sources show the sequence
//             number — a peer that never arrives makes the kernel trap after 20 s (RS_COMM_TIMEOUT_MS) instead of
//             hanging the GPU;
//   4. CONSUME straight out of the local wire block: k-way merge of the gathered top-k lists (the same code as
//             rs_topk_merge), element-wise max of score blocks, or a plain copy.
// Parity = sequence number & 1: a rank can be at most one collective ahead of a peer (it cannot finish call s+1
// before that peer has pushed for s+1, i.e. has finished its kernel of call s), so two buffers never collide.
// All CTAs of the kernel must be co-resident (they all wait in step 3): grids are at most one CTA per SM and are
// launched with the cooperative attribute, so the runtime refuses rather than deadlocks.
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "comm.h"
#include "topk_merge.cuh"

namespace rs {

constexpr uint32_t kCommMagic = 0x43435352u;  // "RSCC"
constexpr size_t kCommFlagBytes = 4096;       // 2 parities x 16 sources x 128 B
constexpr int kCommThreads = kMergeThreads;   // 512

struct CommBlob {  // what rs_comm_export hands out (RS_COMM_HANDLE_BYTES = 128)
  uint32_t magic;
  int32_t world, rank, device;
  int64_t pid;
  uint64_t ptr;
  uint64_t slot_bytes;
  cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(CommBlob) <= 128, "CommBlob must fit RS_COMM_HANDLE_BYTES");

struct CommState {
  int device = 0, num_sms = 0;
  int world = 0, rank = 0;
  size_t slot_bytes = 0;
  uint8_t* block = nullptr;  // this rank's wire block
  uint8_t* peer[kCommMaxWorld] = {};
  bool ipc_opened[kCommMaxWorld] = {};
  bool opened = false;
  unsigned long long seq = 0;
  unsigned* done_ctr = nullptr;
};

struct CommView {  // kernel argument
  uint8_t* block[kCommMaxWorld];
  int world, rank;
  unsigned long long slot_bytes;
  unsigned long long seq;
  unsigned long long timeout_ns;  // how long a CTA waits for its peers before it traps
  unsigned* done_ctr;
};

// ----------------------------------------------------------------------------------------------- device side
__device__ __forceinline__ uint8_t* comm_slot(const CommView& c, int on_rank, int from_rank) {
  return c.block[on_rank] + kCommFlagBytes + ((c.seq & 1ull) * (unsigned long long)c.world + from_rank) * c.slot_bytes;
}
__device__ __forceinline__ unsigned long long* comm_flag(const CommView& c, int on_rank, int from_rank) {
  return reinterpret_cast<unsigned long long*>(c.block[on_rank] + ((c.seq & 1ull) * 16ull + from_rank) * 128ull);
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Step 2: every thread of every CTA calls this after its pushes.
__device__ __forceinline__ void comm_signal(const CommView& c) {
  __threadfence_system();  // this thread's peer stores are performed system-wide ...
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // ... and, cumulatively, those of the whole CTA, before the ticket
    const unsigned ticket = atomicAdd(c.done_ctr, 1u);
    if (ticket == gridDim.x - 1) {
      *c.done_ctr = 0u;  // every CTA of this launch has taken its ticket; the next launch is stream-ordered after us
      __threadfence_system();
      for (int p = 0; p < c.world; ++p) st_release_sys(comm_flag(c, p, c.rank), c.seq + 1ull);
    }
  }
}

// Step 3: every thread of every CTA calls this; returns when all sources' contributions are in the local block.
__device__ __forceinline__ void comm_wait(const CommView& c) {
  if ((int)threadIdx.x < c.world) {
    const unsigned long long* f = comm_flag(c, c.rank, (int)threadIdx.x);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(f) < c.seq + 1ull) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > c.timeout_ns) __trap();  // a peer never arrived — fail loudly, do not hang the box
    }
  }
  __syncthreads();
}

// top-k lists: slot = [ids int64 nq*k_in][scores fp32 nq*k_in]
__global__ void __launch_bounds__(kCommThreads, 1)
    comm_allgather_topk_kernel(const CommView c, const float* __restrict__ loc_scores, const int64_t* __restrict__ loc_ids,
                               int nq, int k_in, int k_out, int cap, int prune, float* out_scores, int64_t* out_ids) {
  extern __shared__ __align__(16) uint8_t smem[];
  const size_t n = (size_t)nq * k_in;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (size_t)gridDim.x * blockDim.x;
  for (size_t i = gtid; i < n; i += gthreads) {
    const int64_t id = loc_ids[i];
    const float sc = loc_scores[i];
    for (int p = 0; p < c.world; ++p) {
      int64_t* dst_ids = reinterpret_cast<int64_t*>(comm_slot(c, p, c.rank));
      dst_ids[i] = id;
      reinterpret_cast<float*>(dst_ids + n)[i] = sc;
    }
  }
  comm_signal(c);
  comm_wait(c);
  const int64_t* ids0 = reinterpret_cast<const int64_t*>(comm_slot(c, c.rank, 0));
  const float* sc0 = reinterpret_cast<const float*>(ids0 + n);
  for (int q = blockIdx.x; q < nq; q += gridDim.x)
    merge_one_query(reinterpret_cast<uint64_t*>(smem), sc0, ids0, c.world, k_in, k_out, cap, prune,
                    (int64_t)(c.slot_bytes / 4), (int64_t)(c.slot_bytes / 8), out_scores, out_ids, q);
}

// plain all-gather of `n16` 16-byte units per rank: out = [world][n16]
__global__ void __launch_bounds__(kCommThreads, 1)
    comm_allgather_kernel(const CommView c, const uint4* __restrict__ local, size_t n16, uint4* out) {
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (size_t)gridDim.x * blockDim.x;
  for (size_t i = gtid; i < n16; i += gthreads) {
    const uint4 v = local[i];
    for (int p = 0; p < c.world; ++p) reinterpret_cast<uint4*>(comm_slot(c, p, c.rank))[i] = v;
  }
  comm_signal(c);
  comm_wait(c);
  for (int p = 0; p < c.world; ++p) {
    const uint4* src = reinterpret_cast<const uint4*>(comm_slot(c, c.rank, p));
    for (size_t i = gtid; i < n16; i += gthreads) out[(size_t)p * n16 + i] = __ldcg(src + i);
  }
}

// element-wise max over ranks of fp32 vectors (every candidate has one owner, the others contribute -inf)
__global__ void __launch_bounds__(kCommThreads, 1)
    comm_allreduce_max_kernel(const CommView c, const float* __restrict__ local, size_t n, float* out) {
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (size_t)gridDim.x * blockDim.x;
  for (size_t i = gtid; i < n; i += gthreads) {
    const float v = local[i];
    for (int p = 0; p < c.world; ++p) reinterpret_cast<float*>(comm_slot(c, p, c.rank))[i] = v;
  }
  comm_signal(c);
  comm_wait(c);
  for (size_t i = gtid; i < n; i += gthreads) {
    float m = __ldcg(reinterpret_cast<const float*>(comm_slot(c, c.rank, 0)) + i);
    for (int p = 1; p < c.world; ++p) m = fmaxf(m, __ldcg(reinterpret_cast<const float*>(comm_slot(c, c.rank, p)) + i));
    out[i] = m;
  }
}

// ----------------------------------------------------------------------------------------------- host side
CommState* comm_create(int device, int num_sms) {
  CommState* s = new CommState();
  s->device = device;
  s->num_sms = num_sms;
  return s;
}

static void comm_release(CommState* s) {
  for (int p = 0; p < kCommMaxWorld; ++p) {
    if (s->ipc_opened[p] && s->peer[p]) cudaIpcCloseMemHandle(s->peer[p]);
    s->ipc_opened[p] = false;
    s->peer[p] = nullptr;
  }
  if (s->block) cudaFree(s->block);
  if (s->done_ctr) cudaFree(s->done_ctr);
  s->block = nullptr;
  s->done_ctr = nullptr;
  s->opened = false;
  s->world = 0;
  s->seq = 0;
  cudaGetLastError();
}

void comm_destroy(CommState* s) {
  if (!s) return;
  comm_release(s);
  delete s;
}

bool comm_is_open(const CommState* s) { return s && s->opened; }
int comm_world(const CommState* s) { return s ? s->world : 0; }
int comm_rank(const CommState* s) { return s ? s->rank : 0; }
size_t comm_slot_bytes(const CommState* s) { return s ? s->slot_bytes : 0; }

int comm_export(CommState* s, int world, int rank, size_t slot_bytes, void* out_blob, std::string* err) {
  if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world) {
    *err = "world must be in [1, " + std::to_string(kCommMaxWorld) + "] and rank in [0, world)";
    return -1;
  }
  comm_release(s);
  slot_bytes = (slot_bytes + 255) / 256 * 256;
  if (slot_bytes < 4096) slot_bytes = 4096;
  const size_t total = kCommFlagBytes + (size_t)2 * world * slot_bytes;
  cudaError_t e = cudaMalloc(&s->block, total);
  if (e == cudaSuccess) e = cudaMemset(s->block, 0, total);
  if (e == cudaSuccess) e = cudaMalloc(&s->done_ctr, 128);
  if (e == cudaSuccess) e = cudaMemset(s->done_ctr, 0, 128);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();  // zeroed before any peer can learn the address
  CommBlob blob{};
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&blob.ipc, s->block);
  if (e != cudaSuccess) {
    *err = std::string("wire block allocation / export: ") + cudaGetErrorString(e);
    comm_release(s);
    return -3;
  }
  s->world = world;
  s->rank = rank;
  s->slot_bytes = slot_bytes;
  blob.magic = kCommMagic;
  blob.world = world;
  blob.rank = rank;
  blob.device = s->device;
  blob.pid = (int64_t)getpid();
  blob.ptr = reinterpret_cast<uint64_t>(s->block);
  blob.slot_bytes = slot_bytes;
  memset(out_blob, 0, 128);
  memcpy(out_blob, &blob, sizeof(blob));
  return 0;
}

int comm_open(CommState* s, const void* blobs, std::string* err) {
  if (!s->block) {
    *err = "rs_comm_export must be called first";
    return -1;
  }
  const uint8_t* bp = static_cast<const uint8_t*>(blobs);
  for (int p = 0; p < s->world; ++p) {
    CommBlob b;
    memcpy(&b, bp + (size_t)p * 128, sizeof(b));
    if (b.magic != kCommMagic || b.world != s->world || b.rank != p || b.slot_bytes != s->slot_bytes) {
      *err = "handle " + std::to_string(p) + " does not describe rank " + std::to_string(p) + " of this " +
             std::to_string(s->world) + "-rank exchange (same world and slot size on every rank, handles in rank order)";
      return -1;
    }
    if (p == s->rank) {
      s->peer[p] = s->block;
      continue;
    }
    if (b.pid == (int64_t)getpid()) {
      // ranks of one process: map the peer's allocation directly
      int can = 0;
      cudaDeviceCanAccessPeer(&can, s->device, b.device);
      if (!can) {
        *err = "device " + std::to_string(s->device) + " cannot access peer device " + std::to_string(b.device);
        return -6;
      }
      cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        *err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
        return -6;
      }
      cudaGetLastError();
      s->peer[p] = reinterpret_cast<uint8_t*>(b.ptr);
    } else {
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        *err = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(p) + "): " + cudaGetErrorString(e);
        return -6;
      }
      s->peer[p] = static_cast<uint8_t*>(ptr);
      s->ipc_opened[p] = true;
    }
  }
  s->opened = true;
  s->seq = 0;
  return 0;
}

void comm_close(CommState* s) { comm_release(s); }

static CommView make_view(CommState* s) {
  CommView v{};
  for (int p = 0; p < s->world; ++p) v.block[p] = s->peer[p];
  v.world = s->world;
  v.rank = s->rank;
  v.slot_bytes = s->slot_bytes;
  v.seq = s->seq++;
  // Ranks reach a collective at different times (host-side skew); 20 s covers that, a dead peer still ends in a trap
  // (= a CUDA error on this rank) instead of a hung device.  RS_COMM_TIMEOUT_MS overrides.
  static const unsigned long long timeout_ms = getenv("RS_COMM_TIMEOUT_MS") ? strtoull(getenv("RS_COMM_TIMEOUT_MS"), nullptr, 10) : 20000ull;
  v.timeout_ns = timeout_ms * 1000000ull;
  v.done_ctr = s->done_ctr;
  return v;
}

template <typename... Args>
static cudaError_t launch_coop(void (*kernel)(Args...), int grid, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kCommThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

int comm_allgather_topk(CommState* s, const float* loc_scores, const int64_t* loc_ids, int nq, int k_in, int k_out,
                        float* out_scores, int64_t* out_ids, cudaStream_t stream, std::string* err) {
  if (!s->opened) {
    *err = "no exchange is open on this handle (rs_comm_export + rs_comm_open)";
    return -1;
  }
  const size_t need = (size_t)nq * k_in * 12;
  if (need > s->slot_bytes) {
    *err = "nq * k_in * 12 = " + std::to_string(need) + " bytes exceed the exchange's slot of " + std::to_string(s->slot_bytes);
    return -1;
  }
  if ((long long)s->world * k_in > 16384 || k_out > 16384) {
    *err = "world * k_in and k_out must be <= 16384";
    return -2;
  }
  int cap, prune;
  merge_plan(s->world, k_in, k_out, &cap, &prune);
  const size_t smem = (size_t)cap * 8;
  cudaError_t e = cudaFuncSetAttribute(comm_allgather_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    const int grid = nq < s->num_sms ? nq : s->num_sms;
    e = launch_coop(comm_allgather_topk_kernel, grid, smem, stream, (const CommView)make_view(s), loc_scores, loc_ids, nq, k_in,
                    k_out, cap, prune, out_scores, out_ids);
  }
  if (e != cudaSuccess) {
    *err = std::string("comm_allgather_topk_kernel launch: ") + cudaGetErrorString(e);
    return -3;
  }
  return 0;
}

static int grid_for_bytes(const CommState* s, size_t bytes_per_rank) {
  // enough CTAs to keep NVLink busy for large blocks, one CTA for a few KB (latency-bound anyway)
  size_t g = (bytes_per_rank * (size_t)s->world + 65535) / 65536;
  if (g < 1) g = 1;
  if (g > (size_t)s->num_sms) g = (size_t)s->num_sms;
  return (int)g;
}

int comm_allgather(CommState* s, const void* local, size_t bytes, void* out, cudaStream_t stream, std::string* err) {
  if (!s->opened) {
    *err = "no exchange is open on this handle (rs_comm_export + rs_comm_open)";
    return -1;
  }
  if (bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(local) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) {
    *err = "rs_allgather: buffers must be 16-byte aligned and bytes a multiple of 16";
    return -1;
  }
  if (bytes > s->slot_bytes) {
    *err = std::to_string(bytes) + " bytes exceed the exchange's slot of " + std::to_string(s->slot_bytes);
    return -1;
  }
  cudaError_t e = launch_coop(comm_allgather_kernel, grid_for_bytes(s, bytes), 0, stream, (const CommView)make_view(s),
                              static_cast<const uint4*>(local), bytes / 16, static_cast<uint4*>(out));
  if (e != cudaSuccess) {
    *err = std::string("comm_allgather_kernel launch: ") + cudaGetErrorString(e);
    return -3;
  }
  return 0;
}

int comm_allreduce_max(CommState* s, const float* local, size_t n, float* out, cudaStream_t stream, std::string* err) {
  if (!s->opened) {
    *err = "no exchange is open on this handle (rs_comm_export + rs_comm_open)";
    return -1;
  }
  if (n * 4 > s->slot_bytes) {
    *err = std::to_string(n * 4) + " bytes exceed the exchange's slot of " + std::to_string(s->slot_bytes);
    return -1;
  }
  cudaError_t e = launch_coop(comm_allreduce_max_kernel, grid_for_bytes(s, n * 4), 0, stream, (const CommView)make_view(s), local,
                              n, out);
  if (e != cudaSuccess) {
    *err = std::string("comm_allreduce_max_kernel launch: ") + cudaGetErrorString(e);
    return -3;
  }
  return 0;
}

}  // namespace rs
