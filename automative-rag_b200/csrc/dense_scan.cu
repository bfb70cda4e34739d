// dense_scan.cu — single-query exact top-k scan over a row-major fp16/bf16 corpus.
//
// Replaces the arithmetic behind QdrantStore.similarity_search_with_score
// (reference src/core/query/retrieval/vectorstore.py:166-214; Qdrant cosine search with a
// payload filter).  One query is a GEMV over the whole corpus, so the kernel is HBM-bound by
// construction: 2*d bytes per passing row, ~2 flops per byte.
//
// Shape of the kernel (B200: 148 SMs, one persistent CTA per SM):
//   * warp 8 is the TMA producer.  It walks the CTA's tiles in order (tile = TILE_ROWS
//     consecutive rows = one contiguous byte range, 16 KB at d = 1024): lane 0 waits for the
//     ring slot, then ONE cp.async.bulk moves a fully passing tile, or lane i issues the run of
//     passing rows starting at row i.  Rows that fail the filter are never read from HBM.
//   * warps 0..7 are consumers; warp w takes tiles w, w+8, ...  A lane reads 16-byte vectors
//     (LDS.128, conflict free), converts to fp32, FMAs against the query held in registers, and
//     the warp butterfly-reduces.  The ring is as deep as shared memory allows (13 x 16 KB =
//     208 KB at k <= 128): Little's law for ~6.5 TB/s x ~2 us loaded latency needs ~90 KB in
//     flight per SM, and a slot is out of flight while its tile is being reduced.
//   * scores become order-preserving u64 keys and go through the CTA's TopKBuffer; a barrier
//     every few rounds decides whether to compact.
//   * every CTA writes its sorted top-k to the workspace; the last CTA to finish (atomic
//     ticket) merges the grid's lists and writes the final (score, id) pairs — no second launch.
#include "common.cuh"
#include "kernels.h"
#include "topk_buffer.cuh"

namespace rs {

constexpr int kScanConsumerWarps = 8;
constexpr int kScanMaxStages = 16;
constexpr int kScanConsumerThreads = kScanConsumerWarps * 32;
constexpr int kScanThreads = kScanConsumerThreads + 32;
constexpr int kConsumerBar = 1;  // named barrier id for the 256 consumer threads

__device__ __forceinline__ uint32_t tile_mask_bits(const uint32_t* __restrict__ mask, int64_t tile, int tile_rows,
                                                   int64_t n) {
  int64_t row0 = tile * tile_rows;
  int64_t left = n - row0;
  uint32_t in_range = left >= tile_rows ? (tile_rows == 32 ? 0xFFFFFFFFu : ((1u << tile_rows) - 1u))
                                        : ((1u << (int)left) - 1u);
  if (mask == nullptr) return in_range;
  uint32_t w = __ldg(mask + (row0 >> 5));  // tile_rows is a power of two <= 32: never straddles a word
  return (w >> (row0 & 31)) & in_range;
}

template <typename T, int NCH>
__global__ void __launch_bounds__(kScanThreads, 1) dense_scan_kernel(const ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tile_rows = p.tile_rows;
  const int S = p.stages;
  const uint32_t row_bytes = (uint32_t)p.d * 2u;
  const uint32_t tile_bytes = row_bytes * tile_rows;

  // shared memory carve-up
  uint8_t* stage_base = smem;                                                  // S * tile_bytes
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem + (size_t)S * tile_bytes);  // [C]
  uint64_t* full_bar = keys + p.buf_cap;
  uint64_t* empty_bar = full_bar + kScanMaxStages;
  uint64_t* thr = empty_bar + kScanMaxStages;
  int* cnt = reinterpret_cast<int*>(thr + 1);
  int* s_flag = cnt + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int64_t num_tiles = (p.n + tile_rows - 1) / tile_rows;
  // tiles of this CTA: blockIdx.x, blockIdx.x + grid, ...  (t-th tile -> global tile blockIdx.x + t * grid)
  const int64_t my_tiles = (num_tiles > blockIdx.x) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t rounds = (my_tiles + kScanConsumerWarps - 1) / kScanConsumerWarps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const uint32_t* mask = p.mask;

  if (warp == kScanConsumerWarps) {
    // ------------------------------------------------------------------ TMA producer warp
    // The warp walks this CTA's tiles IN ORDER, converged: lane 0 waits for the ring slot, then
    // either lane 0 issues one bulk copy for the whole tile or lane i issues the run of passing
    // rows that starts at row i.  Filter bits are fetched 32 tiles at a time (one per lane), one
    // batch ahead of their use.
    const uint64_t pol = policy_evict_first();
    const uint32_t all_bits = tile_rows == 32 ? 0xFFFFFFFFu : ((1u << tile_rows) - 1u);
    auto fetch_bits = [&](int64_t t) -> uint32_t {
      return t < my_tiles ? tile_mask_bits(mask, blockIdx.x + t * gridDim.x, tile_rows, p.n) : 0u;
    };
    uint32_t bits_cur = fetch_bits(lane);
    for (int64_t tb = 0; tb < my_tiles; tb += 32) {
      const uint32_t bits_nxt = fetch_bits(tb + 32 + lane);
      const int lim = (int)min((int64_t)32, my_tiles - tb);
      for (int i = 0; i < lim; ++i) {
        const int64_t t = tb + i;
        const uint32_t bits = __shfl_sync(0xFFFFFFFFu, bits_cur, i);
        const int stage = (int)(t % S);
        const uint32_t par = (uint32_t)((t / S) & 1);
        if (lane == 0) mbar_wait(&empty_bar[stage], par ^ 1u);
        __syncwarp();
        const int64_t tile = blockIdx.x + t * gridDim.x;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.corpus) + (size_t)tile * tile_bytes;
        uint8_t* dst = stage_base + (size_t)stage * tile_bytes;
        if (bits == 0u) {
          if (lane == 0) mbar_arrive(&full_bar[stage]);
        } else if (bits == all_bits) {
          if (lane == 0) {
            mbar_arrive_expect_tx(&full_bar[stage], tile_bytes);
            bulk_g2s(dst, src, tile_bytes, &full_bar[stage], pol);
          }
        } else {
          if (lane == 0) mbar_arrive_expect_tx(&full_bar[stage], row_bytes * __popc(bits));
          __syncwarp();
          const bool run_start = ((bits >> lane) & 1u) && (lane == 0 || !((bits >> (lane - 1)) & 1u));
          if (run_start) {
            const int run = __ffs(~(bits >> lane)) - 1;  // consecutive passing rows from this one
            bulk_g2s(dst + (size_t)lane * row_bytes, src + (size_t)lane * row_bytes, row_bytes * run, &full_bar[stage],
                     pol);
          }
        }
      }
      bits_cur = bits_nxt;
    }
  } else {
    // ------------------------------------------------------------------ consumer warps
    TopKBuffer buf{keys, thr, cnt, p.buf_cap, p.k, (int)threadIdx.x, kScanConsumerThreads, kConsumerBar};
    buf.init();

    // query -> fp32 registers; lane owns elements c*256 + lane*8 + j
    float q[NCH * 8];
    const T* qp = reinterpret_cast<const T*>(p.query);
    float qss = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      int e0 = c * 256 + lane * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = (e0 + j < p.d) ? Cvt<T>::to_float(qp[e0 + j]) : 0.f;
        q[c * 8 + j] = v;
        qss = fmaf(v, v, qss);
      }
    }
    float q_scale = 1.f;
    if (p.metric == 1) {
      qss = warp_sum(qss);
      q_scale = qss > 0.f ? rsqrtf(qss) : 0.f;
      // rsqrtf is approximate (2 ulp); one Newton step makes the scale fp32-accurate
      if (qss > 0.f) q_scale = q_scale * (1.5f - 0.5f * qss * q_scale * q_scale);
    }
    const float* inv_norm = p.inv_norm;
    const int nvec = p.d >> 3;  // 16-byte vectors per row
    // appends per round <= warps * tile_rows; the buffer tolerates C/2 between checks
    const int rounds_per_check = max(1, (p.buf_cap >> 1) / (kScanConsumerWarps * tile_rows));

    int64_t t = warp;
    for (int64_t r = 0; r < rounds; ++r, t += kScanConsumerWarps) {
      if (t < my_tiles) {
        const int stage = (int)(t % S);
        const uint32_t par = (uint32_t)((t / S) & 1);
        const uint8_t* my_stage = stage_base + (size_t)stage * tile_bytes;
        const int64_t tile = blockIdx.x + t * gridDim.x;
        const uint32_t bits = tile_mask_bits(mask, tile, tile_rows, p.n);
        const int64_t row0 = tile * tile_rows;
        float inv = 1.f;
        if (inv_norm != nullptr && lane < tile_rows && ((bits >> lane) & 1u)) inv = __ldg(inv_norm + row0 + lane);
        mbar_wait(&full_bar[stage], par);
        float my_score = 0.f;
        if (bits != 0u) {
          for (int i0 = 0; i0 < tile_rows; i0 += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = i0 + u;
              if (i < tile_rows && ((bits >> i) & 1u)) {
                const uint4* rowp = reinterpret_cast<const uint4*>(my_stage + (size_t)i * row_bytes);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                  const int v = c * 32 + lane;
                  if (v < nvec) acc[u] = dot8<T>(rowp[v], &q[c * 8], acc[u]);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float s = warp_sum(acc[u]);
              if (lane == i0 + u) my_score = s;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);  // slot may be refilled
        // lanes < tile_rows hold one row score each
        const bool live = lane < tile_rows && ((bits >> lane) & 1u);
        const float score = my_score * inv * q_scale;
        const uint64_t key = make_key(score, (uint32_t)(row0 + lane));
        buf.warp_append(live && key > buf.threshold(), key);
      }
      if ((r + 1) % rounds_per_check == 0) buf.maybe_compact();
    }
    buf.compact();  // final: keys[0..k) sorted descending (0 = empty)

    // ------------------------------------------------------------------ cross-CTA merge
    uint64_t* ws = p.ws_keys + (size_t)blockIdx.x * p.k;
    for (int i = threadIdx.x; i < p.k; i += kScanConsumerThreads) ws[i] = keys[i];
    __threadfence();
    named_bar_sync(kConsumerBar, kScanConsumerThreads);
    if (threadIdx.x == 0) {
      unsigned ticket = atomicAdd(p.ticket, 1u);
      *s_flag = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    named_bar_sync(kConsumerBar, kScanConsumerThreads);
    if (*s_flag) {
      __threadfence();
      // The buffer already holds this CTA's own top-k with the matching threshold; stream the
      // other CTAs' sorted lists through it.  Thread t walks list t (+256, ...) J keys at a
      // time; a list is abandoned at its first key <= threshold (lists are sorted).
      const int J = max(1, (buf.C >> 1) / kScanConsumerThreads);
      for (int base = 0; base < (int)gridDim.x; base += kScanConsumerThreads) {
        const int list = base + threadIdx.x;
        bool active = list < (int)gridDim.x && list != (int)blockIdx.x;
        const uint64_t* lp = p.ws_keys + (size_t)list * p.k;
        for (int j0 = 0; j0 < p.k; j0 += J) {
          for (int j = j0; j < min(j0 + J, p.k); ++j) {
            uint64_t key = 0ull;
            if (active) {
              key = __ldcg(lp + j);
              if (key <= buf.threshold()) active = false;
            }
            buf.warp_append(active, key);
          }
          buf.maybe_compact();
        }
      }
      buf.compact();
      for (int i = threadIdx.x; i < p.k; i += kScanConsumerThreads) {
        uint64_t key = keys[i];
        if (key == 0ull) {
          p.out_scores[i] = -INFINITY;
          p.out_ids[i] = -1;
        } else {
          p.out_scores[i] = key_score(key);
          p.out_ids[i] = p.id_base + (int64_t)key_row(key);
        }
      }
      if (threadIdx.x == 0) *p.ticket = 0u;  // ready for the next launch on this stream
    }
  }
}

template <typename T>
static cudaError_t launch_scan_t(const ScanParams& p, int grid, size_t smem, cudaStream_t stream) {
  const int nch = (p.d + 255) / 256;
#define RS_SCAN_CASE(N)                                                                                      \
  {                                                                                                          \
    cudaError_t e = cudaFuncSetAttribute(dense_scan_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                         \
    if (e != cudaSuccess) return e;                                                                          \
    dense_scan_kernel<T, N><<<grid, kScanThreads, smem, stream>>>(p);                                        \
    return cudaGetLastError();                                                                               \
  }
  if (nch <= 1) RS_SCAN_CASE(1)
  if (nch <= 2) RS_SCAN_CASE(2)
  if (nch <= 4) RS_SCAN_CASE(4)
  if (nch <= 8) RS_SCAN_CASE(8)
  RS_SCAN_CASE(16)
#undef RS_SCAN_CASE
}

int scan_tile_rows(int d) {
  int rows = 16384 / (d * 2);
  int tr = 1;
  while (tr * 2 <= rows && tr < 32) tr <<= 1;
  return tr;
}

static size_t scan_fixed_smem(int k) {
  return (size_t)TopKBuffer::capacity_for(k) * 8 + 2 * kScanMaxStages * 8 + 8 + 16;
}

int scan_stages(int d, int k) {
  const size_t tile_bytes = (size_t)scan_tile_rows(d) * d * 2;
  const size_t budget = 227 * 1024 - 1024;  // leave 1 KB for the runtime's reserved shared memory
  int s = (int)((budget - scan_fixed_smem(k)) / tile_bytes);
  return s > kScanMaxStages ? kScanMaxStages : (s < 2 ? 2 : s);
}

size_t scan_smem_bytes(int d, int k) {
  const size_t tile_bytes = (size_t)scan_tile_rows(d) * d * 2;
  return (size_t)scan_stages(d, k) * tile_bytes + scan_fixed_smem(k);
}

cudaError_t launch_dense_scan(ScanParams p, int dtype, int num_sms, cudaStream_t stream) {
  p.tile_rows = scan_tile_rows(p.d);
  p.stages = scan_stages(p.d, p.k);
  p.buf_cap = TopKBuffer::capacity_for(p.k);
  const int64_t num_tiles = (p.n + p.tile_rows - 1) / p.tile_rows;
  int grid = (int)(num_tiles < (int64_t)num_sms ? (num_tiles > 0 ? num_tiles : 1) : num_sms);
  const size_t smem = scan_smem_bytes(p.d, p.k);
  if (dtype == 0) return launch_scan_t<__half>(p, grid, smem, stream);
  return launch_scan_t<__nv_bfloat16>(p, grid, smem, stream);
}

}  // namespace rs
