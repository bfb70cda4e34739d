// topk_merge.cu — small integer/sort kernels around the two scoring stages:
//   * topk_merge_kernel: merge per-shard top-k lists after the one all-gather
//   * rerank_postprocess_kernel: the sort / min-max / blend / [:top_k] tail of
//     ColBERTReranker.rerank (reference src/core/query/llm/rerankers.py:302-343,377-380)
//   * filter_mask_kernel: _build_filter predicate (vectorstore.py:216-276) over columnar
//     metadata -> bit-packed row mask
#include <math_constants.h>

#include "common.cuh"
#include "kernels.h"
#include "topk_merge.cuh"

namespace rs {


// One CTA per query.  The candidates of the nlists sorted lists are loaded into shared memory as
// (orderable score, position) keys and bitonic-sorted; runs of equal score are then ordered by
// ascending id so the result obeys the ABI's (score desc, id asc) total order even when shard
// id ranges interleave.
//
// Only candidates that can still be in the answer are sorted.  With r = ceil(k_out / nlists) and
// b_l = the smallest of list l's first r entries (its r-th best, the lists being sorted), the
// ceil(k_out / r) lists with the largest b_l hold r entries each that score at least the smallest
// of those b_l, T — k_out or more candidates — so nothing below T is needed (everything equal to T
// is kept).  148 lists x 100 (a batched dense search) shrink from a
// 16384-key sort to a few hundred keys.
//
// The body lives in topk_merge.cuh (merge_one_query): the fused gather-and-merge kernel of the multi-GPU exchange
// (comm.cu) runs the same code on lists that peers wrote into this GPU's wire buffer.
__global__ void __launch_bounds__(kMergeThreads) topk_merge_kernel(const float* scores, const int64_t* ids, int nlists,
                                                                   int nq, int k_in, int k_out, int cap, int prune,
                                                                   int64_t sstride, int64_t istride, float* out_scores,
                                                                   int64_t* out_ids, const uint32_t* bound,
                                                                   int bound_groups, const float* bound_scale) {
  extern __shared__ __align__(16) uint8_t smem[];
  // bound: [bound_groups, nq] orderable scores, each standing for enough rows that the minimum over the groups has
  // k_out candidates at or above it (dense_tc5.cu's cross-range words); any zero word = no bound.  The list scores
  // are those scores times bound_scale[q] (>= 0; the same float product the lists were written with, so the order of
  // a score and the bound survives the rounding).
  uint32_t hint = 0u;
  if (bound) {
    hint = 0xFFFFFFFFu;
    for (int g = 0; g < bound_groups; ++g) hint = min(hint, __ldcg(bound + (size_t)g * nq + blockIdx.x));
    if (hint != 0u && bound_scale) hint = f32_orderable(orderable_f32(hint) * __ldcg(bound_scale + blockIdx.x));
  }
  merge_one_query(reinterpret_cast<uint64_t*>(smem), scores, ids, nlists, k_in, k_out, cap, prune, sstride, istride,
                  out_scores, out_ids, (int)blockIdx.x, hint);
}

cudaError_t launch_topk_merge(const float* scores, const int64_t* ids, int nlists, int nq, int k_in, int k_out,
                              int64_t score_list_stride, int64_t id_list_stride, float* out_scores, int64_t* out_ids,
                              cudaStream_t stream, const uint32_t* bound, int bound_groups, const float* bound_scale) {
  if (score_list_stride == 0) score_list_stride = (int64_t)nq * k_in;
  if (id_list_stride == 0) id_list_stride = (int64_t)nq * k_in;
  int cap, prune;
  merge_plan(nlists, k_in, k_out, &cap, &prune);
  const size_t smem = (size_t)cap * 8;
  cudaError_t e = cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  topk_merge_kernel<<<nq, kMergeThreads, smem, stream>>>(scores, ids, nlists, nq, k_in, k_out, cap, prune,
                                                         score_list_stride, id_list_stride, out_scores, out_ids,
                                                         bound, bound_groups, bound_scale);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------- rerank tail
// One CTA per query row.  Stable descending order == sort by key (orderable score, ~position):
// on equal scores the earlier position comes first, exactly Python's stable
// sorted(..., key=score, reverse=True) (rerankers.py:377-380).  With a second score vector the
// reference first orders by the ColBERT score, then min-max normalises both vectors, blends and
// stably re-sorts THAT order (rerankers.py:298-339) — so blended ties keep ColBERT order, which
// is why the second pass keys on the rank from the first pass, not on the input index.
__global__ void __launch_bounds__(kMergeThreads) rerank_postprocess_kernel(const float* __restrict__ scores,
                                                                           const float* __restrict__ other, int n,
                                                                           int cap, float w_a, float w_b, int top_k,
                                                                           int32_t* __restrict__ out_idx,
                                                                           float* __restrict__ out_scores) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem);  // [cap]
  int* perm = reinterpret_cast<int*>(keys + cap);      // [cap] rank after pass 1 -> input index
  float* red = reinterpret_cast<float*>(perm + cap);   // [4 * 32]
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* a = scores + (size_t)q * n;
  const float* b = other ? other + (size_t)q * n : nullptr;
  for (int i = tid; i < cap; i += kMergeThreads) keys[i] = (i < n) ? make_key(a[i], (uint32_t)i) : 0ull;
  bitonic_sort_desc(keys, cap, tid, kMergeThreads, kMergeBar);
  if (b) {
    float amin = CUDART_INF_F, amax = -CUDART_INF_F, bmin = CUDART_INF_F, bmax = -CUDART_INF_F;
    for (int i = tid; i < n; i += kMergeThreads) {
      float x = a[i], y = b[i];
      amin = fminf(amin, x);
      amax = fmaxf(amax, x);
      bmin = fminf(bmin, y);
      bmax = fmaxf(bmax, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      amin = fminf(amin, __shfl_xor_sync(0xFFFFFFFFu, amin, o));
      amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
      bmin = fminf(bmin, __shfl_xor_sync(0xFFFFFFFFu, bmin, o));
      bmax = fmaxf(bmax, __shfl_xor_sync(0xFFFFFFFFu, bmax, o));
    }
    if (lane == 0) {
      red[warp] = amin;
      red[32 + warp] = amax;
      red[64 + warp] = bmin;
      red[96 + warp] = bmax;
    }
    __syncthreads();
    for (int w = 0; w < kMergeThreads / 32; ++w) {
      amin = fminf(amin, red[w]);
      amax = fmaxf(amax, red[32 + w]);
      bmin = fminf(bmin, red[64 + w]);
      bmax = fmaxf(bmax, red[96 + w]);
    }
    const float ar = amax - amin, br = bmax - bmin;
    for (int r = tid; r < cap; r += kMergeThreads) {
      const uint64_t key = keys[r];
      uint64_t key2 = 0ull;
      int src = -1;
      if (key != 0ull) {
        src = (int)key_row(key);
        // (score - min) / range, all-equal -> 1.0 (rerankers.py:302-310, :319-327); blend :330-333
        const float na = ar > 0.f ? (a[src] - amin) / ar : 1.f;
        const float nb = br > 0.f ? (b[src] - bmin) / br : 1.f;
        key2 = make_key(w_a * na + w_b * nb, (uint32_t)r);
      }
      perm[r] = src;
      keys[r] = key2;  // slot r is read and written by this thread only
    }
    bitonic_sort_desc(keys, cap, tid, kMergeThreads, kMergeBar);
  }
  for (int i = tid; i < top_k; i += kMergeThreads) {
    const uint64_t key = (i < cap) ? keys[i] : 0ull;
    if (key == 0ull) {
      out_idx[(size_t)q * top_k + i] = -1;
      out_scores[(size_t)q * top_k + i] = -CUDART_INF_F;
    } else {
      const int r = (int)key_row(key);
      out_idx[(size_t)q * top_k + i] = b ? perm[r] : r;
      out_scores[(size_t)q * top_k + i] = key_score(key);
    }
  }
}

cudaError_t launch_rerank_postprocess(const float* scores, const float* other, int nq, int n, float w_a, float w_b,
                                      int top_k, int32_t* out_idx, float* out_scores, cudaStream_t stream) {
  int cap = 64;
  while (cap < n) cap <<= 1;
  const size_t smem = (size_t)cap * 12 + 128 * 4;
  cudaError_t e =
      cudaFuncSetAttribute(rerank_postprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  rerank_postprocess_kernel<<<nq, kMergeThreads, smem, stream>>>(scores, other, n, cap, w_a, w_b, top_k, out_idx,
                                                                 out_scores);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------- filter -> bitmask
// Thread i tests row i against every clause (AND over clauses, OR inside a clause); a warp
// ballot packs 32 rows into one mask word, so the column reads and the mask write are both
// fully coalesced.  Pure HBM-bound integer work: 4 bytes per row per clause in, 1 bit out.
constexpr int kFilterMaxClauses = 16;
struct FilterClauses {
  const int32_t* col[kFilterMaxClauses];
};

__global__ void __launch_bounds__(256) filter_mask_kernel(FilterClauses fc, int nclauses,
                                                          const int32_t* __restrict__ values,
                                                          const int32_t* __restrict__ val_offsets,
                                                          const uint32_t* __restrict__ tombstone, int64_t n,
                                                          uint32_t* __restrict__ out_mask) {
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nwords;
       w += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t row = w * 32 + lane;
    bool pass = row < n;
    for (int c = 0; c < nclauses && __any_sync(0xFFFFFFFFu, pass); ++c) {
      const int v0 = __ldg(val_offsets + c), v1 = __ldg(val_offsets + c + 1);
      bool hit = false;
      if (pass) {
        const int32_t x = __ldg(fc.col[c] + row);
        for (int v = v0; v < v1; ++v) hit |= (x == __ldg(values + v));
      }
      pass = pass && hit;
    }
    uint32_t word = __ballot_sync(0xFFFFFFFFu, pass);
    if (tombstone) word &= ~__ldg(tombstone + w);
    if (lane == 0) out_mask[w] = word;
  }
}

// ------------------------------------------------------------------------- candidate ownership (sharded MaxSim)
// Documents are owned round-robin (owner = id % world, local index = id / world): global candidate ids -> this
// rank's local document indices, -1 for candidates another rank owns and for padding (an index outside the collection
// is an empty document the MaxSim kernels skip and score -inf).  One launch instead of six element-wise torch ops.
template <typename T>
__global__ void owned_candidates_kernel(const T* __restrict__ cand, int64_t n, int world, int rank, int64_t pool,
                                        int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t id = (int64_t)cand[i];
  if (id >= 0 && pool > 0) id %= pool;  // config 5: document slot = row id mod pool size
  out[i] = (id >= 0 && id % world == rank) ? (int32_t)(id / world) : -1;
}

cudaError_t launch_owned_candidates(const void* cand, int is_i64, int64_t n, int world, int rank, int64_t pool,
                                    int32_t* out, cudaStream_t stream) {
  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads);
  if (is_i64)
    owned_candidates_kernel<int64_t><<<blocks, threads, 0, stream>>>(static_cast<const int64_t*>(cand), n, world, rank, pool, out);
  else
    owned_candidates_kernel<int32_t><<<blocks, threads, 0, stream>>>(static_cast<const int32_t*>(cand), n, world, rank, pool, out);
  return cudaGetLastError();
}

cudaError_t launch_filter_mask(const int32_t* const* cols_host, int nclauses, const int32_t* values_dev,
                               const int32_t* val_offsets_dev, const uint32_t* tombstone, int64_t n,
                               uint32_t* out_mask, int num_sms, cudaStream_t stream) {
  FilterClauses fc{};
  for (int c = 0; c < nclauses; ++c) fc.col[c] = cols_host[c];
  const int64_t nwords = (n + 31) >> 5;
  int64_t blocks = (nwords + 7) / 8;
  const int64_t cap = (int64_t)num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  filter_mask_kernel<<<(int)blocks, 256, 0, stream>>>(fc, nclauses, values_dev, val_offsets_dev, tombstone, n, out_mask);
  return cudaGetLastError();
}

}  // namespace rs

// ------------------------------------------------------------------------- document list -> packed tokens
// ColBERTReranker._compute_maxsim_scores receives its documents as a Python LIST of separately allocated [Ld_i, D]
// tensors (rerankers.py:215-217); rs_maxsim wants one packed [sum Ld, D] buffer.  One launch copies (and converts to
// the compute dtype) every document: CTA (x, i) handles rows 32x .. 32x+31 of document i, 16 bytes of source per
// thread step, fully coalesced.  Sources are device pointers from a table (documents already on the GPU) or offsets
// into one staged upload (documents that came from host memory).
namespace rs {

__device__ __forceinline__ float load_as_float(const void* p, int dtype, size_t i) {
  if (dtype == 2) return reinterpret_cast<const float*>(p)[i];
  if (dtype == 0) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void store_from_float(void* p, int dtype, size_t i, float v) {
  if (dtype == 2) reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == 0) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256) gather_docs_kernel(const void* const* __restrict__ ptrs,
                                                          const uint8_t* __restrict__ staged,
                                                          const int64_t* __restrict__ src_off,
                                                          const int32_t* __restrict__ offsets, int d, int src_dtype,
                                                          int dst_dtype, void* __restrict__ dst) {
  const int doc = blockIdx.y;
  const int row0 = offsets[doc], rows = offsets[doc + 1] - row0;
  const int r_begin = blockIdx.x * 32;
  if (r_begin >= rows) return;
  const int r_end = min(rows, r_begin + 32);
  const void* src = ptrs ? ptrs[doc] : static_cast<const void*>(staged + src_off[doc]);
  const size_t e0 = (size_t)r_begin * d, e1 = (size_t)r_end * d;
  const size_t esz_s = src_dtype == 2 ? 4 : 2, esz_d = dst_dtype == 2 ? 4 : 2;
  uint8_t* out = static_cast<uint8_t*>(dst) + (size_t)row0 * d * esz_d;
  if (src_dtype == dst_dtype && ((reinterpret_cast<uintptr_t>(src) | (size_t)d * esz_s) & 15u) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(src) + e0 * esz_s);
    uint4* d4 = reinterpret_cast<uint4*>(out + e0 * esz_d);
    const size_t n16 = (e1 - e0) * esz_s / 16;
    for (size_t i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = s4[i];
  } else {
    for (size_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) store_from_float(out, dst_dtype, i, load_as_float(src, src_dtype, i));
  }
}

cudaError_t launch_gather_docs(const void* const* ptrs_dev, const uint8_t* staged_dev, const int64_t* src_off_dev,
                               const int32_t* offsets_dev, int nd, int max_len, int d, int src_dtype, int dst_dtype,
                               void* dst, cudaStream_t stream) {
  if (nd <= 0 || max_len <= 0) return cudaSuccess;
  dim3 grid((max_len + 31) / 32, nd);
  gather_docs_kernel<<<grid, 256, 0, stream>>>(ptrs_dev, staged_dev, src_off_dev, offsets_dev, d, src_dtype, dst_dtype, dst);
  return cudaGetLastError();
}

}  // namespace rs
