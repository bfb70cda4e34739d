// topk_merge.cuh — merge of per-shard top-k lists for ONE query by one CTA of kMergeThreads threads.
// Shared by topk_merge_kernel (topk_merge.cu) and the fused gather-and-merge kernel of the multi-GPU exchange
// (comm.cu), where the lists were written by peer GPUs into this GPU's wire buffer while the kernel was already
// running: every input read therefore goes through ld.global.cg (L2, the point of coherence for peer writes) and
// never through the non-coherent read-only path.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace rs {

constexpr int kMergeThreads = 512;
constexpr int kMergeBar = 1;

// capacity of the key buffer (a power of two >= nlists * k_in) and whether the pruning pre-pass pays
inline void merge_plan(int nlists, int k_in, int k_out, int* cap_out, int* prune_out) {
  int cap = 64;
  while (cap < nlists * k_in) cap <<= 1;
  // pruning pays once the full sort is large and needs the lists' r-th entries (r <= k_in) and a small sort of them
  const int r = (k_out + nlists - 1) / nlists;
  *cap_out = cap;
  *prune_out = (cap >= 2048 && nlists >= 2 && nlists <= 2048 && r <= k_in) ? 1 : 0;
}

// t0_hint: orderable score known to have >= k_out candidates at or above it, or 0.
// scores / ids: list l of query q at scores + l * sstride + q * k_in (ids alike with istride); entries with id < 0 are
// padding.  Output [k_out] at out_* + q * k_out in (score desc, id asc) order, padded with (-inf, -1).
// All kMergeThreads threads of the CTA call this; `keys` is shared memory for `cap` u64 keys.
__device__ __forceinline__ void merge_one_query(uint64_t* keys, const float* scores, const int64_t* ids, int nlists, int k_in,
                                                int k_out, int cap, int prune, int64_t sstride, int64_t istride,
                                                float* out_scores, int64_t* out_ids, int q, uint32_t t0_hint = 0u) {
  __shared__ int s_cnt;
  __shared__ uint32_t s_t0;
  const int tid = threadIdx.x;
  const int m = nlists * k_in;
  auto id_at = [&](int i) -> int64_t {
    const int list = i / k_in, j = i - list * k_in;
    return __ldcg(ids + (size_t)list * istride + (size_t)q * k_in + j);
  };
  auto key_of = [&](int i) -> uint64_t {  // key of candidate i (list-major position), 0 when the slot is empty
    const int list = i / k_in, j = i - list * k_in;
    const size_t in_list = (size_t)q * k_in + j;
    if (__ldcg(ids + (size_t)list * istride + in_list) < 0) return 0ull;
    return make_key(__ldcg(scores + (size_t)list * sstride + in_list), (uint32_t)i);
  };

  named_bar_sync(kMergeBar, kMergeThreads);  // the previous query of this CTA (persistent callers) is done with keys[]
  // orderable score below which a candidate cannot be in the answer (0 = keep everything).  A caller that already
  // knows such a bound passes it as t0_hint (the batched dense kernel: its cross-range threshold; its lists are in no
  // particular order, which the bound computed below does not exploit well).
  uint32_t t0 = t0_hint;
  if (prune && t0_hint == 0u) {
    const int r = (k_out + nlists - 1) / nlists;  // <= k_in
    const int need = (k_out + r - 1) / r;         // lists whose r-th entries bound the answer (<= nlists)
    int n2 = 64;
    while (n2 < nlists) n2 <<= 1;                 // <= cap
    // per list: the smallest of its first r entries (its r-th best when the list is sorted; a valid bound for r of
    // its entries in any case, so unsorted input lists are merely pruned less)
    for (int l = tid; l < n2; l += kMergeThreads) {
      uint64_t lo = 0ull;
      if (l < nlists) {
        lo = ~0ull;
        for (int j = 0; j < r; ++j) {
          const uint64_t kj = key_of(l * k_in + j);
          lo = kj < lo ? kj : lo;
        }
        lo = (lo & 0xFFFFFFFF00000000ull) | 1ull;  // score field only; an empty slot among the r gives score 0
      }
      keys[l] = lo;
    }
    bitonic_sort_desc(keys, n2, tid, kMergeThreads, kMergeBar);
    if (tid == 0) s_t0 = (uint32_t)(keys[need - 1] >> 32);  // 0 if fewer than `need` lists have r entries: no pruning
    named_bar_sync(kMergeBar, kMergeThreads);
    t0 = s_t0;
  }
  if (tid == 0) s_cnt = 0;
  named_bar_sync(kMergeBar, kMergeThreads);  // everyone has read keys[] / s_t0 before keys[] is refilled
  int n_sort = cap;
  if (t0 != 0) {
    for (int i = tid; i < m; i += kMergeThreads) {
      const uint64_t key = key_of(i);
      if (key != 0ull && (uint32_t)(key >> 32) >= t0) keys[atomicAdd(&s_cnt, 1)] = key;
    }
    named_bar_sync(kMergeBar, kMergeThreads);
    const int kept = s_cnt;
    n_sort = 64;
    while (n_sort < kept) n_sort <<= 1;
    for (int i = kept + tid; i < n_sort; i += kMergeThreads) keys[i] = 0ull;
  } else {
    for (int i = tid; i < cap; i += kMergeThreads) keys[i] = i < m ? key_of(i) : 0ull;
  }
  bitonic_sort_desc(keys, n_sort, tid, kMergeThreads, kMergeBar);
  // position in the final order: rank inside the run of equal scores is by ascending id
  for (int i = tid; i < min(m, n_sort); i += kMergeThreads) {
    const uint64_t key = keys[i];
    if (key == 0ull) continue;
    const uint32_t so = (uint32_t)(key >> 32);
    const int64_t id = id_at((int)key_row(key));
    int lo = i;
    while (lo > 0 && (uint32_t)(keys[lo - 1] >> 32) == so) --lo;
    int rank = 0;
    bool tie = (lo != i) || (i + 1 < n_sort && (uint32_t)(keys[i + 1] >> 32) == so && keys[i + 1] != 0ull);
    if (tie) {
      for (int j = lo; j < n_sort && keys[j] != 0ull && (uint32_t)(keys[j] >> 32) == so; ++j) {
        const int64_t idj = id_at((int)key_row(keys[j]));
        if (idj < id || (idj == id && j < i)) ++rank;
      }
    } else {
      lo = i;
    }
    const int dst = lo + rank;
    if (dst < k_out) {
      out_scores[(size_t)q * k_out + dst] = key_score(key);
      out_ids[(size_t)q * k_out + dst] = id;
    }
  }
  // padding
  named_bar_sync(kMergeBar, kMergeThreads);
  for (int i = tid; i < k_out; i += kMergeThreads) {
    if (i >= n_sort || keys[i] == 0ull) {
      out_scores[(size_t)q * k_out + i] = -CUDART_INF_F;
      out_ids[(size_t)q * k_out + i] = -1;
    }
  }
}

}  // namespace rs
