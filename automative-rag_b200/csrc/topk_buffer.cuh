// topk_buffer.cuh — CTA-level running top-k over a stream of u64 result keys.
//
// A shared-memory buffer of C (a power of two) keys plus a threshold (the current k-th best key).
// Producers append only keys above the threshold; when the buffer passes its high-water mark
// H = C - slack the CTA sorts it (bitonic, descending), keeps the best k and raises the threshold.
// The caller guarantees that no more than `slack` keys are appended between two `maybe_compact`
// calls, so appends never overflow; C >= 1.5 k + slack leaves at least k/2 appends between two
// compactions.  (The buffer shares the CTA's shared memory with the scan's tile ring: a first
// version with C >= 4k cost k = 1000 half of the ring and 40 % of the scan's bandwidth.)
//
// This replaces a per-thread heap: after the first few hundred rows almost nothing beats the
// threshold (expected appends ~ k * ln(rows / k)), so the common path is one compare per row.
#pragma once
#include "common.cuh"

namespace rs {

struct TopKBuffer {
  uint64_t* keys;     // [C] shared
  uint64_t* thr;      // shared: current threshold key (0 = none yet)
  int* cnt;           // shared: number of valid keys
  int C, hw, k;       // capacity, high-water mark (C - slack), list length
  int tid, nthreads;  // participating threads (named barrier `bar_id`)
  int bar_id;
  uint64_t floor_key = 0ull;  // a key known to be <= the final k-th best (raise_floor); the threshold never drops below it

  __device__ __forceinline__ int high_water() const { return hw; }
  __device__ __forceinline__ int slack() const { return C - hw; }

  static __host__ __device__ __forceinline__ int capacity_for(int k, int slack) {
    int c = 512;
    while (c < k + k / 2 + slack) c <<= 1;
    return c;
  }

  // all participating threads
  __device__ __forceinline__ void init() {
    for (int i = tid; i < C; i += nthreads) keys[i] = 0ull;
    if (tid == 0) {
      *thr = 0ull;
      *cnt = 0;
    }
    named_bar_sync(bar_id, nthreads);
  }

  __device__ __forceinline__ uint64_t threshold() const { return *reinterpret_cast<volatile uint64_t*>(thr); }

  // Warp-cooperative append: every lane of a fully converged warp calls this; lanes with
  // `want` set contribute `key`.
  __device__ __forceinline__ void warp_append(bool want, uint64_t key) {
    unsigned m = __ballot_sync(0xFFFFFFFFu, want);
    if (m == 0) return;
    int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(cnt, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (want) {
      int pos = base + __popc(m & ((1u << lane) - 1u));
      if (pos < C) keys[pos] = key;  // guarded; the caller's contract makes this always true
    }
  }

  // all participating threads; sorts, truncates to k, raises the threshold.  Only the smallest
  // power-of-two prefix that holds the valid keys is sorted (the final compaction of a scan usually
  // finds a few dozen keys, not C), and [n, max(n2, k)) is cleared so keys[0..k) never holds stale data.
  __device__ __forceinline__ void compact() {
    named_bar_sync(bar_id, nthreads);
    const int n = min(*reinterpret_cast<volatile int*>(cnt), C);
    int n2 = 2;
    while (n2 < n) n2 <<= 1;
    const int clear_to = max(n2, min(k, C));
    for (int i = n + tid; i < clear_to; i += nthreads) keys[i] = 0ull;
    bitonic_sort_desc(keys, n2, tid, nthreads, bar_id);
    if (tid == 0) {
      int kept = min(n, k);
      *cnt = kept;
      const uint64_t kth = (kept == k) ? keys[k - 1] : 0ull;
      *thr = kth > floor_key ? kth : floor_key;
    }
    named_bar_sync(bar_id, nthreads);
  }

  // all participating threads, same `bound`: the caller knows that at least k keys of the whole stream are > bound
  // (so nothing <= bound can be in the answer).  Keys already in the buffer stay; they are sorted out by compact().
  __device__ __forceinline__ void raise_floor(uint64_t bound) {
    floor_key = bound;
    if (tid == 0 && bound > *thr) *thr = bound;
    named_bar_sync(bar_id, nthreads);
  }

  // Barrier + uniform decision: compacts iff any thread saw the buffer above high water.
  __device__ __forceinline__ void maybe_compact() {
    bool over = *reinterpret_cast<volatile int*>(cnt) > high_water();
    if (named_bar_or(bar_id, nthreads, over)) compact();
  }
};

}  // namespace rs
