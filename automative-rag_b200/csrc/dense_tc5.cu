// dense_tc5.cu — batched dense top-k: many queries against the corpus as ONE tcgen05 GEMM with the
// per-query top-k fused into the epilogue, so the [nq, n] score matrix never reaches HBM
// (BASELINE config 3: 1024 queries x 10M x 1024 bf16, top-100).
//
// Same arithmetic as the single-query scan (the search behind
// QdrantStore.similarity_search_with_score, reference vectorstore.py:166-214), different shape: with
// hundreds of queries the corpus read is amortised and the stage is tensor-bound
// (2*nq*n*d flops, arithmetic intensity ~nq flop/byte).
//
// Mapping:
//   * scores = Q [nq, d] . C^T [d, n];  M (TMEM lanes) = queries, N (TMEM columns) = corpus rows, so
//     tcgen05.ld 32x32b hands one epilogue THREAD one query row with 32 consecutive corpus rows: the
//     running top-k of a query is thread-private state (threshold in a register), and the common
//     case per 32 scores is 16 FMNMX3 + one compare against the threshold.
//   * CTA tile = 256 queries (two 128-row UMMA tiles, one 128x256 fp32 TMEM accumulator each =
//     512 columns) x 256 corpus rows; K = d streamed in 64-element blocks through a 3-stage TMA ring
//     (A 2 x 16 KB + B 32 KB per stage).  grid = (corpus ranges) x (query groups), one wave; the
//     CTAs of one range run side by side so the corpus is read from HBM once and from L2 after that.
//   * candidates above the threshold go to a per-(CTA, query) buffer in global memory (L2
//     resident) with room for a whole tile of appends; at the end of a tile — after the accumulators
//     have been handed back — the WARP sorts the keys of each query whose buffer passed 192 entries in
//     shared memory and keeps the best k, raising that query's threshold.  Expected k*ln(rows/k)
//     appends per query.
//   * each CTA writes its per-query top-k lists; rs_topk_merge (one more small launch) merges the
//     ranges.  Result order and tie rule are those of the scan: (score desc, id asc).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <math_constants.h>
#include <type_traits>
#include <vector>

#include "tc5.cuh"
#include "tc5_host.h"

namespace rs {

bool tc5_encode(const Tc5State* s, CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, std::string* err);
void* tc5_dense_scratch(Tc5State* s, size_t bytes);
int tc5_num_sms(const Tc5State* s);

constexpr int kDtThreads = 384;   // single-CTA tiles: warps 0 TMA, 1 MMA, 2 TMEM, 4-7 / 8-11 epilogue sets
constexpr int kDtThreadsPair = 256;  // CTA pairs: one epilogue set per CTA
constexpr int kDtBN = 256;        // corpus rows per tile (UMMA N)
constexpr int kDtBK = 64;         // K elements per stage (one 128-byte swizzle row)
constexpr int kDtMT = 2;          // 128-query UMMA tiles per CTA (single-CTA tiles) / CTAs per pair
constexpr int kDtStages = 3;
constexpr int kDtStagesSmall = 4;  // batches of <= 128 queries stage ONE query tile: 48 KB stages, four of them (128 KB of
                                   // corpus in flight per SM instead of 96: the pass is HBM-bound, not tensor-bound)
constexpr int kDtStagesPair = 6;  // a pair's stage is half the size: 128 query rows + 128 corpus rows per CTA
constexpr int kDtMaxGroups = 16;  // groups of ranges behind the cross-range bound (one batch of loads per refresh)
constexpr int kDtMaxGm = 8;       // best scores tracked per (range, query) for the cross-range bound
constexpr int kDtCap = 512;       // candidate buffer entries per (CTA, query): room for a whole tile (256 rows) of
                                  // appends on top of kDtCompactAt, so compaction can wait for the end of the tile
constexpr int kDtCompactAt = 192; // compact a query's buffer once it holds more than this (keeps the sort at 256 keys)
constexpr uint32_t kDtABytes = 128 * 128;      // one 128-row query tile, one K block
constexpr uint32_t kDtBBytes = kDtBN * 128;    // one 256-row corpus tile, one K block
constexpr uint32_t kDtStageBytes = kDtMT * kDtABytes + kDtBBytes;

struct DenseTcParams {
  const void* queries;      // [nq, d] (for the query norms)
  const float* inv_norm;    // [n] or null
  const uint32_t* mask;     // bit mask or null; query q uses the words at mask + q * mask_stride (0 = one shared mask)
  int64_t mask_stride;
  uint64_t* cand;           // [grid, 256, kDtCap] candidate keys
  float* list_scores;       // [ranges, nq, k_list]
  int64_t* list_ids;        // [ranges, nq, k_list]
  float* qscale_out;        // [nq] the factor the list scores carry on top of the threshold words (cosine: 1/|q|)
  int32_t k_list;           // slots per (range, query) list, >= k
  int32_t a_rows;           // query rows per TMA box: 128, or nq rounded up to 8 for a batch below 128 queries
  int64_t n, id_base;
  int32_t nq, d, k, metric;
  int32_t num_ranges, tiles_total;
  uint32_t* gthr;           // [cross_groups, nq] orderable score: max over the group's ranges of the range's gm-th
                            // best row so far (0 = none yet)
  int32_t gm;               // rows a range vouches for, 1..kDtMaxGm, or 0 = no cross-CTA threshold
  int32_t bootstrap;        // 1 = bisect a first threshold out of each range's first tile (0: RS_DENSE_NO_BOOTSTRAP)
  int32_t cross_groups;     // ceil(k / gm) <= min(num_ranges, kDtMaxGroups): the cross-CTA threshold is the minimum over these
  long long* trace;         // diagnostics (RS_DENSE_TRACE=1): [CTAs][16] cycles each role spent waiting, or null
  int32_t* progress;        // [num_ranges, query groups]: corpus tiles whose loads a (range, group) has issued (0 at launch)
  int32_t drift;            // a group loads at most this many tiles ahead of the slowest group of its range (0 = free)
};

// wait (+ cycles spent waiting when a trace buffer is attached)
__device__ __forceinline__ void dt_timed_wait(uint64_t* bar, uint32_t parity, long long& acc, bool tracing) {
  if (!tracing) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t;
}

// warp-level bitonic sort of n (power of two, 64..512) u64 keys in shared memory, descending
// (one warp, __syncwarp only)
__device__ __forceinline__ void warp_sort_desc(uint64_t* keys, int n, int lane) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int r = 0; r < (n >> 6); ++r) {
        const int i = r * 32 + lane;  // n / 2 pairs
        const int lo = 2 * i - (i & (stride - 1));
        bitonic_ce(keys, lo, stride, size);
      }
    }
  }
  __syncwarp();
}

// PAIR = false: one CTA per (range, 256-query group) with up to two 128-query UMMA tiles.  Two active tiles fill
//   TMEM with their accumulators, so the epilogue and the MMAs of consecutive corpus tiles alternate; a single
//   active tile (batches of <= 128 queries) leaves room for two accumulator generations and overlaps them.
// PAIR = true: a CTA PAIR (2-wide cluster = the two SMs of a TPC) per (range, 256-query group) runs ONE
//   tcgen05.mma.cta_group::2 of M = 256: each CTA stages its 128 queries and HALF of the 256 corpus rows (the tensor
//   cores read the other half from the peer's shared memory), and holds its 128 x 256 accumulator in its own TMEM —
//   256 columns, so TMEM takes TWO accumulators and the epilogue of tile t overlaps the MMAs of tile t+1 at the same
//   operand bytes per flop.  The leader (cluster rank 0) issues the MMAs; TMA loads of both CTAs count on the leader's
//   barrier; tcgen05.commit multicasts "stage free" / "accumulator ready" to both; the peer's epilogue hands
//   accumulators back with a remote mbarrier arrive.
// SMALL (single-CTA tiles only): the batch fits one 128-query tile, so a stage holds one A tile and the ring is a
//   stage deeper.
template <bool BF16, bool PAIR, bool SMALL>
__global__ void __launch_bounds__(kDtThreads, 1)
    dense_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c,
                     const DenseTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  static_assert(!(PAIR && SMALL), "SMALL is a single-CTA variant");
  constexpr int kStages = PAIR ? kDtStagesPair : (SMALL ? kDtStagesSmall : kDtStages);
  constexpr int kMT = (PAIR || SMALL) ? 1 : kDtMT;            // 128-query tiles staged by this CTA
  constexpr uint32_t kBRows = PAIR ? kDtBN / 2 : kDtBN;       // corpus rows staged by this CTA
  constexpr uint32_t kBBytes = kBRows * 128;
  constexpr uint32_t kStageBytes = kMT * kDtABytes + kBBytes;
  constexpr int kAccBufs = 2;                                 // accumulator generations in flight (see dbuf)
  constexpr int kEpiWarps = PAIR ? 4 : 8;

  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int range = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, mgroup = blockIdx.y;
  // Range r takes corpus tiles r, r + R, r + 2R, ...: at any moment all ranges — and the query groups of each, which
  // walk the same tiles — are inside one window of R consecutive tiles, so a tile is fetched from DRAM once and
  // served to the other query groups out of L2 even when they drift a few rounds apart (contiguous ranges let the
  // four groups of a range drift by more than the L2 holds: 2.4x DRAM re-reads, profiles/r01_dense_tc5_v2_ncu.txt).
  const int ntiles = (p.tiles_total - range + p.num_ranges - 1) / p.num_ranges;
  const int q0 = mgroup * (kDtMT * 128) + (PAIR ? (int)rank * 128 : 0);
  // active 128-query tiles: a pair always runs its one M = 256 MMA (rows past nq are TMA zero fill)
  const int n_act = PAIR ? 1 : min(kMT, (p.nq - q0 + 127) / 128);
  const int kblocks = p.d / kDtBK;
  // Two accumulator generations (the epilogue of tile t overlaps the MMAs of tile t+1) whenever a generation needs
  // only 256 TMEM columns: always for pairs, and for single-CTA tiles when just one 128-query tile is active.
  const bool dbuf = PAIR || n_act == 1;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* stages = sm;  // kStages x [A tiles | B]
  uint64_t* sort_scratch = reinterpret_cast<uint64_t*>(stages + kStages * kStageBytes);  // [kEpiWarps][kDtCap]
  uint64_t* bars = sort_scratch + kEpiWarps * kDtCap;
  uint64_t* full = bars;                      // kStages (pair: the leader's count both CTAs' bytes)
  uint64_t* empty = full + kStages;           // kStages
  uint64_t* acc_full = empty + kStages;       // kAccBufs
  uint64_t* acc_empty = acc_full + kAccBufs;  // kAccBufs (pair: the leader's collect both CTAs' epilogue warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + kAccBufs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tracing = p.trace != nullptr;
  const int cta_linear = (int)(blockIdx.y * gridDim.x + blockIdx.x);
  long long k_c0 = 0;
  unsigned long long k_g0 = 0;
  if (tracing) {
    k_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_g0));
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < kAccBufs; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], PAIR ? 2 * 4 : 4 * n_act);  // the epilogue warps that drain one accumulator generation
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc_cta2(tmem_ptr, 512);
      tmem_relinquish_cta2();
    } else {
      tmem_alloc(tmem_ptr, 512);
      tmem_relinquish();
    }
  }
  // A batch of fewer than 128 queries is staged as a box of a_rows rows: the rest of the 128-row UMMA operand is
  // zeroed here once and never written again (the 128-byte swizzle permutes inside a row, not across rows).  Letting
  // TMA zero-fill the missing rows of a 128-row box costs more than the copy: 0.43-0.57 ms for 3..32 queries over
  // 1M rows where 64 queries took 0.38.
  const uint32_t a_tx = p.a_rows < 128 ? (uint32_t)p.a_rows * 128u : (uint32_t)(PAIR ? 1 : n_act) * kDtABytes;
  if (!PAIR && p.a_rows < 128) {
    for (int s = 0; s < kStages; ++s) {
      uint4* a = reinterpret_cast<uint4*>(stages + (size_t)s * kStageBytes);
      for (int i = threadIdx.x; i < (int)(kDtABytes / 16); i += blockDim.x) a[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  tc5_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc5_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // One lane per box of a stage (A tile(s), B tile): a single issuing thread spends ~0.3-0.5 us per TMA
    // instruction (barrier round trip + issue), which is more than a K block's worth of MMAs.
    const int nbox = PAIR ? 2 : n_act + 1;
    if (lane < nbox && ntiles > 0) {
      tma_prefetch_desc(lane == nbox - 1 ? &map_c : &map_q);
      uint64_t pol;
      if (lane == nbox - 1)
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));  // corpus: streamed
      else
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));    // queries: re-read per tile
      int it = 0;
      const long long T0 = clock64();
      long long w_empty = 0;
      const int ngroups = (int)gridDim.y;
      for (int ti = 0; ti < ntiles; ++ti) {
        const int t = range + ti * p.num_ranges;
        // Keep the query groups of a range within p.drift tiles of each other: every group streams the SAME corpus
        // tiles, and a tile is served from L2 to the followers only while the leader is not further ahead than the L2
        // holds (free-running groups drifted apart over the ~2000 tiles of a range: 2.4-2.7x the corpus from DRAM).
        // Purely a pacing hint: the poll is bounded, so a group that cannot see its peers simply goes on.
        if (p.drift > 0 && ngroups > 1 && ti >= p.drift) {
          const volatile int32_t* pr = p.progress + (size_t)range * ngroups;
          for (int spin = 0; spin < 4096; ++spin) {
            int lo = 0x7fffffff;
            for (int g = 0; g < ngroups; ++g) lo = min(lo, (int)pr[g]);
            if (lo + p.drift >= ti) break;
            __nanosleep(200);
          }
        }
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (uint32_t)(it / kStages) & 1u;
          dt_timed_wait(&empty[s], ph ^ 1u, w_empty, tracing);
          uint8_t* st = stages + (size_t)s * kStageBytes;
          if (PAIR) {
            // both CTAs' bytes land on the LEADER's barrier; only the leader posts the expectation
            if (rank == 0 && lane == 0) mbar_arrive_expect_tx(&full[s], 2u * kStageBytes);
            const uint32_t lead = mapa_u32(smem_u32(&full[s]), 0);
            if (lane == 0)
              tma_load_2d_cta2(st, &map_q, kb * kDtBK, q0, lead, pol);
            else
              tma_load_2d_cta2(st + kDtABytes, &map_c, kb * kDtBK, t * kDtBN + (int)rank * (int)kBRows, lead, pol);
          } else {
            if (lane == 0) mbar_arrive_expect_tx(&full[s], a_tx + kBBytes);
            __syncwarp((1u << nbox) - 1u);
            if (lane < nbox - 1)
              tma_load_2d(st + lane * kDtABytes, &map_q, kb * kDtBK, q0 + lane * 128, &full[s], pol);
            else
              tma_load_2d(st + kMT * kDtABytes, &map_c, kb * kDtBK, t * kDtBN, &full[s], pol);
          }
        }
        if (p.drift > 0 && ngroups > 1 && rank == 0 && lane == nbox - 1)
          *(volatile int32_t*)(p.progress + (size_t)range * ngroups + mgroup) = ti + 1;
      }
      if (tracing && lane == 0) {
        p.trace[cta_linear * 16 + 10] = w_empty;
        p.trace[cta_linear * 16 + 11] = clock64() - T0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair: leader only)
    // Whole warp, uniform control flow, the issuing lane picked by elect.sync: descriptors stay in uniform registers
    // (with `if (lane == 0)` every tcgen05.mma cost ~10 instructions of 64-bit adds, R2UR moves and an ELECT loop).
    if (ntiles > 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BF16, PAIR ? 256 : 128, kDtBN);
      const uint64_t desc0 = umma_smem_desc_sw128(smem_u32(stages));
      int it = 0;
      const long long T0 = clock64();
      long long w_full = 0, w_acc = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int b = dbuf ? (t & 1) : 0;
        const int use = dbuf ? (t >> 1) : t;  // uses of accumulator generation b so far
        dt_timed_wait(&acc_empty[b], ((uint32_t)use & 1u) ^ 1u, w_acc, tracing);  // the epilogue has drained it
        tc5_fence_after();
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (uint32_t)(it / kStages) & 1u;
          dt_timed_wait(&full[s], ph, w_full, tracing);
          tc5_fence_after();
          if (elect_one_sync()) {
            const uint64_t da0 = desc0 + (uint64_t)(((uint32_t)s * kStageBytes) >> 4);
            const uint64_t db = da0 + (uint64_t)((kMT * kDtABytes) >> 4);
            if (PAIR) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ss_cta2(tmem_base + (uint32_t)b * kDtBN, da0 + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc,
                                 (kb | kk) != 0 ? 1u : 0u);
              umma_commit_cta2(&empty[s], 0b11);
            } else {
              for (int a = 0; a < n_act; ++a) {
                const uint64_t da = da0 + (uint64_t)((a * kDtABytes) >> 4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16_ss(tmem_base + (uint32_t)(dbuf ? b : a) * kDtBN, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2),
                              idesc, (kb | kk) != 0 ? 1u : 0u);
              }
              umma_commit(&empty[s]);
            }
            if (kb == kblocks - 1) {
              if (PAIR) umma_commit_cta2(&acc_full[b], 0b11); else umma_commit(&acc_full[b]);
            }
          }
          __syncwarp();
        }
      }
      if (tracing && lane == 0) {
        p.trace[cta_linear * 16 + 0] = clock64() - T0;
        p.trace[cta_linear * 16 + 1] = w_full;
        p.trace[cta_linear * 16 + 2] = w_acc;
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ------------------------------------------------------------------ epilogue: fused per-query top-k
    const int set = (warp - 4) >> 2, quarter = warp & 3;
    if (set < n_act && ntiles > 0) {
      const int lq = set * 128 + quarter * 32 + lane;  // query within this CTA's tiles
      const int query = q0 + lq;
      const bool valid = query < p.nq;
      uint64_t* my_sort = sort_scratch + (size_t)(warp - 4) * kDtCap;
      // candidate buffer of this (range, query): the CTA index is unique per (range, group, rank)
      uint64_t* my_cand = p.cand + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * (kMT * 128) + lq) * kDtCap;
      // query scale for cosine (the corpus side is inv_norm[] or unit rows)
      float q_scale = 1.f;
      if (p.metric == 1 && valid) {
        const uint4* qv = reinterpret_cast<const uint4*>(p.queries) + (size_t)query * (p.d >> 3);
        float ss = 0.f;
        for (int i = 0; i < (p.d >> 3); ++i) {
          const uint4 v = __ldg(qv + i);
          ss = BF16 ? dot8<__nv_bfloat16>(v, v, ss) : dot8<__half>(v, v, ss);
        }
        q_scale = ss > 0.f ? rsqrtf(ss) : 0.f;
        if (ss > 0.f) q_scale = q_scale * (1.5f - 0.5f * ss * q_scale * q_scale);
      }
      // A row enters the candidate buffer when its score beats `thr` = max of two lower bounds on the final
      // k-th best score of the query:
      //   thr_local  the k-th best of THIS range so far (strict >: a later row of equal score has the higher id);
      //   thr_cross  every range publishes its gm-th best score, gm = ceil(k / ranges); gm rows in each of the
      //              ranges score at least the minimum T of those values, i.e. >= k rows overall, so nothing below
      //              T can be in the answer (rows equal to T stay in: thr_cross is the float just below T).
      // A range sees only 1/ranges of the corpus, so thr_local alone lets ~k/rows_seen of the rows through;
      // T behaves like the k-th best of everything all ranges have seen and cuts that by an order of magnitude.
      float thr_local = -CUDART_INF_F, thr_cross = -CUDART_INF_F, thr = -CUDART_INF_F;
      float tm[kDtMaxGm];  // best scores seen, descending (registers: every index below is static)
#pragma unroll
      for (int j = 0; j < kDtMaxGm; ++j) tm[j] = -CUDART_INF_F;
      const int gm = p.gm;
      auto gm_th = [&]() {
        float r = tm[0];
#pragma unroll
        for (int j = 1; j < kDtMaxGm; ++j) r = (gm == j + 1) ? tm[j] : r;
        return r;
      };
      // this range's group: the first (ranges % G) groups hold one range more than the others
      const int grp_base = p.num_ranges / p.cross_groups, grp_rem = p.num_ranges % p.cross_groups;
      const int my_group = range < grp_rem * (grp_base + 1) ? range / (grp_base + 1)
                                                            : grp_rem + (range - grp_rem * (grp_base + 1)) / grp_base;
      uint32_t* my_gthr = p.gthr + (size_t)my_group * p.nq + (valid ? query : 0);
      int next_refresh = 0;  // tile after which thr_cross is read again
      int cnt = 0;
      const int k = p.k;
      const bool bootstrap = p.bootstrap != 0;

      // the warp sorts lane L's candidate buffer and keeps the best k
      auto compact_lane = [&](int L) {
        const uint64_t* src = (const uint64_t*)__shfl_sync(0xFFFFFFFFu, (unsigned long long)my_cand, L);
        const int c = __shfl_sync(0xFFFFFFFFu, cnt, L);
        int n2 = 64;
        while (n2 < c) n2 <<= 1;  // sort only the smallest power of two that holds the valid keys
        for (int i = lane; i < n2; i += 32) my_sort[i] = i < c ? __ldcg(src + i) : 0ull;
        warp_sort_desc(my_sort, n2, lane);
        const int kept = min(c, k);
        uint64_t* dst = const_cast<uint64_t*>(src);
        for (int i = lane; i < kept; i += 32) __stcg(dst + i, my_sort[i]);
        const uint64_t kth = my_sort[kept - 1 < 0 ? 0 : kept - 1];
        __syncwarp();
        if (lane == L) {
          cnt = kept;
          if (kept == k) {
            thr_local = key_score(kth);
            thr = fmaxf(thr_local, thr_cross);
          }
        }
      };

      const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t lead_acc_empty = PAIR ? mapa_u32(smem_u32(acc_empty), 0) : 0u;
      uint32_t va[32], vb[32];
      const long long T0 = clock64();
      long long w_accf = 0, t_compact = 0, t_boot = 0, t_refresh = 0, t_slow = 0, n_slow = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int b = dbuf ? (t & 1) : 0;
        const int use = dbuf ? (t >> 1) : t;
        dt_timed_wait(&acc_full[b], (uint32_t)use & 1u, w_accf, tracing);
        tc5_fence_after();
        const uint32_t taddr = tlane + (uint32_t)(dbuf ? b : set) * kDtBN;
        const int64_t row_base = (int64_t)(range + t * p.num_ranges) * kDtBN;
        auto consume = [&](uint32_t (&v)[32], int ch) {
          const int64_t r0 = row_base + ch * 32;
          if (r0 >= p.n) return;  // uniform
          if (p.inv_norm) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float w = (r0 + c < p.n) ? __ldg(p.inv_norm + r0 + c) : 0.f;
              v[c] = __float_as_uint(__uint_as_float(v[c]) * w);
            }
          }
          float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            m0 = fmaxf(fmaxf(m0, __uint_as_float(v[c + 0])), __uint_as_float(v[c + 1]));
            m1 = fmaxf(fmaxf(m1, __uint_as_float(v[c + 2])), __uint_as_float(v[c + 3]));
            m2 = fmaxf(fmaxf(m2, __uint_as_float(v[c + 4])), __uint_as_float(v[c + 5]));
            m3 = fmaxf(fmaxf(m3, __uint_as_float(v[c + 6])), __uint_as_float(v[c + 7]));
          }
          const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          if (__any_sync(0xFFFFFFFFu, valid && m > thr)) {
            const long long ts0 = tracing ? clock64() : 0;
            uint32_t bits = (p.n - r0 >= 32) ? 0xFFFFFFFFu : ((1u << (int)(p.n - r0)) - 1u);
            if (p.mask) bits &= __ldg(p.mask + (size_t)(valid ? query : 0) * p.mask_stride + (r0 >> 5));
            if (valid && m > thr) {
              float cbest = -CUDART_INF_F;
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const float s = __uint_as_float(v[c]);
                if (s > thr && ((bits >> c) & 1u)) {
                  __stcg(my_cand + cnt, make_key(s, (uint32_t)(r0 + c)));
                  ++cnt;
                  cbest = fmaxf(cbest, s);
                }
              }
              // Keep the best scores sorted and publish the gm-th when it moves.  One insertion per chunk (its
              // best appended row) keeps this out of the unrolled loop — 32 inlined copies pushed the kernel out of
              // the instruction cache; a second top-gm row inside the same 32 rows only makes the bound looser,
              // never wrong.
              if (gm > 0 && cbest > tm[kDtMaxGm - 1]) {
                const float before = gm_th();
                tm[kDtMaxGm - 1] = cbest;
#pragma unroll
                for (int j = kDtMaxGm - 1; j > 0; --j) {
                  if (tm[j] > tm[j - 1]) {
                    const float x = tm[j - 1];
                    tm[j - 1] = tm[j];
                    tm[j] = x;
                  }
                }
                const float after = gm_th();
                if (after > before) atomicMax(my_gthr, f32_orderable(after));
              }
            }
            if (tracing) {
              t_slow += clock64() - ts0;
              ++n_slow;
            }
          }
        };
        const long long tb0 = tracing ? clock64() : 0;
        if (t == 0 && bootstrap) {
          // ---- first tile of the range: find a threshold before anything is appended.  Without one every row of
          // the tile enters the buffer and each of the warp's 32 queries pays a 256-key sort for it (~19 k cycles
          // each: half the kernel for a 64-query batch over 1M rows).  Bisection between the tile's lowest and
          // highest score, each step one more read of the accumulators (they stay in TMEM), all 32 queries in
          // parallel: the largest midpoint that still leaves >= k admissible rows above it is a valid thr_local.
          float hi = -CUDART_INF_F, lo = CUDART_INF_F, best = -CUDART_INF_F;
          auto scan = [&](auto first_tag, float mid, float& mx, float& mn) {
            constexpr bool FIRST = decltype(first_tag)::value;
            int above = 0;
#pragma unroll 1
            for (int ch = 0; ch < kDtBN / 32; ++ch) {
              const int64_t r0 = row_base + ch * 32;
              if (r0 >= p.n) break;  // uniform
              tmem_ld_32x32(taddr + ch * 32, va);
              uint32_t bits = (p.n - r0 >= 32) ? 0xFFFFFFFFu : ((1u << (int)(p.n - r0)) - 1u);
              if (p.mask) bits &= __ldg(p.mask + (size_t)(valid ? query : 0) * p.mask_stride + (r0 >> 5));
              tmem_ld_wait(va);
              if (!FIRST && !p.inv_norm && __all_sync(0xFFFFFFFFu, bits == 0xFFFFFFFFu)) {
#pragma unroll
                for (int c = 0; c < 32; ++c) above += __uint_as_float(va[c]) > mid ? 1 : 0;
                continue;
              }
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                float sc = __uint_as_float(va[c]);
                if (p.inv_norm) sc *= (r0 + c < p.n) ? __ldg(p.inv_norm + r0 + c) : 0.f;
                if ((bits >> c) & 1u) {
                  above += sc > mid ? 1 : 0;
                  if (FIRST) {
                    mx = fmaxf(mx, sc);
                    mn = fminf(mn, sc);
                  }
                }
              }
            }
            return above;
          };
          float dummy_hi = 0.f, dummy_lo = 0.f;
          scan(std::true_type{}, CUDART_INF_F, hi, lo);
#pragma unroll 1
          for (int it = 0; it < 6; ++it) {
            const float mid = 0.5f * lo + 0.5f * hi;
            const int above = scan(std::false_type{}, mid, dummy_hi, dummy_lo);
            if (above >= k) {
              best = mid;
              lo = mid;
            } else {
              hi = mid;
            }
          }
          if (valid && best > thr_local) {
            thr_local = best;
            thr = fmaxf(thr_local, thr_cross);
          }
          if (tracing) t_boot += clock64() - tb0;
        }
        tmem_ld_32x32(taddr, va);
        tmem_ld_wait(va);
#pragma unroll 1
        for (int ch = 0; ch < kDtBN / 32; ch += 2) {
          tmem_ld_32x32(taddr + (ch + 1) * 32, vb);
          consume(va, ch);
          tmem_ld_wait(vb);
          if (ch + 2 < kDtBN / 32) {
            tmem_ld_32x32(taddr + (ch + 2) * 32, va);
          } else {
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster(lead_acc_empty + (uint32_t)b * 8u); else mbar_arrive(&acc_empty[b]);
            }
          }
          consume(vb, ch + 1);
          if (ch + 2 < kDtBN / 32) tmem_ld_wait(va);
        }
        // The accumulators went back to the MMA warp above, so compaction runs while the tensor core works on
        // the next tile.  Done inside the tile it sat on the critical path: the MMA restarts only when ALL
        // the warps have drained, and some warp compacts in almost every tile.
        uint32_t need = __ballot_sync(0xFFFFFFFFu, cnt > kDtCompactAt);
        const long long tc0 = tracing ? clock64() : 0;
        while (need) {
          const int L = __ffs(need) - 1;
          need &= need - 1;
          compact_lane(L);
        }
        if (tracing) t_compact += clock64() - tc0;
        // refresh the cross-range bound (also off the critical path; stale values are only lower, never wrong).
        // The ranges form G = ceil(k / gm) <= 16 groups; a group's word is the atomic max of what its ranges
        // published, i.e. it stands for gm rows of ONE of its ranges, so the minimum T over the groups has
        // >= G * gm >= k distinct rows at or above it.  (Round 2 started with one word per range and the minimum
        // over all of them: with k = 10 over 148 ranges that is the worst of 148 range maxima instead of the 10th
        // best of group maxima — an order of magnitude more rows passed — and reading 148 words under a saturated
        // memory system cost ~25 k cycles per refresh, more than the rest of a small batch's epilogue.)
        if (gm > 0 && t == next_refresh) {
          // The bound moves like 1/rows_seen: refresh after every tile at first, then at geometrically growing
          // distances.
          // (more than 16 groups — the long lists of k > 128 — cost one round trip per 16 words: refresh half as often)
          next_refresh = t + 1 + t / 8 + (p.cross_groups > kDtMaxGroups ? 1 + t / 8 : 0);
          const long long tr0 = tracing ? clock64() : 0;
          uint32_t lo = 0xFFFFFFFFu;
          const uint32_t* g = p.gthr + (valid ? query : 0);
          for (int j0 = 0; j0 < p.cross_groups; j0 += kDtMaxGroups) {
#pragma unroll
            for (int j = 0; j < kDtMaxGroups; ++j) {  // one L2 round trip: the loads are independent
              const uint32_t v = j0 + j < p.cross_groups ? __ldcg(g + (size_t)(j0 + j) * p.nq) : 0xFFFFFFFFu;
              lo = min(lo, v);
            }
          }
          if (valid && lo > 1u) {  // every group has a published value: the float just below T
            thr_cross = orderable_f32(lo - 1u);
            thr = fmaxf(thr_local, thr_cross);
          }
          if (tracing) t_refresh += clock64() - tr0;
        }
      }
      if (tracing && quarter == 0 && lane == 0 && set == 0) {
        p.trace[cta_linear * 16 + 4] = clock64() - T0;
        p.trace[cta_linear * 16 + 5] = w_accf;
        p.trace[cta_linear * 16 + 6] = t_compact;
        p.trace[cta_linear * 16 + 7] = t_boot;
        p.trace[cta_linear * 16 + 8] = t_refresh;
        p.trace[cta_linear * 16 + 9] = t_slow;
        p.trace[cta_linear * 16 + 12] = n_slow;
      }
      // ---- final: the range's list of every query, kl >= k slots, in NO particular order (the merge does not need
      // one), all 32 queries of the warp at once.  A buffer that fits the list is copied as it is.  A longer one is
      // cut by a score threshold with between k and kl keys above it, found by bisection over the buffer staged in
      // shared memory (the pipeline stages are dead by now: the last acc_full was committed after every MMA that read
      // them).  Only a buffer whose k-th score is tied beyond kl falls back to the warp's sort.  (Sorting every buffer,
      // one query after the other with an L2 round trip each, was a quarter of a 64-query batch over 1M rows at
      // k = 10 and more than half of it at k = 100.)
      const int kl = p.k_list;
      const bool staged = PAIR || n_act == 1;  // 4 epilogue warps x 48 KB
      uint64_t* col = reinterpret_cast<uint64_t*>(stages) + (size_t)(warp - 4) * (kDtCompactAt * 32) + lane;
      uint32_t cut = 0u;        // keys with score bits above this go to the list
      bool from_global = !staged;
      if (staged) {
#pragma unroll 8
        for (int i = 0; i < cnt; ++i) col[i * 32] = __ldcg(my_cand + i);
        bool searching = valid && cnt > kl;
        uint32_t lo = 0xFFFFFFFFu, hi = 0u;
        if (searching) {
          for (int i = 0; i < cnt; ++i) {
            const uint32_t sb = (uint32_t)(col[i * 32] >> 32);
            lo = min(lo, sb);
            hi = max(hi, sb);
          }
          lo = lo > 0u ? lo - 1u : 0u;  // every key is above lo; none is above hi
        }
        bool failed = false;
        while (__any_sync(0xFFFFFFFFu, searching)) {
          if (searching) {
            if (hi - lo <= 1u) {
              failed = true;
              searching = false;
            } else {
              const uint32_t mid = lo + (hi - lo) / 2u;
              int above = 0;
              for (int i = 0; i < cnt; ++i) above += (uint32_t)(col[i * 32] >> 32) > mid ? 1 : 0;
              if (above > kl) {
                lo = mid;
              } else if (above < k) {
                hi = mid;
              } else {
                cut = mid;
                searching = false;
              }
            }
          }
        }
        uint32_t need = __ballot_sync(0xFFFFFFFFu, failed);
        from_global = failed;
        while (need) {
          const int L = __ffs(need) - 1;
          need &= need - 1;
          compact_lane(L);
        }
      } else {
        uint32_t need = __ballot_sync(0xFFFFFFFFu, valid && cnt > kl);
        while (need) {
          const int L = __ffs(need) - 1;
          need &= need - 1;
          compact_lane(L);
        }
      }
      if (valid && range == 0) p.qscale_out[query] = q_scale;
      if (valid) {
        float* ls = p.list_scores + ((size_t)range * p.nq + query) * kl;
        int64_t* li = p.list_ids + ((size_t)range * p.nq + query) * kl;
        int w = 0;
        for (int i = 0; i < cnt; ++i) {
          const uint64_t key = from_global ? __ldcg(my_cand + i) : col[i * 32];
          if (from_global || (uint32_t)(key >> 32) > cut) {
            ls[w] = key_score(key) * q_scale;
            li[w] = p.id_base + (int64_t)key_row(key);
            ++w;
          }
        }
        for (; w < kl; ++w) {
          ls[w] = -CUDART_INF_F;
          li[w] = -1;
        }
      }
    }
  }

  tc5_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();  // pair: the peer may still read this CTA's tiles / barriers
  if (tracing && threadIdx.x == 0) {
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    p.trace[cta_linear * 16 + 13] = clock64() - k_c0;
    p.trace[cta_linear * 16 + 14] = (long long)(g1 - k_g0);
  }
  if (warp == 2) {
    tc5_fence_after();
    if (PAIR) tmem_dealloc_cta2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ lists longer than a range may keep (k > 128)
// For 128 < k <= 1024 the kernel keeps only the k_run <= 128 best rows of every corpus range (ranges are interleaved
// 256-row tiles, so each holds ~k / ranges of a query's top k) and the merge takes the k best of ranges * k_list
// candidates.  That is exact unless some range held MORE than it could keep: a range whose list is full (>= k_run
// valid entries) and whose weakest kept score is not below the merged k-th score may have dropped a row of the answer.
// This kernel flags such queries; the caller re-runs them through the single-query scan (never seen on random data:
// the expected share of a range is k / ranges against k_run slots; thousands of exactly tied scores do trigger it).
__global__ void __launch_bounds__(256) dense_overflow_check_kernel(const float* __restrict__ list_scores,
                                                                   const int64_t* __restrict__ list_ids, int ranges,
                                                                   int nq, int k_list, int k_run,
                                                                   const float* __restrict__ out_scores,
                                                                   const int64_t* __restrict__ out_ids, int k,
                                                                   int32_t* __restrict__ flags) {
  const int q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // merged k-th score; fewer than k results = nothing may have been dropped anywhere
  const bool have_k = out_ids[(size_t)q * k + (k - 1)] >= 0;
  const float kth = have_k ? out_scores[(size_t)q * k + (k - 1)] : -CUDART_INF_F;
  if (threadIdx.x == 0) flags[q] = 0;
  __syncthreads();
  for (int r = warp; r < ranges; r += 8) {
    const size_t base = ((size_t)r * nq + q) * k_list;
    int cnt = 0;
    float lo = CUDART_INF_F;
    for (int j = lane; j < k_list; j += 32) {
      if (list_ids[base + j] >= 0) {
        ++cnt;
        lo = fminf(lo, list_scores[base + j]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
      lo = fminf(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
    }
    if (lane == 0 && cnt >= k_run && lo >= kth) flags[q] = 1;
  }
}

// Sizing shared by tc5_dense_supported and tc5_dense_topk.
struct DtPlan {
  bool pair, two_level;
  int mgroups, tiles_total, ranges, k_run, k_list;
};
static bool dt_plan(int num_sms, int64_t n, int nq, int k, DtPlan* pl) {
  static const int no_pair = getenv("RS_DENSE_NO_PAIR") ? 1 : 0;  // A/B switch for scripts/batch_bench.py
  // pairs cover 256 queries; up to 128 queries fit one CTA's single M tile, where a pair's second CTA would idle
  pl->pair = !no_pair && num_sms >= 2 && nq > 128;
  pl->mgroups = (nq + kDtMT * 128 - 1) / (kDtMT * 128);  // 256 queries per CTA / per CTA pair
  pl->tiles_total = (int)((n + kDtBN - 1) / kDtBN);
  int ranges = (pl->pair ? num_sms / 2 : num_sms) / pl->mgroups;
  if (ranges < 1) ranges = 1;
  if (ranges > pl->tiles_total) ranges = pl->tiles_total;
  pl->two_level = k > 128;
  if (!pl->two_level) {
    // slots per (range, query) list: k plus slack, so that a threshold with [k, k_list] keys above it is easy to
    // find; ranges * k_list keys per query must fit rs_topk_merge's shared memory
    int k_list = std::max(32, k + std::max(k / 2, 22));
    if ((long long)ranges * k_list > 16384) k_list = std::max(k, 16384 / ranges);
    while ((long long)ranges * k_list > 16384) --ranges;
    pl->ranges = ranges;
    pl->k_run = k;
    pl->k_list = k_list;
    return true;
  }
  if (k > 1024 || ranges < 2) return false;
  const int kl_max = 16384 / ranges;                      // >= 110 for <= 148 ranges
  const int k_run = std::min(128, kl_max * 4 / 5);
  pl->ranges = ranges;
  pl->k_run = k_run;
  pl->k_list = std::min(kl_max, k_run + std::max(k_run / 2, 22));
  // a range must be able to hold twice its expected share of the answer (the rest is the overflow check's business)
  return k_run >= 32 && 2ll * k <= (long long)ranges * k_run;
}

// ================================================================================ host side
bool tc5_dense_supported(const Tc5State* s, int64_t n, int d, int nq, int k, const uint32_t* mask,
                         int64_t mask_stride_words, bool worthwhile) {
  (void)mask;
  if (!s) return false;
  // One batched pass over the corpus beats a loop of single-query scans from 4 queries on at every corpus size, and
  // from 2 queries once the corpus is large enough for the pass to be HBM-bound (1M x 1024 rows, k = 10: 0.35 ms for
  // 2..32 queries vs 0.59 ms for two scans; 20k rows: 0.053 vs 0.038 ms; profiles/r02_batch_crossover.txt).
  static const int min_nq = getenv("RS_DENSE_TC_MIN_NQ") ? atoi(getenv("RS_DENSE_TC_MIN_NQ")) : 0;
  if (n < kDtBN) return false;
  if (nq < 2) return false;
  // `worthwhile` (the AUTO choice) adds the profitability rule to what the kernel can do
  if (worthwhile && (min_nq > 0 ? nq < min_nq : (nq < 4 && n < 200000))) return false;
  if (d % kDtBK != 0 || d < kDtBK) return false;
  (void)mask_stride_words;                       // a filter per query is a per-thread mask word in the epilogue
  if (n >= (1ll << 31) * (int64_t)1) return false;
  DtPlan pl;
  return dt_plan(tc5_num_sms(s), n, nq, k, &pl);  // k <= 128 directly, k <= 1024 through the two-level lists
}

int tc5_dense_topk(Tc5State* s, const void* corpus, int64_t n, int d, int dtype, const float* inv_norm, int metric,
                   const void* queries, int nq, const uint32_t* mask, int64_t mask_stride_words, int k, int64_t id_base, float* out_scores,
                   int64_t* out_ids, cudaStream_t stream, int* launched, std::string* err, std::vector<int>* redo) {
  *launched = 0;
  if (redo) redo->clear();
  const int num_sms = tc5_num_sms(s);
  DtPlan pl;
  if (!dt_plan(num_sms, n, nq, k, &pl)) {
    *err = "unsupported (k, corpus) for the batched kernel";
    return -2;
  }
  const bool pair = pl.pair;
  const int mgroups = pl.mgroups, tiles_total = pl.tiles_total, ranges = pl.ranges, k_list = pl.k_list;
  const int k_out = k;  // what the caller asked for
  k = pl.k_run;         // what a range keeps (== k_out unless two_level)
  if (pl.two_level && !redo) {
    *err = "k > 128 needs a fallback list";
    return -2;
  }

  const int a_rows = nq < 128 ? (nq + 7) / 8 * 8 : 128;
  CUtensorMap map_q, map_c;
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)nq};
    const uint64_t strides[1] = {(uint64_t)d * 2};
    const uint32_t box[2] = {kDtBK, (uint32_t)a_rows};
    if (!tc5_encode(s, &map_q, dtype, 2, queries, dims, strides, box, err)) return -2;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)n};
    const uint64_t strides[1] = {(uint64_t)d * 2};
    const uint32_t box[2] = {kDtBK, (uint32_t)(pair ? kDtBN / 2 : kDtBN)};  // a pair's CTA stages half of the tile
    if (!tc5_encode(s, &map_c, dtype, 2, corpus, dims, strides, box, err)) return -2;
  }
  const size_t cand_bytes = (size_t)ranges * mgroups * (kDtMT * 128) * kDtCap * sizeof(uint64_t);
  const size_t ls_bytes = ((size_t)ranges * nq * k_list * sizeof(float) + 255) / 256 * 256;
  const size_t li_bytes = ((size_t)ranges * nq * k_list * sizeof(int64_t) + 255) / 256 * 256;
  const size_t gt_bytes = ((size_t)ranges * nq * sizeof(uint32_t) + 255) / 256 * 256;
  const size_t qs_bytes = ((size_t)nq * sizeof(float) + 255) / 256 * 256;
  const size_t pr_bytes = ((size_t)ranges * mgroups * sizeof(int32_t) + 255) / 256 * 256;
  const size_t fl_bytes = ((size_t)nq * sizeof(int32_t) + 255) / 256 * 256;
  uint8_t* ws = static_cast<uint8_t*>(
      tc5_dense_scratch(s, cand_bytes + ls_bytes + li_bytes + gt_bytes + qs_bytes + pr_bytes + fl_bytes));
  if (!ws) {
    *err = "out of device memory for the candidate buffers";
    return -5;
  }
  DenseTcParams kp{};
  kp.queries = queries;
  kp.inv_norm = metric == 1 ? inv_norm : nullptr;
  kp.mask = mask;
  kp.mask_stride = mask_stride_words;
  kp.cand = reinterpret_cast<uint64_t*>(ws);
  kp.list_scores = reinterpret_cast<float*>(ws + cand_bytes);
  kp.list_ids = reinterpret_cast<int64_t*>(ws + cand_bytes + ls_bytes);
  kp.n = n;
  kp.id_base = id_base;
  kp.nq = nq;
  kp.d = d;
  kp.k = k;
  kp.k_list = k_list;
  kp.a_rows = a_rows;
  kp.metric = metric;
  kp.num_ranges = ranges;
  kp.tiles_total = tiles_total;
  kp.gthr = reinterpret_cast<uint32_t*>(ws + cand_bytes + ls_bytes + li_bytes);
  kp.qscale_out = reinterpret_cast<float*>(ws + cand_bytes + ls_bytes + li_bytes + gt_bytes);
  kp.progress = reinterpret_cast<int32_t*>(ws + cand_bytes + ls_bytes + li_bytes + gt_bytes + qs_bytes);
  static const int drift = getenv("RS_DENSE_DRIFT") ? atoi(getenv("RS_DENSE_DRIFT")) : 2;
  kp.drift = mgroups > 1 ? drift : 0;
  if (kp.drift > 0) {
    cudaError_t me = cudaMemsetAsync(kp.progress, 0, pr_bytes, stream);
    if (me != cudaSuccess) {
      *err = cudaGetErrorString(me);
      return -3;
    }
  }
  static const int cross_off = getenv("RS_DENSE_NO_CROSS_THR") ? 1 : 0;  // A/B switch for scripts/batch_bench.py
  // Rows a range vouches for: as few as the number of ranges allows, down to 64 groups.  The first version capped the
  // groups at 16 (one batch of loads per refresh), which made a range vouch for 7 rows at k = 100 — its 7th best is
  // far below the global 100th; with the refresh reading 16 words per round trip, 50 groups of 2 rows take 64 / 128
  // queries x 1M rows at k = 100 from 0.64 / 0.68 to 0.43 / 0.46 ms (profiles/r02_dense_gm_ab.txt).
  constexpr int kGroupCap = 64;
  const int gm = std::max((k + ranges - 1) / ranges, (k + std::min(ranges, kGroupCap) - 1) / std::min(ranges, kGroupCap));
  kp.gm = (ranges > 1 && gm <= kDtMaxGm && !cross_off && !pl.two_level) ? gm : 0;
  static const int gm_env = getenv("RS_DENSE_GM") ? atoi(getenv("RS_DENSE_GM")) : 0;  // experiment: rows a range vouches for
  if (gm_env > 0 && kp.gm > 0 && !pl.two_level && gm_env <= kDtMaxGm && (k + gm_env - 1) / gm_env <= ranges) kp.gm = gm_env;
  // two-level lists: the bound must vouch for the k_out rows the caller wants, not for the k_run a range keeps: every
  // range publishes its 8th best and the ranges form ceil(k_out / 8) groups — possible while that is <= ranges (the
  // 148 ranges of a batch of <= 128 queries: k_out <= 1184).  Without it a range's own k_run-th best is the only
  // threshold and 80 % of the 32-row chunks of a 16-query batch take the append path (trace: epilogue 71 % in appends
  // and compactions, the MMA warp waiting 30 % for accumulators).
  if (pl.two_level && ranges > 1 && !cross_off && (k_out + kDtMaxGm - 1) / kDtMaxGm <= ranges) kp.gm = kDtMaxGm;
  static const int bootstrap_off = getenv("RS_DENSE_NO_BOOTSTRAP") ? 1 : 0;
  kp.bootstrap = bootstrap_off ? 0 : 1;
  kp.cross_groups = kp.gm > 0 ? std::min(ranges, ((pl.two_level ? k_out : k) + kp.gm - 1) / kp.gm) : 1;
  if (kp.gm > 0) {
    cudaError_t me = cudaMemsetAsync(kp.gthr, 0, gt_bytes, stream);
    if (me != cudaSuccess) {
      *err = cudaGetErrorString(me);
      return -3;
    }
  }
  static const bool trace_on = getenv("RS_DENSE_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  const int n_ctas = (pair ? 2 * ranges : ranges) * mgroups;
  if (trace_on) {
    if (!trace_dev) cudaMalloc(&trace_dev, (size_t)1024 * 16 * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, (size_t)1024 * 16 * sizeof(long long), stream);
    kp.trace = n_ctas <= 1024 ? trace_dev : nullptr;
  }
  static const int no_small = getenv("RS_DENSE_NO_SMALL") ? 1 : 0;  // A/B switch
  const bool small = !pair && nq <= 128 && !no_small;
  const size_t stage_bytes = pair ? (size_t)(kDtABytes + kDtBBytes / 2)
                                  : (small ? (size_t)(kDtABytes + kDtBBytes) : (size_t)kDtStageBytes);
  const size_t smem = 1024 + (size_t)(pair ? kDtStagesPair : (small ? kDtStagesSmall : kDtStages)) * stage_bytes +
                      (size_t)(pair ? 4 : 8) * kDtCap * sizeof(uint64_t) + 256;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = pair ? dim3(2 * ranges, mgroups) : dim3(ranges, mgroups);
  cfg.blockDim = dim3(pair ? kDtThreadsPair : kDtThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pair ? 1 : 0;
  cudaError_t e;
#define RS_DT_LAUNCH(BF, PR, SM)                                                                                    \
  {                                                                                                                 \
    e = cudaFuncSetAttribute(dense_tc5_kernel<BF, PR, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, dense_tc5_kernel<BF, PR, SM>, map_q, map_c, kp);             \
  }
  if (dtype == 1) {
    if (pair) RS_DT_LAUNCH(true, true, false) else if (small) RS_DT_LAUNCH(true, false, true) else RS_DT_LAUNCH(true, false, false)
  } else {
    if (pair) RS_DT_LAUNCH(false, true, false) else if (small) RS_DT_LAUNCH(false, false, true) else RS_DT_LAUNCH(false, false, false)
  }
#undef RS_DT_LAUNCH
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return -3;
  }
  *launched = 1;
  if (trace_on && kp.trace) {
    std::vector<long long> t((size_t)n_ctas * 16);
    cudaStreamSynchronize(stream);
    cudaMemcpy(t.data(), trace_dev, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    auto avg = [&](int slot, int rank_sel) {
      double sum = 0;
      int cnt = 0;
      for (int c = 0; c < n_ctas; ++c)
        if (!pair || rank_sel < 0 || (c & 1) == rank_sel) {
          sum += (double)t[(size_t)c * 16 + slot];
          ++cnt;
        }
      return cnt ? sum / cnt : 0.0;
    };
    fprintf(stderr,
            "[dense_tc5 trace] pair %d ranges %d groups %d tiles/range %d | MMA (leader): total %.0f wait full %.0f acc_empty %.0f | "
            "epilogue warp 4 (leader/peer): total %.0f/%.0f wait acc_full %.0f/%.0f compaction %.0f/%.0f first-tile bisection %.0f refresh %.0f append path %.0f in %.0f chunks | producer (leader/peer): "
            "total %.0f/%.0f wait empty %.0f/%.0f  [cycles, mean over CTAs] | whole kernel %.0f cycles in %.0f ns\n",
            (int)pair, ranges, mgroups, tiles_total / ranges, avg(0, 0), avg(1, 0), avg(2, 0), avg(4, 0), avg(4, 1), avg(5, 0),
            avg(5, 1), avg(6, 0), avg(6, 1), avg(7, 0), avg(8, 0), avg(9, 0), avg(12, 0), avg(11, 0), avg(11, 1), avg(10, 0), avg(10, 1), avg(13, -1), avg(14, -1));
  }
  // the lists are in no particular order; the kernel's final cross-range words bound the answer for the merge
  e = launch_topk_merge(kp.list_scores, kp.list_ids, ranges, nq, k_list, k_out, 0, 0, out_scores, out_ids, stream,
                        kp.gm > 0 ? kp.gthr : nullptr, kp.cross_groups, kp.qscale_out);
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return -3;
  }
  *launched = 2;
  if (pl.two_level) {
    int32_t* flags = reinterpret_cast<int32_t*>(ws + cand_bytes + ls_bytes + li_bytes + gt_bytes + qs_bytes + pr_bytes);
    dense_overflow_check_kernel<<<nq, 256, 0, stream>>>(kp.list_scores, kp.list_ids, ranges, nq, k_list, k, out_scores,
                                                        out_ids, k_out, flags);
    e = cudaGetLastError();
    std::vector<int32_t> hf((size_t)nq);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hf.data(), flags, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // the one host decision of this path
    if (e != cudaSuccess) {
      *err = cudaGetErrorString(e);
      return -3;
    }
    *launched = 3;
    for (int q = 0; q < nq; ++q)
      if (hf[(size_t)q]) redo->push_back(q);
  }
  return 0;
}

}  // namespace rs
