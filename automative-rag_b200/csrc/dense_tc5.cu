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

#include <cstdio>
#include <cstdlib>
#include <math_constants.h>
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
constexpr int kDtStagesPair = 6;  // a pair's stage is half the size: 128 query rows + 128 corpus rows per CTA
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
  float* list_scores;       // [ranges, nq, k]
  int64_t* list_ids;        // [ranges, nq, k]
  int64_t n, id_base;
  int32_t nq, d, k, metric;
  int32_t num_ranges, tiles_total;
  uint32_t* gthr;           // [num_ranges, nq] orderable score of each range's gm-th best row so far (0 = none yet)
  int32_t gm;               // ceil(k / num_ranges) in 1..kDtMaxGm, or 0 = no cross-CTA threshold
  long long* trace;         // diagnostics (RS_DENSE_TRACE=1): [CTAs][16] cycles each role spent waiting, or null
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
template <bool BF16, bool PAIR>
__global__ void __launch_bounds__(kDtThreads, 1)
    dense_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c,
                     const DenseTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int kStages = PAIR ? kDtStagesPair : kDtStages;
  constexpr int kMT = PAIR ? 1 : kDtMT;                       // 128-query tiles staged by this CTA
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
  const int n_act = PAIR ? 1 : min(kDtMT, (p.nq - q0 + 127) / 128);
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
      for (int ti = 0; ti < ntiles; ++ti) {
        const int t = range + ti * p.num_ranges;
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
            if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)n_act * kDtABytes + kBBytes);
            __syncwarp((1u << nbox) - 1u);
            if (lane < nbox - 1)
              tma_load_2d(st + lane * kDtABytes, &map_q, kb * kDtBK, q0 + lane * 128, &full[s], pol);
            else
              tma_load_2d(st + kMT * kDtABytes, &map_c, kb * kDtBK, t * kDtBN, &full[s], pol);
          }
        }
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
      uint32_t* my_gthr = p.gthr + (size_t)range * p.nq + (valid ? query : 0);
      const int cross_every = max(1, p.num_ranges / 32);  // tiles between refreshes of thr_cross
      int cnt = 0;
      const int k = p.k;

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
      long long w_accf = 0, t_compact = 0;
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
                if (after > before) __stcg(my_gthr, f32_orderable(after));
              }
            }
          }
        };
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
        // refresh the cross-range bound (also off the critical path; stale values are only lower, never wrong)
        if (gm > 0 && valid && (t % cross_every) == cross_every - 1) {
          uint32_t lo = 0xFFFFFFFFu;
          const uint32_t* g = p.gthr + query;
          for (int c = 0; c < p.num_ranges; ++c) lo = min(lo, __ldcg(g + (size_t)c * p.nq));
          if (lo > 1u) {  // every range has published: the float just below T
            thr_cross = orderable_f32(lo - 1u);
            thr = fmaxf(thr_local, thr_cross);
          }
        }
      }
      if (tracing && quarter == 0 && lane == 0 && set == 0) {
        p.trace[cta_linear * 16 + 4] = clock64() - T0;
        p.trace[cta_linear * 16 + 5] = w_accf;
        p.trace[cta_linear * 16 + 6] = t_compact;
      }
      // ---- final: every query's buffer sorted, best k written as (score, id) lists of this range
      for (int L = 0; L < 32; ++L) {
        compact_lane(L);
        const int qL = q0 + set * 128 + quarter * 32 + L;
        if (qL < p.nq) {  // uniform
          const int c = __shfl_sync(0xFFFFFFFFu, cnt, L);
          const float sc = __shfl_sync(0xFFFFFFFFu, q_scale, L);
          float* ls = p.list_scores + ((size_t)range * p.nq + qL) * k;
          int64_t* li = p.list_ids + ((size_t)range * p.nq + qL) * k;
          for (int i = lane; i < k; i += 32) {
            if (i < c) {
              const uint64_t key = my_sort[i];
              ls[i] = key_score(key) * sc;
              li[i] = p.id_base + (int64_t)key_row(key);
            } else {
              ls[i] = -CUDART_INF_F;
              li[i] = -1;
            }
          }
        }
        __syncwarp();
      }
    }
  }

  tc5_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();  // pair: the peer may still read this CTA's tiles / barriers
  if (warp == 2) {
    tc5_fence_after();
    if (PAIR) tmem_dealloc_cta2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================ host side
bool tc5_dense_supported(const Tc5State* s, int64_t n, int d, int nq, int k, const uint32_t* mask,
                         int64_t mask_stride_words) {
  (void)mask;
  if (!s) return false;
  // One batched pass over the corpus beats a loop of single-query scans from 4 queries on (1M x 1024 rows, k = 10:
  // 0.60 vs 1.15 ms at 4 queries, 1.1 vs 9.2 ms at 32; profiles/r01_batch_crossover.txt).
  static const int min_nq = getenv("RS_DENSE_TC_MIN_NQ") ? atoi(getenv("RS_DENSE_TC_MIN_NQ")) : 4;
  if (nq < min_nq || n < kDtBN) return false;
  if (d % kDtBK != 0 || d < kDtBK) return false;
  if (k > 128) return false;
  (void)mask_stride_words;                       // a filter per query is a per-thread mask word in the epilogue
  if (n >= (1ll << 31) * (int64_t)1) return false;
  return true;
}

cudaError_t launch_topk_merge(const float* scores, const int64_t* ids, int nlists, int nq, int k_in, int k_out,
                              int64_t score_list_stride, int64_t id_list_stride, float* out_scores, int64_t* out_ids,
                              cudaStream_t stream);

int tc5_dense_topk(Tc5State* s, const void* corpus, int64_t n, int d, int dtype, const float* inv_norm, int metric,
                   const void* queries, int nq, const uint32_t* mask, int64_t mask_stride_words, int k, int64_t id_base, float* out_scores,
                   int64_t* out_ids, cudaStream_t stream, int* launched, std::string* err) {
  *launched = 0;
  const int num_sms = tc5_num_sms(s);
  static const int no_pair = getenv("RS_DENSE_NO_PAIR") ? 1 : 0;  // A/B switch for scripts/batch_bench.py
  // pairs cover 256 queries; up to 128 queries fit one CTA's single M tile, where a pair's second CTA would idle
  const bool pair = !no_pair && num_sms >= 2 && nq > 128;
  const int mgroups = (nq + kDtMT * 128 - 1) / (kDtMT * 128);  // 256 queries per CTA / per CTA pair
  const int tiles_total = (int)((n + kDtBN - 1) / kDtBN);
  int ranges = (pair ? num_sms / 2 : num_sms) / mgroups;
  if (ranges < 1) ranges = 1;
  if (ranges > tiles_total) ranges = tiles_total;
  while ((long long)ranges * k > 16384) --ranges;  // rs_topk_merge limit

  CUtensorMap map_q, map_c;
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)nq};
    const uint64_t strides[1] = {(uint64_t)d * 2};
    const uint32_t box[2] = {kDtBK, 128};
    if (!tc5_encode(s, &map_q, dtype, 2, queries, dims, strides, box, err)) return -2;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)n};
    const uint64_t strides[1] = {(uint64_t)d * 2};
    const uint32_t box[2] = {kDtBK, (uint32_t)(pair ? kDtBN / 2 : kDtBN)};  // a pair's CTA stages half of the tile
    if (!tc5_encode(s, &map_c, dtype, 2, corpus, dims, strides, box, err)) return -2;
  }
  const size_t cand_bytes = (size_t)ranges * mgroups * (kDtMT * 128) * kDtCap * sizeof(uint64_t);
  const size_t ls_bytes = ((size_t)ranges * nq * k * sizeof(float) + 255) / 256 * 256;
  const size_t li_bytes = ((size_t)ranges * nq * k * sizeof(int64_t) + 255) / 256 * 256;
  const size_t gt_bytes = (size_t)ranges * nq * sizeof(uint32_t);
  uint8_t* ws = static_cast<uint8_t*>(tc5_dense_scratch(s, cand_bytes + ls_bytes + li_bytes + gt_bytes));
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
  kp.metric = metric;
  kp.num_ranges = ranges;
  kp.tiles_total = tiles_total;
  kp.gthr = reinterpret_cast<uint32_t*>(ws + cand_bytes + ls_bytes + li_bytes);
  static const int cross_off = getenv("RS_DENSE_NO_CROSS_THR") ? 1 : 0;  // A/B switch for scripts/batch_bench.py
  const int gm = (k + ranges - 1) / ranges;
  kp.gm = (ranges > 1 && gm <= kDtMaxGm && !cross_off) ? gm : 0;
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
  const size_t stage_bytes = pair ? (size_t)(kDtABytes + kDtBBytes / 2) : (size_t)kDtStageBytes;
  const size_t smem = 1024 + (size_t)(pair ? kDtStagesPair : kDtStages) * stage_bytes +
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
#define RS_DT_LAUNCH(BF, PR)                                                                                        \
  {                                                                                                                 \
    e = cudaFuncSetAttribute(dense_tc5_kernel<BF, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, dense_tc5_kernel<BF, PR>, map_q, map_c, kp);                 \
  }
  if (dtype == 1) {
    if (pair) RS_DT_LAUNCH(true, true) else RS_DT_LAUNCH(true, false)
  } else {
    if (pair) RS_DT_LAUNCH(false, true) else RS_DT_LAUNCH(false, false)
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
            "epilogue warp 4 (leader/peer): total %.0f/%.0f wait acc_full %.0f/%.0f compaction %.0f/%.0f | producer (leader/peer): "
            "total %.0f/%.0f wait empty %.0f/%.0f  [cycles, mean over CTAs]\n",
            (int)pair, ranges, mgroups, tiles_total / ranges, avg(0, 0), avg(1, 0), avg(2, 0), avg(4, 0), avg(4, 1), avg(5, 0),
            avg(5, 1), avg(6, 0), avg(6, 1), avg(11, 0), avg(11, 1), avg(10, 0), avg(10, 1));
  }
  e = launch_topk_merge(kp.list_scores, kp.list_ids, ranges, nq, k, k, 0, 0, out_scores, out_ids, stream);
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return -3;
  }
  *launched = 2;
  return 0;
}

}  // namespace rs
