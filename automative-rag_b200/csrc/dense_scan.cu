// dense_scan.cu — single-query exact top-k scan over a row-major fp16/bf16 corpus.
//
// Replaces the arithmetic behind QdrantStore.similarity_search_with_score
// (reference src/core/query/retrieval/vectorstore.py:166-214; Qdrant cosine search with a
// payload filter).  One query is a GEMV over the whole corpus, so the kernel is HBM-bound by
// construction: 2*d bytes per passing row, ~2 flops per byte.
//
// Shape of the kernel (B200: 148 SMs, one persistent CTA per SM):
//   * warp 13 is the TMA producer.  It takes runs of 1..32 consecutive mask words (one word = 32
//     consecutive rows) from an atomic counter, 32 at a time while plenty is left and fewer towards
//     the end, so a CTA that starts late or streams slower takes less.  The producer turns passing rows into TILES of TILE_ROWS rows (4 rows =
//     8 KB at d = 1024) in a shared-memory ring, up to 8 tiles per warp pass — lane l claims ring
//     slot l of the pass, waits for it and issues its copies, so the mbarrier round trips of the
//     pass overlap (a one-lane producer cost ~600 cycles per tile and capped the kernel at 5.7 TB/s):
//       - a word with >= 24 of its 32 rows passing is issued as contiguous tiles, ONE
//         cp.async.bulk each, with the tile's filter bits in the slot so the consumers skip the few
//         failing rows;
//       - otherwise the word's passing row ids are appended to a queue (32 words per warp pass:
//         a shuffle scan of the popcounts gives each lane its offset, then it walks its own set
//         bits) and GATHER tiles of TILE_ROWS rows are issued from the queue, one bulk copy per run
//         of consecutive rows.  Every gather tile is full no matter how sparse the filter is, and
//         rows that fail the filter are never read from HBM.
//     Each ring slot carries its row ids (or first row + filter bits) for the consumers.
//   * warps 0..NW-1 are consumers; warp w owns ring slots w and w+NW.  A lane reads 16-byte vectors
//     of the staged rows (LDS.128, conflict free) and FMAs them against the query, which stays
//     packed 16-bit in registers (FHFMA: 16-bit x 16-bit -> fp32 accumulate, exact products), and
//     the warp butterfly-reduces.  The ring is as deep as shared memory allows (26 x 8 KB = 208 KB
//     at k <= 128); with two slots per consumer one tile is in flight while the other is reduced.
//   * scores become order-preserving u64 keys and go through the CTA's TopKBuffer; a barrier
//     every few rounds decides whether to compact.  The stream ends with a padded round and a
//     round of END markers, so all consumer warps leave the loop in the same round.
//   * every CTA writes its sorted top-k to the workspace; the last CTA to finish (atomic
//     ticket) merges the grid's lists and writes the final (score, id) pairs — no second launch.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"
#include "topk_buffer.cuh"

namespace rs {

// One consumer warp per ring slot (slot w is always reduced by warp w).  mbarrier waits name a phase
// by PARITY, which only tells apart the current phase and the one before it; if slots were shared
// between warps a consumer could wait for a slot's use u+1 before use u has landed and be waved
// through on stale data (seen on hardware as an intermittent hang).  One owner per slot keeps every
// waiter at most one phase away from its barrier.
// A consumer may own two slots (w and w + NW), alternating: while it reduces one tile the other is
// in flight, so the time a slot spends being reduced stops subtracting from the bytes in flight.
constexpr int kScanMaxWarps = 13;                  // consumer warps
constexpr int kScanMaxStages = 2 * kScanMaxWarps;  // ring slots
constexpr int kScanProducerWarp = kScanMaxWarps;   // warps 0..12 consume, warp 13 produces
constexpr int kScanThreads = (kScanMaxWarps + 1) * 32;
constexpr int kConsumerBar = 1;   // named barrier id for the consumer threads
constexpr int kRowQueue = 2048;   // pending passing rows (power of two >= 32 words x 32 rows + a tile)
constexpr int kSlotContig = 0x100;  // slot_n flag: row ids are slot_rows[0] + lane (else slot_rows[lane])
constexpr int kSlotEnd = -1;
constexpr int kTournamentMaxK = 48;   // cross-CTA merge: tournament up to this k (~0.15 us per result + 1.5 us; the deployed
                                      // k of 20-40, mode_config.py), streamed merge above (11-13 us at k = 33..100)
constexpr int kTournamentLists = 5;   // lists per lane of the tournament warp (grid <= 160)
constexpr int kDenseWordBits = 24;  // words with >= 24 of 32 rows passing are staged whole (<= 25% extra bytes)

__device__ __forceinline__ uint64_t policy_evict_normal_() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// TMA tile::gather4: FOUR arbitrary rows of a 2-D tensor in one instruction.  The corpus is described to TMA as
// [n][row_bytes / 8] 8-byte elements (box = one whole row, <= 256 elements = 2 KB), so one gather4 moves four
// passing rows = a whole 8 KB gather tile at d = 1024 — a quarter of the bulk-copy instructions of the
// copy-per-row path, whose per-instruction cost is what holds sparse filters below the HBM rate.
__device__ __forceinline__ void tma_gather4(void* dst_smem, const CUtensorMap* map, uint32_t r0, uint32_t r1, uint32_t r2,
                                            uint32_t r3, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;" ::"r"(smem_u32(dst_smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(0), "r"((int)r0), "r"((int)r1), "r"((int)r2), "r"((int)r3), "l"(policy)
      : "memory");
}

// diagnostics: thread 0 stamps phase i of its CTA when a trace buffer is attached (rs_set_scan_trace)
__device__ __forceinline__ void scan_trace(const ScanParams& p, int i) {
  if (p.trace != nullptr && threadIdx.x == 0) {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[blockIdx.x * 8 + i] = t;
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(kScanThreads, 1)
    dense_scan_kernel(const ScanParams p, const __grid_constant__ CUtensorMap gmap) {
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
  int* slot_n = s_flag + 1;                                            // [kScanMaxStages]
  uint32_t* slot_rows = reinterpret_cast<uint32_t*>(slot_n + kScanMaxStages);  // [kScanMaxStages][32]
  uint32_t* rowq = slot_rows + kScanMaxStages * 32;                    // [kRowQueue]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NW = p.consumers;           // consumer warps in use; S is a multiple of NW
  const int consumer_threads = NW * 32;  // consumer warps >= NW have no slot: they leave at once

  // Programmatic dependent launch: let the NEXT query's scan (launched with the PDL attribute by
  // rs_dense_topk's query loop) take over each SM as soon as this CTA leaves it, so its launch,
  // prologue and first TMA round trip overlap this grid's tail and cross-CTA merge.  Consecutive
  // scans share nothing but the workspace, which alternates between two buffers; the
  // griddepcontrol.wait below (before this grid publishes into its buffer) orders grid N+2 after
  // grid N, the previous user of the same buffer.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  scan_trace(p, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  scan_trace(p, 1);

  if (warp == kScanProducerWarp) {
    // ------------------------------------------------------------------ TMA producer warp
    const uint64_t pol = p.l2_policy == 0 ? policy_evict_first() : (p.l2_policy == 1 ? policy_evict_normal_() : policy_evict_last());
    const uint32_t* mask = p.mask;
    const int64_t num_words = (p.n + 31) >> 5;
    // Work is handed out through one atomic counter per launch, GUIDED: a grab takes up to 32
    // consecutive mask words (one per lane, as many as a warp pass handles) while plenty of work is
    // left and shrinks to remaining/grid words (>= unit_words) towards the end.  CTAs that start
    // late (the SM that ran the previous query's merge) or sit on a slower path to HBM take fewer
    // words, and the grid finishes within about one word's time of each other.
    const int64_t wmin = p.unit_words, gmax = p.grab_max;
    // The first first_words words of every CTA are fixed (CTA b starts at word b * first_words
    // without a round trip to the counter); the counter hands out the words after grid * first_words.
    const int64_t dyn_base = (int64_t)p.first_words * gridDim.x;
    unsigned long long grab_pending = 0ull;
    int64_t seen = 0;     // latest counter value this CTA has observed
    int g_pending = 0;    // words requested by the outstanding grab
    auto grab_issue = [&]() {  // lane 0 takes the next words; the result is read later (grab_result)
      int64_t g = (num_words - seen) / (int64_t)gridDim.x;
      g = g < wmin ? wmin : (g > gmax ? gmax : g);
      g_pending = (int)g;
      if (lane == 0) grab_pending = atomicAdd(p.unit_counter, (unsigned long long)g);
    };
    auto grab_result = [&](int64_t& start, int& count) {
      start = dyn_base + (int64_t)__shfl_sync(0xFFFFFFFFu, grab_pending, 0);
      seen = start + g_pending;
      const int64_t left = num_words - start;
      count = left <= 0 ? 0 : (left < g_pending ? (int)left : g_pending);
    };
    const uint8_t* corpus = reinterpret_cast<const uint8_t*>(p.corpus);
    // Ring cursor, warp-uniform, kept incrementally (no 64-bit div/mod on the issue path):
    // the next tile goes to slot `slot` in ring pass `phase`; `posmod` = position % NW.
    int slot = 0, posmod = 0;
    uint32_t phase = 0;
    uint32_t head = 0, tail = 0;  // row queue (warp-uniform)
    const int batch_max = min(S, p.batch_max);  // contiguous tiles issued per warp pass, one per lane
    const int gather_batch = min(S, p.gather_batch);  // gather tiles per pass

    auto word_bits = [&](int64_t gw) -> uint32_t {  // filter bits of global mask word gw
      if (gw >= num_words) return 0u;
      const int64_t left = p.n - (gw << 5);
      const uint32_t in_range = left >= 32 ? 0xFFFFFFFFu : ((1u << (int)left) - 1u);
      return mask ? (__ldg(mask + gw) & in_range) : in_range;
    };
    // Lanes with `active` each claim one ring slot (consecutive slots in lane order) and wait, in
    // parallel, until its consumer has released it.  The single-lane version of this loop cost
    // ~600 cycles per tile (mbarrier round trips, 64-bit modulo) and capped the whole kernel.
    auto claim = [&](bool active, int& stage) {
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, active);
      const int idx = __popc(m & ((1u << lane) - 1u));
      const int cnt = __popc(m);
      stage = slot + idx;
      uint32_t ph = phase;
      if (stage >= S) {
        stage -= S;
        ph ^= 1u;
      }
      if (active) mbar_wait(&empty_bar[stage], ph ^ 1u);
      slot += cnt;
      if (slot >= S) {
        slot -= S;
        phase ^= 1u;
      }
      posmod = (posmod + cnt) % NW;
    };
    // lane l < count stages tile (row0 + l * tile_rows) whole if any of its rows pass (bits tb)
    auto issue_contig = [&](bool active, uint32_t row0, uint32_t tbits) {
      int stage;
      claim(active, stage);
      if (active) {
        slot_rows[stage * 32] = row0;
        slot_rows[stage * 32 + 1] = tbits;  // which of the staged rows pass the filter
        slot_n[stage] = tile_rows | kSlotContig;
        mbar_arrive_expect_tx(&full_bar[stage], tile_bytes);
        bulk_g2s(stage_base + (size_t)stage * tile_bytes, corpus + (size_t)row0 * row_bytes, tile_bytes, &full_bar[stage],
                 pol);
      }
      __syncwarp();
    };
    // Gather tiles: lane (t, j) = (lane / tile_rows, lane % tile_rows) copies queued row
    // head + t*nr + j into row j of tile t's slot — one bulk copy per RUN of consecutive rows.
    // With a row tensor map (p.gather4) a full tile goes out as one gather4 per four rows.
    const int lanes_per_tile_shift = 31 - __clz(tile_rows);
    const bool g4 = p.gather4 != 0;
    auto issue_gather = [&](int ntiles, int nr) {
      const int t = lane >> lanes_per_tile_shift, j = lane & (tile_rows - 1);
      const bool in = t < ntiles && j < nr;
      const bool leader = in && j == 0;
      int stage;
      claim(leader, stage);
      stage = __shfl_sync(0xFFFFFFFFu, stage, t << lanes_per_tile_shift);
      const uint32_t row = in ? rowq[(head + t * nr + j) & (kRowQueue - 1)] : 0xFFFFFFFFu;
      const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, row, 1);
      const bool start = !in || j == 0 || row != prev + 1u;  // lanes outside a tile terminate runs too
      const uint32_t starts = __ballot_sync(0xFFFFFFFFu, start);
      if (in) slot_rows[stage * 32 + j] = row;
      __syncwarp();
      if (leader) {
        slot_n[stage] = nr;
        mbar_arrive_expect_tx(&full_bar[stage], row_bytes * nr);
      }
      __syncwarp();
      const uint32_t r1 = __shfl_down_sync(0xFFFFFFFFu, row, 1);
      const uint32_t r2 = __shfl_down_sync(0xFFFFFFFFu, row, 2);
      const uint32_t r3 = __shfl_down_sync(0xFFFFFFFFu, row, 3);
      if (g4 && nr == tile_rows) {
        if (in && (j & 3) == 0)
          tma_gather4(stage_base + (size_t)stage * tile_bytes + (size_t)j * row_bytes, &gmap, row, r1, r2, r3,
                      &full_bar[stage], pol);
      } else if (in && start) {
        const uint32_t later = lane == 31 ? 0u : (starts >> (lane + 1));
        const int run = later ? __ffs(later) : (32 - lane);
        bulk_g2s(stage_base + (size_t)stage * tile_bytes + (size_t)j * row_bytes, corpus + (size_t)row * row_bytes,
                 row_bytes * run, &full_bar[stage], pol);
      }
      __syncwarp();
      head += ntiles * nr;
    };
    auto issue_marker = [&](int count, int marker) {  // lane l < count posts an empty (0) or END slot
      int stage;
      const bool active = lane < count;
      claim(active, stage);
      if (active) {
        slot_n[stage] = marker;
        mbar_arrive(&full_bar[stage]);
      }
      __syncwarp();
    };

    const uint32_t all_bits = tile_rows == 32 ? 0xFFFFFFFFu : ((1u << tile_rows) - 1u);
    const int tiles_per_word = 32 / tile_rows;
    // Two grabs are in flight ahead of the words being issued: the next grab's filter words are
    // loading and the grab after it is outstanding while the current words' tiles are issued.
    int64_t w_cur, w_nxt;
    int c_cur, c_nxt;
    seen = dyn_base;
    grab_issue();
    if (p.first_words > 0) {
      w_cur = (int64_t)blockIdx.x * p.first_words;
      c_cur = p.first_words;
    } else {
      grab_result(w_cur, c_cur);
      if (c_cur > 0) grab_issue();
    }
    uint32_t bits_cur = lane < c_cur ? word_bits(w_cur + lane) : 0u;
    while (c_cur > 0) {
      grab_result(w_nxt, c_nxt);
      const uint32_t bits_nxt = lane < c_nxt ? word_bits(w_nxt + lane) : 0u;
      if (c_nxt > 0) grab_issue();
      // ---- up to 32 words at once, one per lane
      const uint32_t w = bits_cur;  // 0 beyond the grab / the corpus
      const uint32_t row0 = (uint32_t)((w_cur + lane) << 5);
      const bool dense = __popc(w) >= kDenseWordBits && (int64_t)row0 + 32 <= p.n;
      // sparse words: append their passing row ids to the queue (exclusive scan of the popcounts
      // gives every lane its offset; each lane then walks its own set bits)
      const int mine = dense ? 0 : __popc(w);
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      if (mine) {
        uint32_t qpos = tail + (uint32_t)(incl - mine);
        uint32_t ww = w;
        while (ww) {
          const int bpos = __ffs(ww) - 1;
          ww &= ww - 1;
          rowq[(qpos++) & (kRowQueue - 1)] = row0 + bpos;
        }
      }
      tail += (uint32_t)total;
      __syncwarp();
      int full_tiles = (int)(tail - head) / tile_rows;
      while (full_tiles > 0) {
        const int nt = min(full_tiles, min(gather_batch, 32 >> lanes_per_tile_shift));
        issue_gather(nt, tile_rows);
        full_tiles -= nt;
      }
      // (nearly) full words: stage their tiles whole and let the consumers skip the few failing
      // rows — cheaper than gathering 24+ rows one by one.  Tiles with no passing row are skipped.
      uint32_t dense_lanes = __ballot_sync(0xFFFFFFFFu, dense);
      while (dense_lanes) {
        const int l = __ffs(dense_lanes) - 1;
        dense_lanes &= dense_lanes - 1;
        const uint32_t wl = __shfl_sync(0xFFFFFFFFu, w, l);
        const uint32_t rl = __shfl_sync(0xFFFFFFFFu, row0, l);
        for (int t0 = 0; t0 < tiles_per_word; t0 += batch_max) {
          const int t = t0 + lane;
          const bool in = lane < batch_max && t < tiles_per_word;
          const uint32_t tb = in ? ((wl >> (t * tile_rows)) & all_bits) : 0u;
          issue_contig(tb != 0u, rl + t * tile_rows, tb);
        }
      }
      bits_cur = bits_nxt;
      w_cur = w_nxt;
      c_cur = c_nxt;
    }
    if (tail != head) issue_gather(1, (int)(tail - head));
    while (posmod) issue_marker(min(NW - posmod, batch_max), 0);  // pad the last round
    for (int left = NW; left > 0; left -= batch_max) issue_marker(min(left, batch_max), kSlotEnd);  // END for all
  } else if (warp < NW) {
    // ------------------------------------------------------------------ consumer warps
    TopKBuffer buf{keys, thr, cnt, p.buf_cap, p.buf_hw, p.k, (int)threadIdx.x, consumer_threads, kConsumerBar};
    buf.init();

    // query stays packed 16-bit in registers; lane owns elements c*256 + lane*8 .. +7
    uint4 q[NCH];
    float qss = 0.f;
    {
      const uint4* qv = reinterpret_cast<const uint4*>(p.query);  // 16-byte aligned (checked by the ABI)
      const int nvq = p.d >> 3;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int vi = c * 32 + lane;
        q[c] = vi < nvq ? __ldg(qv + vi) : make_uint4(0, 0, 0, 0);
        qss = dot8<T>(q[c], q[c], qss);
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
    // appends per round <= warps * tile_rows; the buffer tolerates C - hw between checks
    const int rounds_per_check = p.rounds_per_check;

    const bool two_slots = S == 2 * NW;  // this warp alternates between slots warp and warp + NW
    scan_trace(p, 2);
    for (int r = 0;; ++r) {
      const int stage = warp + ((two_slots && (r & 1)) ? NW : 0);
      mbar_wait(&full_bar[stage], (uint32_t)((two_slots ? (r >> 1) : r) & 1));
      if (r == 0) scan_trace(p, 3);
      const int sn = slot_n[stage];
      if (sn == kSlotEnd) {  // all consumer warps see END in the same round
        break;
      }
      const int nr = sn & 0xFF;
      if (nr > 0) {
        const bool contig = (sn & kSlotContig) != 0;
        uint32_t row = 0;
        if (lane < nr) row = contig ? slot_rows[stage * 32] + lane : slot_rows[stage * 32 + lane];
        const uint32_t tbits = contig ? slot_rows[stage * 32 + 1] : 0xFFFFFFFFu;
        const bool live = lane < nr && ((tbits >> lane) & 1u);
        float inv = 1.f;
        if (inv_norm != nullptr && live) inv = __ldg(inv_norm + row);
        float my_score = 0.f;
        // ---- rows staged in shared memory by TMA
        const uint8_t* my_stage = stage_base + (size_t)stage * tile_bytes;
        for (int i0 = 0; i0 < nr; i0 += 4) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u;
            if (i < nr && ((tbits >> i) & 1u)) {
              const uint4* rowp = reinterpret_cast<const uint4*>(my_stage + (size_t)i * row_bytes);
#pragma unroll
              for (int c = 0; c < NCH; ++c) {
                const int v = c * 32 + lane;
                if (v < nvec) acc[u] = dot8<T>(rowp[v], q[c], acc[u]);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float s = warp_sum(acc[u]);
            if (lane == i0 + u) my_score = s;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);  // slot may be refilled
        const float score = my_score * inv * q_scale;
        const uint64_t key = make_key(score, row);
        buf.warp_append(live && key > buf.threshold(), key);
      } else {
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
      }
      if ((r + 1) % rounds_per_check == 0) buf.maybe_compact();
    }
    scan_trace(p, 4);
    buf.compact();  // final: keys[0..k) sorted descending (0 = empty)
    scan_trace(p, 5);

    // ------------------------------------------------------------------ cross-CTA merge
    asm volatile("griddepcontrol.wait;" ::: "memory");  // previous grid fully done (no-op without PDL)
    uint64_t* ws = p.ws_keys + (size_t)blockIdx.x * p.k;
    for (int i = threadIdx.x; i < p.k; i += consumer_threads) ws[i] = keys[i];
    __threadfence();
    named_bar_sync(kConsumerBar, consumer_threads);
    if (threadIdx.x == 0) {
      unsigned ticket = atomicAdd(p.ticket, 1u);
      *s_flag = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    named_bar_sync(kConsumerBar, consumer_threads);
    scan_trace(p, 6);
    if (*s_flag && p.k <= kTournamentMaxK && gridDim.x <= 32 * kTournamentLists &&
        (size_t)gridDim.x * p.k * sizeof(uint64_t) <= (size_t)S * tile_bytes) {
      // ---- small k: tournament.  Every CTA's sorted list is copied into the (now idle) tile ring
      // with one round of independent loads; then ONE warp pops the k winners: lane l holds the heads
      // of lists l, l+32, ...; each step is a warp max (two REDUX) and one shared-memory load by the
      // winning lane.  ~60 ns per result instead of a barrier-separated compaction per key.
      __threadfence();
      uint64_t* all = reinterpret_cast<uint64_t*>(stage_base);
      const int total = (int)gridDim.x * p.k;
      for (int i0 = threadIdx.x; i0 < total; i0 += consumer_threads * 4) {
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * consumer_threads;
          v[u] = i < total ? __ldcg(p.ws_keys + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * consumer_threads;
          if (i < total) all[i] = v[u];
        }
      }
      named_bar_sync(kConsumerBar, consumer_threads);
      if (warp == 0) {
        uint64_t head[kTournamentLists];
        int pos[kTournamentLists];
        uint64_t best = 0ull;  // this lane's largest head; only the winning lane's changes per step
#pragma unroll
        for (int j = 0; j < kTournamentLists; ++j) {
          const int list = lane + 32 * j;
          pos[j] = list * p.k;
          head[j] = list < (int)gridDim.x ? all[pos[j]] : 0ull;
          best = head[j] > best ? head[j] : best;
        }
        for (int t = 0; t < p.k; ++t) {
          // keys are unique (the row id is part of the key), so exactly one lane holds the maximum
          const uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, (uint32_t)(best >> 32));
          const uint32_t lo = __reduce_max_sync(0xFFFFFFFFu, (uint32_t)(best >> 32) == hi ? (uint32_t)best : 0u);
          const uint64_t win = ((uint64_t)hi << 32) | lo;
          if (win == 0ull) {  // every list exhausted: fewer than k rows passed the filter
            for (int i = t + lane; i < p.k; i += 32) {
              p.out_scores[i] = -INFINITY;
              p.out_ids[i] = -1;
            }
            break;
          }
          if (best == win) {
            p.out_scores[t] = key_score(win);
            p.out_ids[t] = p.id_base + (int64_t)key_row(win);
            best = 0ull;
#pragma unroll
            for (int j = 0; j < kTournamentLists; ++j) {
              if (head[j] == win) {
                const int list = lane + 32 * j;
                ++pos[j];
                head[j] = pos[j] < (list + 1) * p.k ? all[pos[j]] : 0ull;
              }
              best = head[j] > best ? head[j] : best;
            }
          }
        }
        if (lane == 0) {  // ready for the next launch that uses this workspace / this counter
          *p.ticket = 0u;
          *p.unit_counter = 0ull;
        }
      }
      scan_trace(p, 7);
    } else if (*s_flag) {
      __threadfence();
      // The buffer already holds this CTA's own top-k with the matching threshold; stream the
      // other CTAs' sorted lists through it.  Thread t walks list t (+lists_per_pass, ...); a list is
      // abandoned at its first key <= threshold (lists are sorted).
      // First a bound from the lists themselves: with j = ceil(k / lists) - 1, every list holds j + 1 keys >= its
      // own j-th entry, so at least lists * (j + 1) >= k keys are >= T = the minimum of those entries and nothing
      // below T can be in the answer.  Every CTA saw a random 1/148 of the corpus, so T sits close to the true k-th
      // score and only a few thousand of the 148 * k keys are streamed at all (k = 1000: the merge took ~110 us of a
      // 420 us launch when the threshold had to climb from this CTA's own k-th entry).  A list with fewer than
      // j + 1 entries (key 0) switches the bound off.
      {
        const int j = (p.k + (int)gridDim.x - 1) / (int)gridDim.x - 1;
        uint64_t* tmin = reinterpret_cast<uint64_t*>(stage_base);  // the tile ring is idle by now
        if (threadIdx.x == 0) *tmin = ~0ull;
        named_bar_sync(kConsumerBar, consumer_threads);
        uint64_t v = ~0ull;
        for (int l = threadIdx.x; l < (int)gridDim.x; l += consumer_threads) {
          const uint64_t e = __ldcg(p.ws_keys + (size_t)l * p.k + j);
          v = e < v ? e : v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const uint64_t w = __shfl_xor_sync(0xFFFFFFFFu, v, o);
          v = w < v ? w : v;
        }
        if (lane == 0) atomicMin(reinterpret_cast<unsigned long long*>(tmin), (unsigned long long)v);
        named_bar_sync(kConsumerBar, consumer_threads);
        const uint64_t T = *reinterpret_cast<volatile uint64_t*>(tmin);
        named_bar_sync(kConsumerBar, consumer_threads);
        if (T != 0ull && T != ~0ull) buf.raise_floor(T - 1ull);
      }
      const int lists_per_pass = min(consumer_threads, buf.slack());  // appends between two checks
      constexpr int KB = 8;  // keys fetched per thread per batch: KB independent L2 loads, one latency
      for (int base = 0; base < (int)gridDim.x; base += lists_per_pass) {
        const int list = base + threadIdx.x;
        bool active = (int)threadIdx.x < lists_per_pass && list < (int)gridDim.x && list != (int)blockIdx.x;
        const uint64_t* lp = p.ws_keys + (size_t)list * p.k;
        for (int j0 = 0; j0 < p.k; j0 += KB) {
          uint64_t kb[KB];
#pragma unroll
          for (int u = 0; u < KB; ++u) kb[u] = (active && j0 + u < p.k) ? __ldcg(lp + j0 + u) : 0ull;
#pragma unroll
          for (int u = 0; u < KB; ++u) {
            if (j0 + u < p.k) {  // uniform
              if (active && kb[u] <= buf.threshold()) active = false;  // sorted list: nothing further can enter
              buf.warp_append(active, kb[u]);
              // after the list HEADS the threshold must rise at once (k-th best of own list + all
              // heads already bounds the answer from below), so compact unconditionally there
              if (base == 0 && j0 == 0 && u == 0 && buf.floor_key == 0ull)
                buf.compact();
              else
                buf.maybe_compact();
            }
          }
          if (!named_bar_or(kConsumerBar, consumer_threads, active)) break;  // every list exhausted
        }
      }
      buf.compact();
      for (int i = threadIdx.x; i < p.k; i += consumer_threads) {
        uint64_t key = keys[i];
        if (key == 0ull) {
          p.out_scores[i] = -INFINITY;
          p.out_ids[i] = -1;
        } else {
          p.out_scores[i] = key_score(key);
          p.out_ids[i] = p.id_base + (int64_t)key_row(key);
        }
      }
      if (threadIdx.x == 0) {  // ready for the next launch that uses this workspace / this counter
        *p.ticket = 0u;
        *p.unit_counter = 0ull;
      }
      scan_trace(p, 7);
    }
  }
}

template <typename T>
static cudaError_t launch_scan_t(const ScanParams& p, const CUtensorMap& gmap, int grid, size_t smem, bool pdl,
                                 cudaStream_t stream) {
  const int nch = (p.d + 255) / 256;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kScanThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
#define RS_SCAN_CASE(N)                                                                                      \
  {                                                                                                          \
    cudaError_t e = cudaFuncSetAttribute(dense_scan_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                         \
    if (e != cudaSuccess) return e;                                                                          \
    return cudaLaunchKernelEx(&cfg, dense_scan_kernel<T, N>, p, gmap);                                            \
  }
  if (nch <= 1) RS_SCAN_CASE(1)
  if (nch <= 2) RS_SCAN_CASE(2)
  if (nch <= 4) RS_SCAN_CASE(4)
  if (nch <= 8) RS_SCAN_CASE(8)
  RS_SCAN_CASE(16)
#undef RS_SCAN_CASE
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

int scan_tile_rows(int d) {
  // tuning knobs for experiments (scripts/scan_sweep.py); the defaults are the measured optimum
  static const int tile_bytes_target = env_int("RS_SCAN_TILE_BYTES", 8192);
  int rows = tile_bytes_target / (d * 2);
  int tr = 1;
  while (tr * 2 <= rows && tr < 32) tr <<= 1;
  return tr;
}

// Shared-memory plan of a launch: the top-k buffer first, the tile ring in what is left.
struct ScanPlan {
  int tile_rows, consumers, stages, buf_cap, buf_hw, rounds_per_check;
  size_t smem;
};

// coresident: size the CTA for HALF an SM.  Inside one rs_dense_topk call the launches are chained (programmatic
// dependent launch), and a full-SM CTA leaves its SM idle from its final compaction to the next query's first tile
// (~7 us per query: compaction, publish, launch, barrier init, prologue, first TMA round trip).  With two CTAs per SM
// query j+1 streams while query j finishes and starts: the same 24 tiles in flight per SM, no idle gap —
// 1M x 1024, 64 queries per call: 293 -> 279 us per query; 125 k rows (one GPU's share of 1M at 8 GPUs): the fixed
// cost was a quarter of the launch.  A lone launch on half a ring is much slower (305 -> 416 us), so the plan is
// used for every launch of a multi-query call and never for a single-query call.
static ScanPlan scan_plan(int d, int k, bool coresident = false) {
  static const int max_slots = env_int("RS_SCAN_STAGES", kScanMaxStages);
  ScanPlan pl{};
  pl.tile_rows = scan_tile_rows(d);
  // the buffer is checked every `rounds_per_check` consumer rounds; a round appends at most one key per staged
  // row (<= 13 warps x tile_rows): about 256 appends between checks
  const int per_round = kScanMaxWarps * pl.tile_rows;
  pl.rounds_per_check = 256 / per_round > 1 ? 256 / per_round : 1;
  const int slack = per_round * pl.rounds_per_check;
  pl.buf_cap = TopKBuffer::capacity_for(k, slack);
  pl.buf_hw = pl.buf_cap - slack;
  const size_t fixed = (size_t)pl.buf_cap * 8 + 2 * kScanMaxStages * 8 + 8 + 16 + kScanMaxStages * 4 +
                       kScanMaxStages * 32 * 4 + kRowQueue * 4;
  const size_t tile_bytes = (size_t)pl.tile_rows * d * 2;
  // leave 1 KB per CTA for the runtime's reserved shared memory (228 KB per SM in all)
  const size_t budget = coresident ? (228 * 1024) / 2 - 1024 - 512 : 227 * 1024 - 1024;
  int s = budget > fixed ? (int)((budget - fixed) / tile_bytes) : 0;
  s = s > kScanMaxStages ? kScanMaxStages : s;
  s = s > max_slots ? max_slots : s;
  if (s < 2) s = 2;
  // Two slots per consumer (one tile in flight while the other is reduced) as soon as 8 consumers can have
  // them: 11 consumers x 2 slots beat 13 x 1 — bytes in flight matter more than the two extra warps.
  if (s >= 16) {
    pl.consumers = s / 2 < kScanMaxWarps ? s / 2 : kScanMaxWarps;
    pl.stages = 2 * pl.consumers;
  } else {
    pl.consumers = s < kScanMaxWarps ? s : kScanMaxWarps;
    pl.stages = pl.consumers;
  }
  pl.smem = (size_t)pl.stages * tile_bytes + fixed;
  return pl;
}

// Two CTAs per SM need <= 73 registers per thread at 448 threads (the d <= 1024 instantiations; tests/test_abi.py
// checks the built library) and a ring that still covers the HBM latency: >= 12 slots of 8 KB in each half.
bool scan_coresident_ok(int d, int k) {
  static const bool off = getenv("RS_SCAN_NO_CORESIDENT") != nullptr;
  if (off || d > 1024) return false;
  const ScanPlan pl = scan_plan(d, k, true);
  return pl.stages >= 12 && (size_t)pl.tile_rows * d * 2 >= 4096 && pl.smem <= (228 * 1024) / 2 - 1024;
}

int scan_stages(int d, int k) { return scan_plan(d, k).stages; }

void scan_plan_query(int d, int k, int64_t out[7], bool chained) {
  const bool co = chained && scan_coresident_ok(d, k);
  const ScanPlan pl = scan_plan(d, k, co);
  out[0] = pl.tile_rows;
  out[1] = pl.consumers;
  out[2] = pl.stages;
  out[3] = pl.buf_cap;
  out[4] = pl.buf_hw;
  out[5] = pl.rounds_per_check;
  out[6] = (int64_t)pl.smem;
  if (chained) out[7] = co ? 1 : 0;
}

size_t scan_smem_bytes(int d, int k) { return scan_plan(d, k).smem; }

// Smallest grab, in mask words: at least 64 KB of rows, so that the counter's round trip (~1 us)
// stays hidden behind the ring (26 x 8 KB already issued ahead) also at the very end.
static int scan_unit_words(int d) {
  static const int forced = env_int("RS_SCAN_UNIT_WORDS", 0);
  if (forced > 0) return forced > 32 ? 32 : forced;
  const int64_t word_bytes = 32ll * d * 2;
  const int64_t lo = (64 * 1024 + word_bytes - 1) / word_bytes;
  return lo > 32 ? 32 : (int)lo;
}

bool scan_gather4_supported(int d) {
  // one box = one whole row of <= 256 8-byte elements; a gather4 lands four rows = a multiple of 128 bytes
  return d % 16 == 0 && d <= 1024 && scan_tile_rows(d) >= 4;
}

cudaError_t launch_dense_scan(ScanParams p, int dtype, int num_sms, bool pdl, cudaStream_t stream,
                              const CUtensorMap* gather_map, bool chained) {
  const bool co = chained && scan_coresident_ok(p.d, p.k);
  const ScanPlan pl = scan_plan(p.d, p.k, co);
  // Ring slots the producer claims per pass.  It waits for ALL of them before issuing any, so on the 12-slot half
  // ring a pass of 8 gather tiles (four random rows each: the longest round trip) stalls on the slowest slot of two
  // thirds of the ring: p = 0.5 / 0.25 / 0.1 ran at 0.82 / 0.81 / 0.78 of HBM with 8 and at 1.00 / 0.99 / 0.87 with 4
  // (2: 1.03 / 1.00 / 0.90 but p = 0.03 0.74 -> 0.64).  The 26-slot ring and contiguous tiles are best with 8
  // (profiles/r02_scan_coresident_ab.txt).
  static const int batch_env = env_int("RS_SCAN_BATCH", 0), gbatch_env = env_int("RS_SCAN_GATHER_BATCH", 0);
  p.batch_max = batch_env > 0 ? (batch_env > 8 ? 8 : batch_env) : 8;
  p.gather_batch = gbatch_env > 0 ? (gbatch_env > 8 ? 8 : gbatch_env) : (co ? 4 : 8);
  alignas(64) CUtensorMap gmap;
  if (gather_map != nullptr && p.mask != nullptr && scan_gather4_supported(p.d)) {
    gmap = *gather_map;
    p.gather4 = 1;
  } else {
    memset(&gmap, 0, sizeof(gmap));
    p.gather4 = 0;
  }
  p.tile_rows = pl.tile_rows;
  p.consumers = pl.consumers;
  p.stages = pl.stages;
  p.buf_cap = pl.buf_cap;
  p.buf_hw = pl.buf_hw;
  p.rounds_per_check = pl.rounds_per_check;
  static const int l2_policy = env_int("RS_SCAN_L2_POLICY", 0);
  p.l2_policy = l2_policy;
  const int64_t num_words = (p.n + 31) / 32;  // a mask word covers 32 rows
  p.unit_words = scan_unit_words(p.d);
  static const int grab_max = env_int("RS_SCAN_GRAB_MAX", 32);
  p.grab_max = grab_max < p.unit_words ? p.unit_words : (grab_max > 32 ? 32 : grab_max);
  const int64_t num_units = (num_words + p.unit_words - 1) / p.unit_words;
  int grid = (int)(num_units < (int64_t)num_sms ? (num_units > 0 ? num_units : 1) : num_sms);
  // static head start: about a quarter of each CTA's share, whole grabs only
  int64_t first = num_words / ((int64_t)grid * 4);
  first = first > p.grab_max ? p.grab_max : first;
  p.first_words = first < p.unit_words ? 0 : (int32_t)first;
  const size_t smem = pl.smem;
  if (dtype == 0) return launch_scan_t<__half>(p, gmap, grid, smem, pdl, stream);
  return launch_scan_t<__nv_bfloat16>(p, gmap, grid, smem, pdl, stream);
}

}  // namespace rs
