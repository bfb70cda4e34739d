// maxsim_tc5.cu — shared-candidate ColBERT MaxSim on tcgen05 tensor cores with TMEM
// accumulators: the batched shape of ColBERTReranker.batch_rerank_queries
// (reference src/core/query/llm/rerankers.py:583-593 — documents encoded once, every query
// scored against the same list with _compute_maxsim_scores :215-265).
//
// The stage is one contraction S = Q_all [nq*lq, d] . D_all^T [d, n_tokens] followed by a
// segmented max over each document's tokens and a weighted sum over each query's tokens.
// Mapping onto the hardware:
//   * M (TMEM lanes) = query tokens, N (TMEM columns) = document tokens; K = d (64 or 128) fits one
//     shared-memory stage, so there is no K pipeline.  tcgen05.ld 32x32b hands thread t of an
//     epilogue warp ONE query-token row with 32 consecutive document tokens in registers, so the
//     max over document tokens is a register-only FMNMX3 chain and the sum over a query's 32 tokens
//     is one warp shuffle reduction.  The token-score matrix S exists only in TMEM.
//   * CTA PAIRS (cta_group::2, a 2-wide cluster = the two SMs of a TPC): tcgen05.mma of M = 256;
//     each CTA stages its own 128 query rows and HALF of the document tokens of a tile (the tensor
//     cores read the other half from the peer's shared memory) and holds its 128 rows of the
//     accumulator in its own TMEM.  A pair keeps MG = 2 "pair tiles" (2 x 256 query rows) resident,
//     so every 64 KB document tile feeds 512 query rows and each SM pulls only 32 KB of it.
//   * a 256-token stage is consumed as TWO N = 128 MMA groups per pair tile, each into its own
//     128-column TMEM slot (slot = pair tile * 2 + half): four slots rotate, so an epilogue set has
//     1536 cycles to drain a slot before the tensor cores need it again.  Each CTA stages tokens
//     [64c, 64c+64) and [128+64c, 128+64c+64) of a tile (c = its rank), which makes the columns of
//     both half-tile accumulators contiguous in token order.
//   * warp roles (both CTAs): 0 = TMA producer (one lane per box), 1 = MMA issuer (leader CTA
//     only, one elected thread), 2 = TMEM allocator, 4..7 = epilogue set 0 (pair tile 0), 8..11 =
//     epilogue set 1 (warp % 4 selects the TMEM lane quarter).  TMA loads of both CTAs count on the
//     LEADER's barriers; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs;
//     the epilogues hand slots back with a (remote) arrive on the leader's barrier.
//   * work split: the documents are cut into F PHASES of equal token counts (about 24 MB each, so a
//     phase stays L2-resident while the query groups re-read it); inside a phase the
//     (query group, token) space is cut into one equal piece per pair, a piece being one or two
//     SEGMENTS (group, first doc, last doc) cut at document starts.  Every pair gets the same number of
//     tokens in every phase, whatever the number of groups.  The segment list of a pair is computed once
//     at kernel start (warp-parallel 32-ary searches over doc_offsets) and kept in shared memory.
//   * THE EPILOGUE PACES THIS KERNEL (K = 128 is only 1024 tensor cycles per 128 x 256 accumulator):
//     with two epilogue warps per scheduler every instruction costs ~4-5 cycles of issue latency, so
//     what counts is instructions per 32 scores.  The hot path is one tcgen05.ld, one wait, 16
//     three-input max ops and one compare; a document boundary inside a chunk (one per ~9 chunks at
//     300 tokens) is split with a switch on the boundary's 4-column group (static register indices,
//     ~35 instructions); document ends are prefetched 32 at a time (one per lane, fetched by shuffle).
//     History (profiles/r02_maxsim_trace_*.txt): predicated 32-deep max chains at every boundary and a
//     per-chunk piece scheduler both cost ~80 instructions per chunk = ~4000 cycles per drain, 4x what
//     the tensor cores need.
//   * query-token sums that span several warps (lq > 32) are combined through shared memory in a
//     fixed order — results are run-to-run identical, like every other kernel of the library.
#include <cuda.h>
#include <math_constants.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "tc5.cuh"
#include "tc5_host.h"

namespace rs {

constexpr int kTcThreads = 384;
constexpr int kTcTile = 256;    // document tokens per shared-memory stage; each CTA of the pair stages half
constexpr int kTcBN = 128;      // document tokens per MMA group (UMMA N) == columns of a TMEM slot
constexpr int kTcMG = 2;        // resident pair tiles (256 query rows each)
constexpr int kTcSlots = 4;     // TMEM slots: pair tile x half of the stage
constexpr int kTcStages = 4;    // B ring depth (each stage: 128 tokens x d per CTA)
constexpr int kTcTmemCols = 512;
constexpr int kTcEpiBar = 2;    // named barriers 2, 3: the four warps of epilogue set 0 / 1
constexpr int kTcMaxSegs = 64;  // segments per pair (phases x (groups per piece + 1))
constexpr int kTcMaxPhases = 6;

struct MaxSimTcParams {
  const float* q_weight;
  const int32_t* doc_offsets;
  float* out;
  int32_t nq, lq, lq_pad, nd;
  int32_t num_pair_tiles;  // ceil(128-row query tiles / 2)
  int32_t num_mgroups;     // G: groups of kTcMG pair tiles
  int32_t phases;          // F: document phases (L2-sized)
  long long* trace;        // diagnostics (RS_MAXSIM_TRACE=1): [grid][16] cycle counters per role, or null
};

struct TcSeg {
  int g, d0, d1;
};

// wait (+ cycles spent waiting, only when a trace buffer is attached: two clock reads per wait are not free for the
// single MMA-issuing thread)
__device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, long long& acc, bool tracing) {
  if (!tracing) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t;
}

__device__ __forceinline__ float tc_reference_weight(int i, int lq) {
  return (lq > 2 && (i == 0 || i == lq - 1)) ? 0.f : 1.f;  // rerankers.py:255-261
}

__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// first index i in [lo, hi) with off[i] >= want, or hi; off[] non-decreasing.  One warp, 32-ary search.
__device__ __forceinline__ int warp_lower_bound(const int32_t* off, int lo, int hi, long long want, int lane) {
  while (hi - lo > 32) {
    const int step = (hi - lo + 31) / 32;
    const int i = lo + lane * step;
    const bool less = (i < hi) && ((long long)__ldg(off + i) < want);
    const int cnt = __popc(__ballot_sync(0xFFFFFFFFu, less));  // probes below `want` form a prefix
    if (cnt == 0) return lo;
    const int nhi = min(hi, lo + cnt * step);
    lo = lo + (cnt - 1) * step + 1;
    hi = nhi;
  }
  const int i = lo + lane;
  const bool less = (i < hi) && ((long long)__ldg(off + i) < want);
  return lo + __popc(__ballot_sync(0xFFFFFFFFu, less));
}

// l = max v[0 .. b), r = max v[b .. 32) for a boundary b in [8J, 8J + 8): static register indices everywhere,
// only the eight columns around the boundary are predicated.
template <int J>
__device__ __forceinline__ void tc_split8(const uint32_t (&v)[32], int b, float& l, float& r) {
  float l0 = -CUDART_INF_F, l1 = -CUDART_INF_F, r0 = -CUDART_INF_F, r1 = -CUDART_INF_F;
#pragma unroll
  for (int c = 0; c < 8 * J; c += 2) {
    l0 = fmaxf(l0, __uint_as_float(v[c]));
    l1 = fmaxf(l1, __uint_as_float(v[c + 1]));
  }
#pragma unroll
  for (int c = 8 * J + 8; c < 32; c += 2) {
    r0 = fmaxf(r0, __uint_as_float(v[c]));
    r1 = fmaxf(r1, __uint_as_float(v[c + 1]));
  }
#pragma unroll
  for (int c = 8 * J; c < 8 * J + 8; c += 2) {
    const float x = __uint_as_float(v[c]), y = __uint_as_float(v[c + 1]);
    l0 = fmaxf(l0, c < b ? x : -CUDART_INF_F);
    r0 = fmaxf(r0, c < b ? -CUDART_INF_F : x);
    l1 = fmaxf(l1, c + 1 < b ? y : -CUDART_INF_F);
    r1 = fmaxf(r1, c + 1 < b ? -CUDART_INF_F : y);
  }
  l = fmaxf(l0, l1);
  r = fmaxf(r0, r1);
}

template <bool BF16, int KH>
__global__ void __launch_bounds__(kTcThreads, 1)
    maxsim_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                      const MaxSimTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kABytesKH = 128 * 128;            // one K half (64 elements) of this CTA's 128 query rows
  constexpr uint32_t kBBytesKH = (kTcTile / 2) * 128;  // one K half of this CTA's 128 document tokens of a stage
  constexpr uint32_t kBoxBytes = 64 * 128;             // one TMA box of the document stream: 64 tokens x 64 elements
  constexpr uint32_t kABytes = kABytesKH * KH;
  constexpr uint32_t kBBytes = kBBytesKH * KH;

  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), num_pairs = (int)(gridDim.x >> 1);
  const int G = p.num_mgroups, F = p.phases;

  // ---- shared memory carve-up (1024-byte aligned: SWIZZLE_128B atoms)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* smA = sm;                                 // [kTcMG][KH][128 rows x 128 B]
  uint8_t* smB = smA + kTcMG * kABytes;              // [kTcStages][KH][2 halves x 64 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + kTcStages * kBBytes);
  uint64_t* a_full = bars;                 // 1: query tiles of the current segment have landed (leader's counts both CTAs)
  uint64_t* a_empty = bars + 1;            // 1: every MMA of the segment has read them
  uint64_t* b_full = bars + 2;             // kTcStages (leader's)
  uint64_t* b_empty = b_full + kTcStages;  // kTcStages
  uint64_t* acc_full = b_empty + kTcStages;    // kTcSlots
  uint64_t* acc_empty = acc_full + kTcSlots;   // kTcSlots (leader's collect both CTAs' epilogue warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + kTcSlots);
  int* s_nseg = reinterpret_cast<int*>(tmem_ptr + 2);
  int* s_cut = s_nseg + 2;                                   // [kTcMaxPhases + 1] first document of each phase
  int* s_end = s_cut + kTcMaxPhases + 2;                     // [kTcMaxPhases][2 ends][group, document] of this pair's pieces
  TcSeg* segs = reinterpret_cast<TcSeg*>(s_end + 4 * kTcMaxPhases);  // [kTcMaxSegs]
  float* parts = reinterpret_cast<float*>(segs + kTcMaxSegs);        // [2 sets][2 buffers][4 quarters] partial sums (lq > 32)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tracing = p.trace != nullptr;

  if (warp == 0 && lane == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < kTcSlots; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 2 * 4);  // the 4 warps of an epilogue set, in both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_cta2(tmem_ptr, kTcTmemCols);
    tmem_relinquish_cta2();
  }
  // ---- this pair's segments.  Phase f = documents [cut[f], cut[f+1]) (equal token counts); inside a phase the linear
  // space (group g, token) is cut into num_pairs equal pieces at document starts.  Both CTAs compute the same list.
  const long long T = __ldg(p.doc_offsets + p.nd);
  if (warp >= 4 && warp - 4 < F - 1) {  // round 1: the F - 1 inner phase cuts, one warp each
    const int f = warp - 4 + 1;
    const int c = warp_lower_bound(p.doc_offsets, 0, p.nd, T * f / F, lane);
    if (lane == 0) s_cut[f] = c;
  }
  if (threadIdx.x == 0) {
    s_cut[0] = 0;
    s_cut[F] = p.nd;
  }
  __syncthreads();
  if (warp < 2 * F) {  // round 2: both ends of this pair's piece in every phase, one warp each
    const int f = warp >> 1, end = warp & 1;
    const int D0 = s_cut[f], D1 = s_cut[f + 1];
    const long long t_lo = __ldg(p.doc_offsets + D0), Tf = (long long)__ldg(p.doc_offsets + D1) - t_lo;
    int g = 0, d = D0;
    if (Tf > 0) {
      const long long pos = (long long)G * Tf * (pair + end) / num_pairs;
      g = (int)(pos / Tf);
      d = warp_lower_bound(p.doc_offsets, D0, D1, t_lo + pos % Tf, lane);
    }
    if (lane == 0) {
      s_end[(f * 2 + end) * 2] = g;
      s_end[(f * 2 + end) * 2 + 1] = d;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int f = 0; f < F; ++f) {
      const int D0 = s_cut[f], D1 = s_cut[f + 1];
      if (D0 >= D1) continue;
      const int g_lo = s_end[(f * 2) * 2], d_lo = s_end[(f * 2) * 2 + 1];
      const int g_hi = s_end[(f * 2 + 1) * 2], d_hi = s_end[(f * 2 + 1) * 2 + 1];
      for (int g = g_lo; g <= g_hi && g < G; ++g) {
        const int a = g == g_lo ? d_lo : D0, b = g == g_hi ? d_hi : D1;
        if (a < b && n < kTcMaxSegs) segs[n++] = TcSeg{g, a, b};
      }
    }
    s_nseg[0] = n;
  }
  tc5_fence_before();
  cluster_sync_all();
  tc5_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr);
  const int nseg = s_nseg[0];
  const int qpt = 128 / p.lq_pad;  // queries per 128-row tile

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // lane = half * KH + kh loads one 64-token box of every stage: tokens [128 half + 64 rank, +64) of the tile, K half
    // kh.  The query tiles of a segment are loaded by lanes 0 .. KH-1.
    if (lane < 2 * KH && nseg > 0) {
      const int kh = lane % KH, half = lane / KH;
      tma_prefetch_desc(lane == 0 ? &map_q : &map_d);
      const uint64_t pol = policy_evict_normal();
      const uint32_t lead_a_full = mapa_u32(smem_u32(a_full), 0);
      int j = 0;  // stages issued so far (ring position)
      const long long T0 = clock64();
      long long w_be = 0;
      for (int si = 0; si < nseg; ++si) {
        const TcSeg sg = segs[si];
        const int n_act = min(kTcMG, p.num_pair_tiles - sg.g * kTcMG);
        const int tok0 = __ldg(p.doc_offsets + sg.d0);
        const int ntiles = (__ldg(p.doc_offsets + sg.d1) - tok0 + kTcTile - 1) / kTcTile;
        if (half == 0) {
          mbar_wait(a_empty, ((uint32_t)si & 1u) ^ 1u);  // previous segment's MMAs are done with A (both CTAs)
          if (rank == 0 && lane == 0) mbar_arrive_expect_tx(a_full, 2u * (uint32_t)n_act * kABytes);
          for (int a = 0; a < n_act; ++a)
            tma_load_3d_cta2(smA + a * kABytes + kh * kABytesKH, &map_q, kh * 64, 0,
                             ((sg.g * kTcMG + a) * 2 + (int)rank) * qpt, lead_a_full, pol);
        }
        for (int t = 0; t < ntiles; ++t, ++j) {
          const int s = j % kTcStages;
          const uint32_t ph = (uint32_t)(j / kTcStages) & 1u;
          timed_wait(&b_empty[s], ph ^ 1u, w_be, tracing);
          if (rank == 0 && lane == 0) mbar_arrive_expect_tx(&b_full[s], 2u * kBBytes);
          tma_load_2d_cta2(smB + s * kBBytes + kh * kBBytesKH + half * kBoxBytes, &map_d, kh * 64,
                           tok0 + t * kTcTile + half * kTcBN + (int)rank * 64, mapa_u32(smem_u32(&b_full[s]), 0), pol);
        }
      }
      if (p.trace && lane == 0) {
        p.trace[blockIdx.x * 16 + 10] = w_be;
        p.trace[blockIdx.x * 16 + 11] = clock64() - T0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // The whole warp runs the loop with uniform control flow and lane 0 issues (predicated): the shared-memory
    // descriptors then live in uniform registers.  With `if (lane == 0)` around the loop every tcgen05.mma cost 10-12
    // instructions (64-bit adds + four R2UR moves) and the 163 instructions per 512-cycle MMA group made this ONE
    // thread, not the tensor cores, the pace of the kernel (profiles/r02_maxsim_trace_v6.txt: 654 cycles per group).
    if (rank == 0 && nseg > 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BF16, 256, kTcBN);
      int j = 0;
      int uses0 = 0, uses1 = 0;  // uses of the slots of pair tile 0 / 1 so far (phase of acc_empty / acc_full)
      const long long T0 = clock64();
      long long w_bf = 0, w_ae = 0, w_af = 0;
      // Descriptors are built once: the address field (bits [0,14) = byte address >> 4) of a tile at another offset is
      // the base descriptor plus offset >> 4 (all of shared memory is below 2^18 bytes).
      const uint64_t desc_a0 = umma_smem_desc_sw128(smem_u32(smA));
      const uint64_t desc_b0 = umma_smem_desc_sw128(smem_u32(smB));
      for (int si = 0; si < nseg; ++si) {
        const TcSeg sg = segs[si];
        const int n_act = min(kTcMG, p.num_pair_tiles - sg.g * kTcMG);
        const int tok0 = __ldg(p.doc_offsets + sg.d0);
        const int ntiles = (__ldg(p.doc_offsets + sg.d1) - tok0 + kTcTile - 1) / kTcTile;
        timed_wait(a_full, (uint32_t)si & 1u, w_af, tracing);
        for (int t = 0; t < ntiles; ++t, ++j) {
          const int s = j % kTcStages;
          const uint32_t ph = (uint32_t)(j / kTcStages) & 1u;
          timed_wait(&b_full[s], ph, w_bf, tracing);
          tc5_fence_after();
          const uint64_t desc_bs = desc_b0 + (uint64_t)((uint32_t)s * (kBBytes >> 4));
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int a = 0; a < kTcMG; ++a) {
              if (a < n_act) {
                constexpr int kDummy = 0;
                (void)kDummy;
                const int slot = a * 2 + half;
                timed_wait(&acc_empty[slot], ((uint32_t)(a == 0 ? uses0 : uses1) & 1u) ^ 1u, w_ae, tracing);
                tc5_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                  for (int kh = 0; kh < KH; ++kh) {
                    const uint64_t da = desc_a0 + (uint64_t)((a * kABytes + kh * kABytesKH) >> 4);
                    const uint64_t db = desc_bs + (uint64_t)((kh * kBBytesKH + half * kBoxBytes) >> 4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)  // 4 x UMMA_K(16 elements = 32 B) per 128-byte swizzle row
                      umma_f16_ss_cta2(tmem_base + (uint32_t)slot * kTcBN, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2),
                                       idesc, (kh | kk) != 0 ? 1u : 0u);
                  }
                  umma_commit_cta2(&acc_full[slot], 0b11);  // slot ready for its epilogue set in both CTAs
                  if (half == 1 && a == n_act - 1) umma_commit_cta2(&b_empty[s], 0b11);  // stage reusable (both CTAs)
                }
                __syncwarp();
              }
            }
          }
          ++uses0;
          uses1 += n_act > 1 ? 1 : 0;
        }
        if (elect_one_sync()) umma_commit_cta2(a_empty, 0b11);  // query tiles reusable once the segment's MMAs are done
        __syncwarp();
      }
      if (p.trace && lane == 0) {
        p.trace[blockIdx.x * 16 + 0] = clock64() - T0;
        p.trace[blockIdx.x * 16 + 1] = w_bf;
        p.trace[blockIdx.x * 16 + 2] = w_ae;
        p.trace[blockIdx.x * 16 + 3] = w_af;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue sets (set == pair tile)
    const int set = (warp - 4) >> 2;
    const int quarter = warp & 3;  // TMEM lanes 32*quarter .. +31
    const uint32_t lead_acc_empty = mapa_u32(smem_u32(&acc_empty[set * 2]), 0);  // + 8 bytes for the second half
    const int qpw = p.lq_pad >> 5;  // warps (lane quarters) per query: 1, 2 or 4
    int use = 0;                    // half-tiles drained so far: slot = set * 2 + (use & 1), phase = (use >> 1) & 1
    int nfin = 0;                   // documents finalized so far (buffer of the cross-warp sum)
    const long long T0 = clock64();
    long long w_full = 0;
    const int32_t* off = p.doc_offsets;
    for (int si = 0; si < nseg; ++si) {
      const TcSeg sg = segs[si];
      const int n_act = min(kTcMG, p.num_pair_tiles - sg.g * kTcMG);
      if (set >= n_act) continue;  // this set's pair tile does not exist in this group
      const int d0 = sg.d0, d1 = sg.d1;
      const int tok0 = __ldg(off + d0);
      const int ntiles = (__ldg(off + d1) - tok0 + kTcTile - 1) / kTcTile;
      const int mt = ((sg.g * kTcMG + set) << 1) + (int)rank;  // 128-row query tile of this CTA
      const int row = quarter * 32 + lane;                     // row in the 128-row tile
      const int query = mt * qpt + row / p.lq_pad;             // uniform across the warp (lq_pad % 32 == 0)
      const int tok = row % p.lq_pad;
      const bool q_valid = query < p.nq;
      float w = 0.f;
      if (q_valid && tok < p.lq)
        w = p.q_weight ? __ldg(p.q_weight + (size_t)query * p.lq + tok) : tc_reference_weight(tok, p.lq);
      // Ends (relative to tok0) of the documents doc .. doc + 31, one per lane; the batch after it is already in flight.
      auto load_ends = [&](int first) {
        const int dd = first + lane;
        return dd < d1 ? __ldg(off + dd + 1) - tok0 : INT_MAX;
      };
      int doc = d0, e_idx = 0;
      int e_cur = load_ends(d0), e_nxt = load_ends(d0 + 32);
      int e0 = __shfl_sync(0xFFFFFFFFu, e_cur, 0);  // end of the current document (INT_MAX past the segment)
      float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
      float* out_row = p.out + (size_t)(q_valid ? query : 0) * p.nd;

      auto finalize = [&]() {
        const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        const float part = warp_sum(w != 0.f ? w * m : 0.f);
        if (qpw == 1) {
          if (lane == 0 && q_valid) out_row[doc] = part;
        } else {
          // the qpw warps of a query add their parts in a fixed order (all four warps of the set see the same
          // document boundaries, so they meet here once per document)
          float* pp = parts + ((set * 2 + (nfin & 1)) << 2);
          if (lane == 0) pp[quarter] = part;
          named_bar_sync(kTcEpiBar + set, 128);
          if (lane == 0 && (quarter & (qpw - 1)) == 0 && q_valid) {
            float sum = pp[quarter];
            for (int i = 1; i < qpw; ++i) sum += pp[quarter + i];
            out_row[doc] = sum;
          }
          ++nfin;
        }
        m0 = m1 = m2 = m3 = -CUDART_INF_F;
        ++doc;
        if (++e_idx == 32) {
          e_cur = e_nxt;
          e_nxt = load_ends(doc + 32);
          e_idx = 0;
        }
        e0 = __shfl_sync(0xFFFFFFFFu, e_cur, e_idx);
      };
      // One 32-column chunk starting at column c0 of the segment.  Common case (no document ends inside it): 16
      // three-input max ops in 4 independent chains.  A document ending at column b of the chunk: switch on b's
      // 4-column group -> left / right maxima with static register indices (tc_split4), finalize, and the right
      // part opens the next document.  Further boundaries inside the same chunk (documents shorter than 32 tokens)
      // take a predicated loop.
      auto consume = [&](const uint32_t (&v)[32], int c0) {
        if (e0 > c0 + 32) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            m0 = fmaxf(fmaxf(m0, __uint_as_float(v[c + 0])), __uint_as_float(v[c + 1]));
            m1 = fmaxf(fmaxf(m1, __uint_as_float(v[c + 2])), __uint_as_float(v[c + 3]));
            m2 = fmaxf(fmaxf(m2, __uint_as_float(v[c + 4])), __uint_as_float(v[c + 5]));
            m3 = fmaxf(fmaxf(m3, __uint_as_float(v[c + 6])), __uint_as_float(v[c + 7]));
          }
          return;
        }
        const int b = min(max(e0 - c0, 0), 32);  // warp-uniform: the current document ends before column b of this chunk
        float l = -CUDART_INF_F, r = -CUDART_INF_F;
        switch (b >> 3) {
          case 0: tc_split8<0>(v, b, l, r); break;
          case 1: tc_split8<1>(v, b, l, r); break;
          case 2: tc_split8<2>(v, b, l, r); break;
          case 3: tc_split8<3>(v, b, l, r); break;
          default: tc_split8<3>(v, 32, l, r); break;  // b == 32: everything belongs to the ending document
        }
        m0 = fmaxf(m0, l);
        int done = b;       // columns [0, done) are accounted for
        bool rare = false;  // more than one document ends inside this chunk (documents shorter than 32 tokens)
        for (;;) {
          finalize();
          const bool more = e0 <= c0 + 32;
          if (!more && !rare) {
            m0 = r;  // the rest of the chunk opens the next document
            break;
          }
          const int b2 = more ? min(max(e0 - c0, done), 32) : 32;
          float seg = -CUDART_INF_F;
#pragma unroll
          for (int c = 0; c < 32; ++c) seg = fmaxf(seg, (c >= done && c < b2) ? __uint_as_float(v[c]) : -CUDART_INF_F);
          m0 = seg;
          done = b2;
          rare = true;
          if (!more) break;
        }
      };

      const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * 2) * kTcBN;
      uint32_t va[32], vb[32];
#pragma unroll 1
      for (int h = 0; h < 2 * ntiles; ++h, ++use) {  // half-tiles of 128 tokens, in token order
        const int half = use & 1;
        timed_wait(&acc_full[set * 2 + half], (uint32_t)(use >> 1) & 1u, w_full, tracing);
        tc5_fence_after();
        const uint32_t taddr = tlane + (uint32_t)half * kTcBN;
        const int cbase = h * kTcBN;
        tmem_ld_32x32(taddr, va);
        tmem_ld_wait(va);
#pragma unroll 1
        for (int cp = 0; cp < 2; ++cp) {  // two chunk pairs; one load is always in flight behind the reduction
          tmem_ld_32x32(taddr + cp * 64 + 32, vb);
          consume(va, cbase + cp * 64);
          tmem_ld_wait(vb);
          if (cp == 0) {
            tmem_ld_32x32(taddr + 64, va);
          } else {
            // the whole slot is in registers: hand it back to the MMA warp of the leader
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_acc_empty + (uint32_t)half * 8u);
          }
          consume(vb, cbase + cp * 64 + 32);
          if (cp == 0) tmem_ld_wait(va);
        }
      }
    }
    if (p.trace && quarter == 0 && lane == 0) {
      p.trace[blockIdx.x * 16 + 4 + set * 3] = clock64() - T0;
      p.trace[blockIdx.x * 16 + 5 + set * 3] = w_full;
    }
  }

  tc5_fence_before();
  cluster_sync_all();  // the peer may still read this CTA's tiles / signal its barriers
  if (warp == 2) {
    tc5_fence_after();
    tmem_dealloc_cta2(tmem_base, kTcTmemCols);
  }
}

// ================================================================================ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Tc5State {
  int device = 0;
  int num_sms = 0;
  EncodeTiledFn encode = nullptr;
  // scratch of the batched dense path (dense_tc5.cu)
  void* dense_ws = nullptr;
  size_t dense_ws_bytes = 0;
};

Tc5State* tc5_create(int device, int num_sms) {
  Tc5State* s = new Tc5State();
  s->device = device;
  s->num_sms = num_sms;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    s->encode = reinterpret_cast<EncodeTiledFn>(fn);
  else
    cudaGetLastError();
  return s;
}

void tc5_destroy(Tc5State* s) {
  if (!s) return;
  if (s->dense_ws) cudaFree(s->dense_ws);
  delete s;
}

void* tc5_dense_scratch(Tc5State* s, size_t bytes) {
  if (bytes > s->dense_ws_bytes) {
    if (s->dense_ws) cudaFree(s->dense_ws);
    s->dense_ws = nullptr;
    s->dense_ws_bytes = 0;
    if (cudaMalloc(&s->dense_ws, bytes) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    s->dense_ws_bytes = bytes;
  }
  return s->dense_ws;
}

bool tc5_encode(const Tc5State* s, CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, std::string* err) {
  if (!s->encode) {
    if (err) *err = "cuTensorMapEncodeTiled entry point not available";
    return false;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = s->encode(map, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                         (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return false;
  }
  return true;
}

// The corpus as [n_rows][row_bytes / 8] 8-byte elements, box = one whole row, no swizzle: the map the
// single-query scan's tile::gather4 copies read (dense_scan.cu), four arbitrary rows per instruction.
bool tc5_encode_rows(const Tc5State* s, CUtensorMap* map, const void* base, uint64_t n_rows, uint32_t row_bytes,
                     std::string* err) {
  if (!s || !s->encode) {
    if (err) *err = "cuTensorMapEncodeTiled entry point not available";
    return false;
  }
  if (row_bytes % 16 != 0 || row_bytes / 8 > 256 || n_rows == 0) {
    if (err) *err = "row tensor map: unsupported row size";
    return false;
  }
  cuuint64_t gdim[2] = {row_bytes / 8, n_rows};
  cuuint64_t gstr[1] = {row_bytes};
  cuuint32_t bdim[2] = {row_bytes / 8, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = s->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), gdim, gstr, bdim, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(rows) failed with CUresult " + std::to_string((int)r);
    return false;
  }
  return true;
}

int tc5_num_sms(const Tc5State* s) { return s->num_sms; }
bool tc5_has_encode(const Tc5State* s) { return s->encode != nullptr; }

bool tc5_maxsim_supported(const Tc5State* s, int nq, int lq, int d, int nd, const int32_t* cand,
                          const int32_t* out_argmax) {
  if (!s || !s->encode || s->num_sms < 2) return false;
  if (cand != nullptr || out_argmax != nullptr) return false;
  if (d != 64 && d != 128) return false;
  if (lq < 1 || lq > 128 || nd < 1) return false;
  const int lq_pad = lq <= 32 ? 32 : (lq <= 64 ? 64 : 128);
  return (long long)nq * lq_pad >= 128;  // at least one full 128-row tile of query tokens
}

int tc5_maxsim(Tc5State* s, const MaxSimParams& p, int dtype, cudaStream_t stream, int* launched, std::string* err) {
  *launched = 0;
  const int lq_pad = p.lq <= 32 ? 32 : (p.lq <= 64 ? 64 : 128);
  const int qpt = 128 / lq_pad;
  const int num_m_tiles = (p.nq + qpt - 1) / qpt;          // 128-row query tiles
  const int num_pair_tiles = (num_m_tiles + 1) / 2;         // 256-row tiles of a CTA pair
  const int mgroups = (num_pair_tiles + kTcMG - 1) / kTcMG;  // G
  // One pair per TPC (never more pairs than (group, document) units); documents in L2-sized phases.
  long long units = (long long)mgroups * p.nd;
  const int pairs = (int)(units < s->num_sms / 2 ? units : s->num_sms / 2);
  static const int phase_mb = getenv("RS_MAXSIM_PHASE_MB") ? atoi(getenv("RS_MAXSIM_PHASE_MB")) : 24;
  long long phases = ((long long)p.n_tokens * p.d * 2 + ((long long)phase_mb << 20) - 1) / ((long long)phase_mb << 20);
  if (phases < 1) phases = 1;
  if (phases > kTcMaxPhases) phases = kTcMaxPhases;
  while (phases > 1 && phases * ((mgroups + pairs - 1) / pairs + 2) > kTcMaxSegs) --phases;

  CUtensorMap map_q, map_d;
  {
    const uint64_t dims[3] = {(uint64_t)p.d, (uint64_t)p.lq, (uint64_t)p.nq};
    const uint64_t strides[2] = {(uint64_t)p.d * 2, (uint64_t)p.lq * p.d * 2};
    const uint32_t box[3] = {64, (uint32_t)lq_pad, (uint32_t)qpt};
    if (!tc5_encode(s, &map_q, dtype, 3, p.q, dims, strides, box, err)) return -2;
  }
  {
    const uint64_t dims[2] = {(uint64_t)p.d, (uint64_t)p.n_tokens};
    const uint64_t strides[1] = {(uint64_t)p.d * 2};
    const uint32_t box[2] = {64, 64};  // a CTA stages its half of a 256-token tile as two 64-token boxes
    if (!tc5_encode(s, &map_d, dtype, 2, p.doc_tokens, dims, strides, box, err)) return -2;
  }
  MaxSimTcParams kp{};
  kp.q_weight = p.q_weight;
  kp.doc_offsets = p.doc_offsets;
  kp.out = p.out_scores;
  kp.nq = p.nq;
  kp.lq = p.lq;
  kp.lq_pad = lq_pad;
  kp.nd = p.nd;
  kp.num_pair_tiles = num_pair_tiles;
  kp.num_mgroups = mgroups;
  kp.phases = (int)phases;
  // diagnostics: RS_MAXSIM_TRACE=1 prints, per launch, the cycles each role spent waiting (stderr; synchronises)
  static const bool trace_on = getenv("RS_MAXSIM_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  if (trace_on) {
    if (!trace_dev) cudaMalloc(&trace_dev, (size_t)s->num_sms * 16 * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, (size_t)s->num_sms * 16 * sizeof(long long), stream);
    kp.trace = trace_dev;
  }
  const int kh = p.d / 64;
  const size_t smem = 1024 + (size_t)kTcMG * 128 * 128 * kh + (size_t)kTcStages * (kTcTile / 2) * 128 * kh + 2048;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaSuccess;
#define RS_TC_LAUNCH(BF, KHV)                                                                                         \
  {                                                                                                                   \
    e = cudaFuncSetAttribute(maxsim_tc5_kernel<BF, KHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, maxsim_tc5_kernel<BF, KHV>, map_q, map_d, kp);                 \
  }
  if (dtype == 1) {
    if (kh == 1) RS_TC_LAUNCH(true, 1) else RS_TC_LAUNCH(true, 2)
  } else {
    if (kh == 1) RS_TC_LAUNCH(false, 1) else RS_TC_LAUNCH(false, 2)
  }
#undef RS_TC_LAUNCH
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return -3;
  }
  *launched = 1;
  if (trace_on) {
    std::vector<long long> t((size_t)2 * pairs * 16);
    cudaStreamSynchronize(stream);
    cudaMemcpy(t.data(), trace_dev, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    auto avg = [&](int slot, int rank_sel) {
      double sum = 0;
      int n = 0;
      for (int c = 0; c < 2 * pairs; ++c)
        if (rank_sel < 0 || (c & 1) == rank_sel) {
          sum += (double)t[(size_t)c * 16 + slot];
          ++n;
        }
      return n ? sum / n : 0.0;
    };
    fprintf(stderr,
            "[maxsim_tc5 trace] pairs %d groups %d phases %d | leader MMA: total %.0f wait b_full %.0f acc_empty %.0f a_full %.0f | "
            "epilogue set0 (leader/peer): total %.0f/%.0f wait acc_full %.0f/%.0f | set1: total %.0f/%.0f wait %.0f/%.0f | "
            "producer (leader/peer): total %.0f/%.0f wait b_empty %.0f/%.0f  [cycles, mean over CTAs]\n",
            pairs, mgroups, (int)phases, avg(0, 0), avg(1, 0), avg(2, 0), avg(3, 0), avg(4, 0), avg(4, 1), avg(5, 0), avg(5, 1),
            avg(7, 0), avg(7, 1), avg(8, 0), avg(8, 1), avg(11, 0), avg(11, 1), avg(10, 0), avg(10, 1));
  }
  return 0;
}

}  // namespace rs
