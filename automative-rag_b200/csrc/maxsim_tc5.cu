// maxsim_tc5.cu — shared-candidate ColBERT MaxSim on tcgen05 tensor cores with TMEM
// accumulators: the batched shape of ColBERTReranker.batch_rerank_queries
// (reference src/core/query/llm/rerankers.py:583-593 — documents encoded once, every query
// scored against the same list with _compute_maxsim_scores :215-265).
//
// The stage is one contraction S = Q_all [nq*lq, d] . D_all^T [d, n_tokens] followed by a
// segmented max over each document's tokens and a weighted sum over each query's tokens.
// Mapping onto the hardware:
//   * M (TMEM lanes) = query tokens, N (TMEM columns) = document tokens, 256 per tile; K = d (64 or
//     128) fits one shared-memory stage, so there is no K pipeline.  tcgen05.ld 32x32b hands thread t
//     of an epilogue warp ONE query-token row with 32 consecutive document tokens in registers, so
//     the max over document tokens is a register-only FMNMX3 chain and the sum over a query's 32
//     tokens is one warp shuffle reduction.  The token-score matrix S exists only in TMEM.
//   * CTA PAIRS (cta_group::2, a 2-wide cluster = the two SMs of a TPC): one tcgen05.mma of M = 256
//     per K step; each CTA stages its own 128 query rows and HALF of the 256 document tokens of a
//     tile (the tensor cores read the other half from the peer's shared memory), and holds its
//     128 x 256 accumulator in its own TMEM.  A pair keeps MG = 2 "pair tiles" (2 x 256 query rows)
//     resident, so every 64 KB document tile feeds 512 query rows and each SM pulls only 32 KB of
//     it: half the L2->SM bytes and half the shared-memory operand reads per flop of the
//     single-CTA version.  TMEM holds one accumulator per resident pair tile (2 x 256 columns);
//     epilogue warp set s drains accumulator s while the tensor cores fill the other one.
//   * warp roles (both CTAs): 0 = TMA producer (one lane per K half), 1 = MMA issuer (leader CTA
//     only, one elected thread), 2 = TMEM allocator, 4..7 = epilogue set 0, 8..11 = epilogue set 1
//     (warp % 4 selects the TMEM lane quarter).  TMA loads of both CTAs count on the LEADER's
//     barriers; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs; the
//     epilogues hand accumulators back with a (remote) arrive on the leader's barrier.
//   * work split: G query groups (512 query rows each) x R document ranges (equal token counts),
//     pair p = (group p % G, range p / G).  The G pairs of one range walk the same documents at the
//     same time, so a document tile comes from DRAM once and is served to the other groups out of
//     L2 (the previous split let the groups drift a whole corpus apart: 6.2x DRAM re-reads).  With
//     more groups than pairs, pair p runs groups p, p + P, ... over all documents.
//   * query-token sums that span several warps (lq > 32) are combined through shared memory in a
//     fixed order — results are run-to-run identical, like every other kernel of the library.
#include <cuda.h>
#include <math_constants.h>

#include <mutex>

#include "tc5.cuh"
#include "tc5_host.h"

namespace rs {

constexpr int kTcThreads = 384;
constexpr int kTcBN = 256;      // document tokens per tile (UMMA N); each CTA of the pair stages half
constexpr int kTcMG = 2;        // resident pair tiles (256 query rows each) == TMEM accumulators per CTA
constexpr int kTcStages = 4;    // B ring depth (each stage: 128 tokens x d per CTA)
constexpr int kTcTmemCols = 512;
constexpr int kTcEpiBar = 2;    // named barriers 2, 3: the four warps of epilogue set 0 / 1

struct MaxSimTcParams {
  const float* q_weight;
  const int32_t* doc_offsets;
  float* out;
  int32_t nq, lq, lq_pad, nd;
  int32_t num_pair_tiles;  // ceil(128-row query tiles / 2)
  int32_t num_mgroups;     // G: groups of kTcMG pair tiles
  int32_t ranges;          // R: document ranges per group (1 when G >= pairs)
};

__device__ __forceinline__ float tc_reference_weight(int i, int lq) {
  return (lq > 2 && (i == 0 || i == lq - 1)) ? 0.f : 1.f;  // rerankers.py:255-261
}

__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

template <bool BF16, int KH>
__global__ void __launch_bounds__(kTcThreads, 1)
    maxsim_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                      const MaxSimTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kABytesKH = 128 * 128;          // one K half (64 elements) of this CTA's 128 query rows
  constexpr uint32_t kBBytesKH = (kTcBN / 2) * 128;  // one K half of this CTA's 128 document tokens
  constexpr uint32_t kABytes = kABytesKH * KH;
  constexpr uint32_t kBBytes = kBBytesKH * KH;

  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), num_pairs = (int)(gridDim.x >> 1);
  const int G = p.num_mgroups, R = p.ranges;

  // ---- shared memory carve-up (1024-byte aligned: SWIZZLE_128B atoms)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* smA = sm;                                 // [kTcMG][KH][128 rows x 128 B]
  uint8_t* smB = smA + kTcMG * kABytes;              // [kTcStages][KH][128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + kTcStages * kBBytes);
  uint64_t* a_full = bars;                 // 1: query tiles of the current group have landed (leader's counts both CTAs)
  uint64_t* a_empty = bars + 1;            // 1: every MMA of the group has read them
  uint64_t* b_full = bars + 2;             // kTcStages (leader's)
  uint64_t* b_empty = b_full + kTcStages;  // kTcStages
  uint64_t* acc_full = b_empty + kTcStages;  // kTcMG
  uint64_t* acc_empty = acc_full + kTcMG;    // kTcMG (leader's collect both CTAs' epilogue warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + kTcMG);
  int* s_range = reinterpret_cast<int*>(tmem_ptr + 2);   // [2] first / one-past-last document of this pair's range
  float* parts = reinterpret_cast<float*>(s_range + 2);  // [2 sets][2 buffers][4 quarters] partial sums (lq > 32)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < kTcMG; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 2 * 4);  // the 4 warps of an epilogue set, in both CTAs
    }
    fence_mbar_init();
    // This pair's document range: range r of R, cut where the running token count passes r/R of the total, so
    // ragged documents give equal work.  Both CTAs compute the same bounds.
    int d0 = 0, d1 = p.nd;
    if (R > 1) {
      const int r = pair / G;
      const int64_t total = __ldg(p.doc_offsets + p.nd);
      auto cut = [&](int rr) -> int {  // first document whose first token is at or after rr/R of all tokens
        if (rr <= 0) return 0;
        if (rr >= R) return p.nd;
        const int64_t want = total * rr / R;
        int lo = 0, hi = p.nd;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((int64_t)__ldg(p.doc_offsets + mid) < want) lo = mid + 1; else hi = mid;
        }
        return lo;
      };
      d0 = cut(r);
      d1 = cut(r + 1);
    }
    s_range[0] = d0;
    s_range[1] = d1;
  }
  if (warp == 2) {
    tmem_alloc_cta2(tmem_ptr, kTcTmemCols);
    tmem_relinquish_cta2();
  }
  tc5_fence_before();
  cluster_sync_all();
  tc5_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr);
  const int d0 = s_range[0], d1 = s_range[1];

  const int qpt = 128 / p.lq_pad;  // queries per 128-row tile
  // groups of this pair: one (R > 1: group pair % G over its document range) or pair, pair + P, ... (all documents)
  const int g_first = R > 1 ? pair % G : pair;
  const int g_step = R > 1 ? G : num_pairs;
  const int tok0 = d0 < d1 ? __ldg(p.doc_offsets + d0) : 0;
  const int ntiles = d0 < d1 ? (__ldg(p.doc_offsets + d1) - tok0 + kTcBN - 1) / kTcBN : 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: lane kh loads K half kh
    if (lane < KH && ntiles > 0) {
      tma_prefetch_desc(lane == 0 ? &map_q : &map_d);
      const uint64_t pol = policy_evict_normal();
      const uint32_t lead_a_full = mapa_u32(smem_u32(a_full), 0);
      int j = 0;  // B tiles issued so far (ring position)
      int seg_i = 0;
      for (int g = g_first; g < G; g += g_step, ++seg_i) {
        const int n_act = min(kTcMG, p.num_pair_tiles - g * kTcMG);
        mbar_wait(a_empty, ((uint32_t)seg_i & 1u) ^ 1u);  // previous group's MMAs are done with A (both CTAs)
        if (rank == 0 && lane == 0) mbar_arrive_expect_tx(a_full, 2u * (uint32_t)n_act * kABytes);
        for (int a = 0; a < n_act; ++a)
          tma_load_3d_cta2(smA + a * kABytes + lane * kABytesKH, &map_q, lane * 64, 0,
                           ((g * kTcMG + a) * 2 + (int)rank) * qpt, lead_a_full, pol);
        for (int t = 0; t < ntiles; ++t, ++j) {
          const int s = j % kTcStages;
          const uint32_t ph = (uint32_t)(j / kTcStages) & 1u;
          mbar_wait(&b_empty[s], ph ^ 1u);
          if (rank == 0 && lane == 0) mbar_arrive_expect_tx(&b_full[s], 2u * kBBytes);
          tma_load_2d_cta2(smB + s * kBBytes + lane * kBBytesKH, &map_d, lane * 64,
                           tok0 + t * kTcBN + (int)rank * (kTcBN / 2), mapa_u32(smem_u32(&b_full[s]), 0), pol);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (lane == 0 && rank == 0 && ntiles > 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BF16, 256, kTcBN);
      int j = 0;
      int uses[kTcMG] = {0, 0};  // accumulator uses so far (phase of acc_empty / acc_full)
      int seg_i = 0;
      for (int g = g_first; g < G; g += g_step, ++seg_i) {
        const int n_act = min(kTcMG, p.num_pair_tiles - g * kTcMG);
        mbar_wait(a_full, (uint32_t)seg_i & 1u);
        for (int t = 0; t < ntiles; ++t, ++j) {
          const int s = j % kTcStages;
          const uint32_t ph = (uint32_t)(j / kTcStages) & 1u;
          mbar_wait(&b_full[s], ph);
          tc5_fence_after();
          for (int a = 0; a < n_act; ++a) {
            mbar_wait(&acc_empty[a], ((uint32_t)uses[a] & 1u) ^ 1u);
            ++uses[a];
            tc5_fence_after();
#pragma unroll
            for (int kh = 0; kh < KH; ++kh) {
              const uint64_t da = umma_smem_desc_sw128(smem_u32(smA + a * kABytes + kh * kABytesKH));
              const uint64_t db = umma_smem_desc_sw128(smem_u32(smB + s * kBBytes + kh * kBBytesKH));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)  // 4 x UMMA_K(16 elements = 32 B) per 128-byte swizzle row
                umma_f16_ss_cta2(tmem_base + (uint32_t)a * kTcBN, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc,
                                 (kh | kk) != 0 ? 1u : 0u);
            }
            umma_commit_cta2(&acc_full[a], 0b11);  // accumulator a ready for its epilogue set in both CTAs
          }
          umma_commit_cta2(&b_empty[s], 0b11);  // B stage reusable (both CTAs) once the MMAs have read it
        }
        umma_commit_cta2(a_empty, 0b11);  // query tiles reusable once every MMA of the group has completed
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue sets
    const int set = (warp - 4) >> 2;
    const int quarter = warp & 3;  // TMEM lanes 32*quarter .. +31
    const uint32_t lead_acc_empty = mapa_u32(smem_u32(&acc_empty[set]), 0);
    const int qpw = p.lq_pad >> 5;  // warps (lane quarters) per query: 1, 2 or 4
    int use = 0;                    // uses of this set's accumulator so far (phase of acc_full)
    int fin = 0;                    // documents finalized so far (buffer of the cross-warp sum)
    for (int g = g_first; g < G && ntiles > 0; g += g_step) {
      const int n_act = min(kTcMG, p.num_pair_tiles - g * kTcMG);
      if (set >= n_act) continue;  // this set's pair tile does not exist in this group
      const int mt = ((g * kTcMG + set) << 1) + (int)rank;  // 128-row query tile of this CTA
      const int row = quarter * 32 + lane;                  // row in the 128-row tile
      const int query = mt * qpt + row / p.lq_pad;          // uniform across the warp (lq_pad % 32 == 0)
      const int tok = row % p.lq_pad;
      const bool q_valid = query < p.nq;
      float w = 0.f;
      if (q_valid && tok < p.lq)
        w = p.q_weight ? __ldg(p.q_weight + (size_t)query * p.lq + tok) : tc_reference_weight(tok, p.lq);
      const int32_t* off = p.doc_offsets;
      int doc = d0;
      int e0 = __ldg(off + d0 + 1) - tok0;
      int e1 = (d0 + 2 <= d1) ? __ldg(off + d0 + 2) - tok0 : INT_MAX;
      int e2 = (d0 + 3 <= d1) ? __ldg(off + d0 + 3) - tok0 : INT_MAX;
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
          float* pp = parts + ((set * 2 + (fin & 1)) << 2);
          if (lane == 0) pp[quarter] = part;
          named_bar_sync(kTcEpiBar + set, 128);
          if (lane == 0 && (quarter & (qpw - 1)) == 0 && q_valid) {
            float sum = pp[quarter];
            for (int i = 1; i < qpw; ++i) sum += pp[quarter + i];
            out_row[doc] = sum;
          }
          ++fin;
        }
        m0 = m1 = m2 = m3 = -CUDART_INF_F;
        ++doc;
        e0 = e1;
        e1 = e2;
        e2 = (doc + 3 <= d1) ? __ldg(off + doc + 3) - tok0 : INT_MAX;
        if (doc >= d1) e0 = INT_MAX;
      };
      // One 32-column chunk.  Common case (no document ends inside it): 16 three-input max ops in
      // 4 independent chains.  Otherwise each segment [a, b) of the chunk is reduced with a
      // predicated max (static register indexing, no per-column branches) and the document is
      // finalized ONCE per boundary: the cold path stays small, which matters because an unrolled
      // per-column version (32 inlined finalize bodies per chunk) made the kernel 24k instructions
      // and the profile was dominated by instruction-cache misses (stall_no_inst).
      auto consume = [&](const uint32_t (&v)[32], int c0) {
        int a = 0;
        while (e0 <= c0 + 32) {  // warp-uniform
          const int b = e0 - c0;  // 0..32: the current document ends before column b of this chunk
          float seg = -CUDART_INF_F;
#pragma unroll
          for (int c = 0; c < 32; ++c) seg = fmaxf(seg, (c >= a && c < b) ? __uint_as_float(v[c]) : -CUDART_INF_F);
          m0 = fmaxf(m0, seg);
          finalize();
          a = b;
        }
        if (a == 0) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            m0 = fmaxf(fmaxf(m0, __uint_as_float(v[c + 0])), __uint_as_float(v[c + 1]));
            m1 = fmaxf(fmaxf(m1, __uint_as_float(v[c + 2])), __uint_as_float(v[c + 3]));
            m2 = fmaxf(fmaxf(m2, __uint_as_float(v[c + 4])), __uint_as_float(v[c + 5]));
            m3 = fmaxf(fmaxf(m3, __uint_as_float(v[c + 6])), __uint_as_float(v[c + 7]));
          }
        } else if (a < 32) {
          float seg = -CUDART_INF_F;
#pragma unroll
          for (int c = 0; c < 32; ++c) seg = fmaxf(seg, (c >= a) ? __uint_as_float(v[c]) : -CUDART_INF_F);
          m0 = fmaxf(m0, seg);
        }
      };

      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)set * kTcBN;
      uint32_t va[32], vb[32];
      for (int j = 0; j < ntiles; ++j, ++use) {
        mbar_wait(&acc_full[set], (uint32_t)use & 1u);
        tc5_fence_after();
        const int cbase = j * kTcBN;
        tmem_ld_32x32(taddr, va);
        tmem_ld_wait(va);
#pragma unroll 1
        for (int ch = 0; ch < kTcBN / 32; ch += 2) {
          tmem_ld_32x32(taddr + (ch + 1) * 32, vb);  // in flight while chunk ch is reduced
          consume(va, cbase + ch * 32);
          tmem_ld_wait(vb);
          if (ch + 2 < kTcBN / 32) {
            tmem_ld_32x32(taddr + (ch + 2) * 32, va);
          } else {
            // all of this accumulator is in registers: hand the TMEM slot back to the MMA warp of the leader
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_acc_empty);
          }
          consume(vb, cbase + (ch + 1) * 32);
          if (ch + 2 < kTcBN / 32) tmem_ld_wait(va);
        }
      }
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
  // G groups x R document ranges of pairs; the G pairs of a range stream the same documents side by side.
  const int pairs_avail = s->num_sms / 2;
  int ranges = 1, pairs = pairs_avail;
  if (mgroups < pairs_avail) {
    ranges = pairs_avail / mgroups;
    if (ranges > p.nd) ranges = p.nd;
    if (ranges < 1) ranges = 1;
    pairs = mgroups * ranges;
  }

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
    const uint32_t box[2] = {64, (uint32_t)(kTcBN / 2)};  // each CTA of a pair stages half of a tile
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
  kp.ranges = ranges;
  const int kh = p.d / 64;
  const size_t smem = 1024 + (size_t)kTcMG * 128 * 128 * kh + (size_t)kTcStages * (kTcBN / 2) * 128 * kh + 512;
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
  return 0;
}

}  // namespace rs
