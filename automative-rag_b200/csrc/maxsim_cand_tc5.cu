// maxsim_cand_tc5.cu — per-query-candidate ColBERT MaxSim on tcgen05 tensor cores: the
// retrieve-then-rerank shape, where every query scores ITS OWN candidate documents.
//
// Reference: ColBERTReranker.rerank -> _colbert_rerank -> _compute_maxsim_scores
// (src/core/query/llm/rerankers.py:351-385, :215-265): one query [Lq, D] against the token
// embeddings of the documents retrieved for that query; per document S = Q . D^T (:247), max over
// document tokens (:250), weighted sum over query tokens (:255-261).  The arg-max token per query
// token and the per-token maxima serve _explain_colbert_matches (:489-501).  D is the BERT hidden
// size as deployed (768, :118-120,159) or a projected 64 / 128.
//
// Unlike the shared-candidate kernel (maxsim_tc5.cu) no document tile is reused by a second
// query, so the stage is HBM-bound: 2 * D bytes per (query, candidate) token at 32 flops per
// byte.  The kernel is therefore organised around the document stream:
//   * a work item is a (query, candidate) PAIR; the flattened pair list is cut into equal
//     contiguous spans, one per SM.  Pairs are query-major, so a CTA sees few query changes.
//     Pairs whose candidate is not a document of this collection (index < 0 or >= nd: padding, or a
//     candidate owned by another GPU in the sharded stage) and empty documents never enter the
//     pipeline: the producer writes their -inf score itself.
//   * a document is cut into CHUNKS of <= 128 tokens (equal parts, rounded up to 32); a chunk is
//     streamed K-block by K-block: a ring STAGE holds KPS (1 or 2) blocks of 64 elements x <= 128
//     tokens, fetched by TMA as 32-row boxes straight from the packed token buffer (lanes of the
//     producer warp issue the boxes of a stage in parallel); rows past the document's end belong to
//     the next document (or are zero-filled past the buffer) and are masked in the epilogue.
//   * one tcgen05.mma group per chunk: D[128 x n] += Q[128 x 64] . block[n x 64]^T over the d / 64
//     K blocks.  Only the query's lq_pad rows are staged (A block kb at A + kb * lq_pad * 128 bytes);
//     the tensor core reads 128 rows from there, so accumulator rows >= lq_pad hold garbage that no
//     epilogue warp ever loads.  Four 128-column TMEM accumulators decouple the tensor pipe from
//     the epilogue.  The tensor pipe idles most of the time — it only has to keep up with HBM.
//   * epilogue warp q (TMEM lane quarter q, active when lq > 32 q) owns query tokens 32q..32q+31:
//     tcgen05.ld gives each thread its token's scores against 32 document tokens; running max
//     (and, for the explanations path, the first arg-max) over the document's chunks, then w . max
//     summed over the warp (and over the active warps in a fixed order, so results are run-to-run
//     identical).
//   * the producer describes every chunk in a small shared-memory ring (tokens, valid tokens,
//     first/last chunk of the document, query switch, output index), so the MMA and epilogue warps
//     never touch cand[] / doc_offsets[] themselves.
#include <cuda.h>
#include <math_constants.h>

#include "tc5.cuh"
#include "tc5_host.h"

namespace rs {

constexpr int kCdThreads = 256;  // warps: 0 producer, 1 MMA issuer, 2 TMEM allocator, 4..7 epilogue
constexpr int kCdBN = 128;       // most document tokens per chunk (UMMA N)
constexpr int kCdBox = 32;       // rows per TMA box
constexpr int kCdMaxStages = 5;  // ring depth (fewer when a wide query tile leaves less shared memory)
constexpr int kCdSlots = 4;      // TMEM accumulators (128 columns each)
constexpr int kCdDescRing = 16;  // > slots + 2: a descriptor is dead before its entry is reused
constexpr int kCdTmemCols = 512;
constexpr int kCdEpiBar = 2;     // named barrier of the active epilogue warps
constexpr int kCdABudget = 96 * 1024;  // both query buffers: 2 * lq_pad * d * 2 bytes

enum : int { kCdFirst = 1, kCdLast = 2, kCdNewQuery = 4, kCdEnd = 8 };

struct CandParams {
  const float* q_weight;       // [nq, lq] or null
  const int32_t* doc_offsets;  // [nd + 1]
  const int32_t* cand;         // [nq, nc] or null (candidate j of every query is document j)
  float* out;                  // [nq, nc]
  int32_t* out_argmax;         // null or [nq, nc, lq]
  float* out_tokmax;           // null or [nq, nc, lq]
  int64_t pairs;               // nq * nc
  int32_t nq, lq, lq_pad, nd, nc;
  int32_t kblocks;             // d / 64
  int32_t stages;              // ring depth in use
  int32_t a_rows;              // rows of a query K block in shared memory: 128 (zero-padded) when both buffers fit, else lq_pad
};

struct ChunkDesc {
  int32_t n;        // tokens fed to the MMA (multiple of 32)
  int32_t valid;    // tokens that belong to the document
  int32_t flags;
  int32_t query;
  int32_t tok_off;  // index of the chunk's first token inside its document
  int32_t pad;
  int64_t pair;     // output index
};

__device__ __forceinline__ float cd_reference_weight(int i, int lq) {
  return (lq > 2 && (i == 0 || i == lq - 1)) ? 0.f : 1.f;  // rerankers.py:255-261
}

__device__ __forceinline__ uint64_t cd_policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

template <bool BF16, int KPS, bool ARGMAX>
__global__ void __launch_bounds__(kCdThreads, 1)
    maxsim_cand_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                           const CandParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kBlkBytes = 128 * 128;             // one K block (64 elements) of a 128-token chunk
  constexpr uint32_t kStageBytes = kBlkBytes * KPS;
  constexpr uint32_t kBoxBytes = kCdBox * 128;

  const int64_t p_begin = p.pairs * blockIdx.x / gridDim.x;
  const int64_t p_end = p.pairs * (blockIdx.x + 1) / gridDim.x;
  if (p_begin >= p_end) return;  // uniform: nothing allocated yet

  // ---- shared memory carve-up (1024-byte aligned: SWIZZLE_128B atoms)
  const int KB = p.kblocks, S = p.stages;
  const uint32_t a_blk = (uint32_t)p.a_rows * 128u;  // bytes of one K block of the query tile
  const uint32_t a_bytes = a_blk * (uint32_t)KB;     // one query buffer
  const uint32_t a_tx = (uint32_t)p.lq_pad * 128u * (uint32_t)KB;  // bytes TMA delivers per query (lq_pad rows per block)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* smA = sm;                        // [2][KB][a_rows x 128 B]
  uint8_t* smB = smA + 2 * a_bytes;         // [S][KPS][128 rows x 128 B]  (with a_rows < 128 the tensor core's 128-row
                                            //  window of the last query block ends inside this ring: garbage rows)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)S * kStageBytes);
  uint64_t* a_full = bars;                       // 2
  uint64_t* a_empty = a_full + 2;                // 2
  uint64_t* b_full = a_empty + 2;                // kCdMaxStages
  uint64_t* b_empty = b_full + kCdMaxStages;     // kCdMaxStages
  uint64_t* acc_full = b_empty + kCdMaxStages;   // kCdSlots
  uint64_t* acc_empty = acc_full + kCdSlots;     // kCdSlots
  ChunkDesc* desc = reinterpret_cast<ChunkDesc*>(acc_empty + kCdSlots);  // [kCdDescRing], written by the producer
  ChunkDesc* acc_desc = desc + kCdDescRing;                                // [kCdSlots], written by the MMA warp
  float* parts = reinterpret_cast<float*>(acc_desc + kCdSlots);            // [2][4] per-quarter partial sums
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(parts + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nquad = p.lq_pad / 32;  // active epilogue warps

  if (warp == 0 && lane == 0) {
    for (int a = 0; a < 2; ++a) {
      mbar_init(&a_full[a], 1);
      mbar_init(&a_empty[a], 1);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < kCdSlots; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], nquad);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, kCdTmemCols);
    tmem_relinquish();
  }
  if (p.a_rows == 128) {
    // Narrow rows (d <= 192): full 128-row query tiles, rows >= lq_pad zeroed once and never touched again (TMA refreshes
    // the first lq_pad rows).  Measured on config 4b, same box: 3.10 ms with zero padding, 3.37 ms when the tensor core's
    // window ran over live ring data instead (profiles/r02_maxsim_cand_ab.txt).
    for (uint32_t i = threadIdx.x * 16; i < 2 * a_bytes; i += kCdThreads * 16)
      *reinterpret_cast<uint4*>(smA + i) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to TMA / UMMA
  }
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp)
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_d);
    }
    const uint64_t pol = cd_policy_evict_normal();
    uint32_t seq = 0;   // chunks issued so far (only used modulo powers of two)
    int rs = 0;         // ring cursor: next stage and the parity of its current use — kept incrementally, a 64-bit
    uint32_t rph = 0;   // modulo by the run-time ring depth per stage cost the producer a third of its time
    int a_seq = -1;     // query switches so far - 1
    int cur_query = -1;
    // document bounds of pair (base + lane), fetched one batch of 32 pairs ahead of their use; a pair without tokens
    // (candidate outside the collection, empty document) is scored right here and never enters the pipeline
    auto fetch = [&](int64_t pair, int& o0, int& o1) {
      o0 = o1 = 0;
      if (pair < p_end) {
        const int j = (int)(pair % p.nc);
        const int doc = p.cand ? __ldg(p.cand + pair) : j;
        if (doc >= 0 && doc < p.nd) {
          o0 = __ldg(p.doc_offsets + doc);
          o1 = __ldg(p.doc_offsets + doc + 1);
        }
        if (o1 <= o0) {
          p.out[pair] = -CUDART_INF_F;
          if (ARGMAX) {
            for (int t = 0; t < p.lq; ++t) {
              if (p.out_argmax) p.out_argmax[pair * p.lq + t] = 0;
              if (p.out_tokmax) p.out_tokmax[pair * p.lq + t] = -CUDART_INF_F;
            }
          }
        }
      }
    };
    int o0n, o1n;
    fetch(p_begin + lane, o0n, o1n);
    for (int64_t pb = p_begin; pb < p_end; pb += 32) {
      const int o0l = o0n, o1l = o1n;
      fetch(pb + 32 + lane, o0n, o1n);
      uint32_t live = __ballot_sync(0xFFFFFFFFu, o1l > o0l);  // pairs of this batch that have tokens
      while (live) {
        const int i = __ffs(live) - 1;
        live &= live - 1;
        const int64_t pair = pb + i;
        const int o0 = __shfl_sync(0xFFFFFFFFu, o0l, i), o1 = __shfl_sync(0xFFFFFFFFu, o1l, i);
        const int query = (int)(pair / p.nc);
        const int ld = o1 - o0;
        const int nch = (ld + kCdBN - 1) / kCdBN;
        const int per = (((ld + nch - 1) / nch) + kCdBox - 1) / kCdBox * kCdBox;  // chunk length, multiple of 32
        for (int c = 0; c < nch; ++c, ++seq) {
          const int valid = max(0, min(per, ld - c * per));
          const int n = (valid + kCdBox - 1) / kCdBox * kCdBox;
          int flags = (c == 0 ? kCdFirst : 0) | (c == nch - 1 ? kCdLast : 0);
          if (query != cur_query) {  // warp-uniform: stage the new query's tokens in the other A buffer
            flags |= kCdNewQuery;
            cur_query = query;
            ++a_seq;
            const int ab = a_seq & 1;
            if (lane == 0) {
              mbar_wait(&a_empty[ab], ((uint32_t)(a_seq >> 1) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&a_full[ab], a_tx);
            }
            __syncwarp();
            for (int kb = lane; kb < KB; kb += 32)
              tma_load_3d(smA + ab * a_bytes + kb * a_blk, &map_q, kb * 64, 0, query, &a_full[ab], pol);
          }
          if (lane == 0) {
            ChunkDesc& dd = desc[seq % kCdDescRing];
            dd.n = n;
            dd.valid = valid;
            dd.flags = flags;
            dd.query = query;
            dd.tok_off = c * per;
            dd.pair = pair;
          }
          const int nbox = n / kCdBox;
          for (int st = 0; st < KB / KPS; ++st) {
            const int s = rs;
            mbar_wait(&b_empty[s], rph ^ 1u);
            if (lane == 0) mbar_arrive_expect_tx(&b_full[s], (uint32_t)n * 128u * KPS);
            __syncwarp();
            if (lane < nbox * KPS) {
              const int b = lane / KPS, k2 = lane % KPS;
              tma_load_2d(smB + (size_t)s * kStageBytes + k2 * kBlkBytes + b * kBoxBytes, &map_d, (st * KPS + k2) * 64,
                          o0 + c * per + b * kCdBox, &b_full[s], pol);
            }
            if (++rs == S) {
              rs = 0;
              rph ^= 1u;
            }
          }
        }
      }
    }
    {  // END marker: a descriptor and one (empty) ring stage
      const int s = rs;
      mbar_wait(&b_empty[s], rph ^ 1u);
      if (lane == 0) {
        ChunkDesc& dd = desc[seq % kCdDescRing];
        dd.n = 0;
        dd.valid = 0;
        dd.flags = kCdEnd;
        dd.query = -1;
        dd.tok_off = 0;
        dd.pair = -1;
        mbar_arrive(&b_full[s]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int a_seq = -1;
      int rs = 0;        // ring cursor (see the producer)
      uint32_t rph = 0;
      for (uint32_t seq = 0;; ++seq) {
        // the first stage of the chunk (or the END marker's stage) also publishes the chunk's descriptor
        int s = rs;
        mbar_wait(&b_full[s], rph);
        const ChunkDesc dd = desc[seq % kCdDescRing];
        const int slot = (int)(seq % kCdSlots);
        mbar_wait(&acc_empty[slot], ((seq / kCdSlots) & 1u) ^ 1u);
        tc5_fence_after();
        if (dd.flags & kCdEnd) {
          acc_desc[slot] = dd;
          __threadfence_block();
          umma_commit(&acc_full[slot]);
          break;
        }
        if (dd.flags & kCdNewQuery) {
          if (a_seq >= 0) umma_commit(&a_empty[a_seq & 1]);  // free once every MMA of the previous query is done
          ++a_seq;
          mbar_wait(&a_full[a_seq & 1], (uint32_t)(a_seq >> 1) & 1u);
          tc5_fence_after();
        }
        const uint32_t idesc = (1u << 4) | ((BF16 ? 1u : 0u) << 7) | ((BF16 ? 1u : 0u) << 10) |
                               ((uint32_t)(dd.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint8_t* At = smA + (a_seq & 1) * a_bytes;
        for (int st = 0; st < KB / KPS; ++st) {
          if (st > 0) {
            s = rs;
            mbar_wait(&b_full[s], rph);
            tc5_fence_after();
          }
#pragma unroll
          for (int k2 = 0; k2 < KPS; ++k2) {
            const int kb = st * KPS + k2;
            const uint64_t da = umma_smem_desc_sw128(smem_u32(At + kb * a_blk));
            const uint64_t db = umma_smem_desc_sw128(smem_u32(smB + (size_t)s * kStageBytes + k2 * kBlkBytes));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)  // 4 x UMMA_K(16 elements = 32 B) per 128-byte swizzle row
              umma_f16_ss(tmem_base + (uint32_t)slot * kCdBN, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc,
                          (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&b_empty[s]);
          if (++rs == S) {
            rs = 0;
            rph ^= 1u;
          }
        }
        acc_desc[slot] = dd;
        __threadfence_block();
        umma_commit(&acc_full[slot]);  // fires once the MMAs above have completed
      }
    }
  } else if (warp >= 4 && warp - 4 < nquad) {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp - 4;  // == warp % 4: TMEM lanes 32*quarter .. +31
    const int tok = quarter * 32 + lane;
    float w = 0.f;
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F;
    int am = 0;  // ARGMAX: document token of the running maximum m0 (the first one on ties)
    int docs_done = 0;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // max over the first `cnt` (1..32) of 32 columns; `t0` = document-relative index of column 0
    auto consume = [&](const uint32_t (&v)[32], int cnt, int t0) {
      if (ARGMAX) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float x = __uint_as_float(v[c]);
          if (c < cnt && x > m0) {  // strict: keeps the first maximal token, as the reference's argmax over a row
            m0 = x;
            am = t0 + c;
          }
        }
      } else if (cnt >= 32) {
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          m0 = fmaxf(fmaxf(m0, __uint_as_float(v[c + 0])), __uint_as_float(v[c + 1]));
          m1 = fmaxf(fmaxf(m1, __uint_as_float(v[c + 2])), __uint_as_float(v[c + 3]));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) m0 = fmaxf(m0, c < cnt ? __uint_as_float(v[c]) : -CUDART_INF_F);
      }
    };
    uint32_t va[32], vb[32];
    for (uint32_t seq = 0;; ++seq) {
      const int slot = (int)(seq % kCdSlots);
      mbar_wait(&acc_full[slot], (seq / kCdSlots) & 1u);
      tc5_fence_after();
      const ChunkDesc dd = acc_desc[slot];
      if (dd.flags & kCdEnd) break;
      if (dd.flags & kCdNewQuery) {
        w = 0.f;
        if (tok < p.lq) w = p.q_weight ? __ldg(p.q_weight + (size_t)dd.query * p.lq + tok) : cd_reference_weight(tok, p.lq);
      }
      if (dd.flags & kCdFirst) {
        m0 = m1 = -CUDART_INF_F;
        am = 0;
      }
      const uint32_t taddr = tlane + (uint32_t)slot * kCdBN;
      const int nblk = dd.n >> 5;  // 1..4, warp-uniform
      tmem_ld_32x32(taddr, va);
      tmem_ld_wait(va);
      if (nblk > 1) tmem_ld_32x32(taddr + 32, vb);
      consume(va, dd.valid, dd.tok_off);
      if (nblk > 1) {
        tmem_ld_wait(vb);
        if (nblk > 2) tmem_ld_32x32(taddr + 64, va);
        consume(vb, dd.valid - 32, dd.tok_off + 32);
        if (nblk > 2) {
          tmem_ld_wait(va);
          if (nblk > 3) tmem_ld_32x32(taddr + 96, vb);
          consume(va, dd.valid - 64, dd.tok_off + 64);
          if (nblk > 3) {
            tmem_ld_wait(vb);
            consume(vb, dd.valid - 96, dd.tok_off + 96);
          }
        }
      }
      // the accumulator is in registers: hand the TMEM slot back
      tc5_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
      if (dd.flags & kCdLast) {
        const float m = fmaxf(m0, m1);
        if (ARGMAX && tok < p.lq) {
          if (p.out_argmax) p.out_argmax[dd.pair * p.lq + tok] = am;
          if (p.out_tokmax) p.out_tokmax[dd.pair * p.lq + tok] = m;
        }
        const float part = warp_sum(w != 0.f ? w * m : 0.f);
        if (nquad == 1) {
          if (lane == 0) p.out[dd.pair] = part;
        } else {
          float* pp = parts + (docs_done & 1) * 4;
          if (lane == 0) pp[quarter] = part;
          named_bar_sync(kCdEpiBar, nquad * 32);
          if (quarter == 0 && lane == 0) {
            float sum = pp[0];
            for (int qd = 1; qd < nquad; ++qd) sum += pp[qd];
            p.out[dd.pair] = sum;
          }
          ++docs_done;
        }
      }
    }
  }

  tc5_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc5_fence_after();
    tmem_dealloc(tmem_base, kCdTmemCols);
  }
}

// ================================================================================ host side
static int cand_lq_pad(int lq) { return lq <= 32 ? 32 : (lq <= 64 ? 64 : (lq <= 96 ? 96 : 128)); }

// rows of a query K block in shared memory: the full zero-padded 128 when both buffers fit the budget, else lq_pad
static int cand_a_rows(int lq_pad, int d) { return (size_t)2 * 128 * d * 2 <= (size_t)kCdABudget ? 128 : lq_pad; }

// ring depth that fits next to the two query buffers, or 0 when the shape does not fit at all
static int cand_stages(int lq_pad, int d, int kps) {
  const size_t a_bytes = (size_t)2 * cand_a_rows(lq_pad, d) * d * 2;
  if (a_bytes > (size_t)kCdABudget) return 0;
  const size_t budget = 225 * 1024 - 1024 /*alignment*/ - 2048 /*barriers, descriptors*/ - a_bytes;
  int s = (int)(budget / ((size_t)128 * 128 * kps));
  return s > kCdMaxStages ? kCdMaxStages : s;
}

bool tc5_maxsim_cand_supported(const Tc5State* s, int nq, int lq, int d, int nd, int nc) {
  if (!s || !tc5_has_encode(s)) return false;
  if (d < 64 || d > 1024 || (d % 64) != 0) return false;
  if (!(lq >= 1 && lq <= 128 && nd >= 1 && nc >= 1 && nq >= 1)) return false;
  const int kps = (d % 128) == 0 ? 2 : 1;
  return cand_stages(cand_lq_pad(lq), d, kps) >= 2;
}

int tc5_maxsim_cand(Tc5State* s, const MaxSimParams& p, int dtype, cudaStream_t stream, int* launched, std::string* err) {
  *launched = 0;
  const int lq_pad = cand_lq_pad(p.lq);
  const int nc = p.cand ? p.nc : p.nd;
  const long long pairs = (long long)p.nq * nc;
  const int num_sms = tc5_num_sms(s);
  const int grid_x = (int)(pairs < num_sms ? pairs : num_sms);
  const int kps = (p.d % 128) == 0 ? 2 : 1;  // K blocks per ring stage
  const int stages = cand_stages(lq_pad, p.d, kps);
  if (stages < 2) {
    *err = "query tile does not fit shared memory";
    return -2;
  }

  CUtensorMap map_q, map_d;
  {
    const uint64_t dims[3] = {(uint64_t)p.d, (uint64_t)p.lq, (uint64_t)p.nq};
    const uint64_t strides[2] = {(uint64_t)p.d * 2, (uint64_t)p.lq * p.d * 2};
    const uint32_t box[3] = {64, (uint32_t)lq_pad, 1};
    if (!tc5_encode(s, &map_q, dtype, 3, p.q, dims, strides, box, err)) return -2;
  }
  {
    const uint64_t dims[2] = {(uint64_t)p.d, (uint64_t)p.n_tokens};
    const uint64_t strides[1] = {(uint64_t)p.d * 2};
    const uint32_t box[2] = {64, (uint32_t)kCdBox};
    if (!tc5_encode(s, &map_d, dtype, 2, p.doc_tokens, dims, strides, box, err)) return -2;
  }
  CandParams kp{};
  kp.q_weight = p.q_weight;
  kp.doc_offsets = p.doc_offsets;
  kp.cand = p.cand;
  kp.out = p.out_scores;
  kp.out_argmax = p.out_argmax;
  kp.out_tokmax = p.out_tokmax;
  kp.pairs = pairs;
  kp.nq = p.nq;
  kp.lq = p.lq;
  kp.lq_pad = lq_pad;
  kp.nd = p.nd;
  kp.nc = nc;
  kp.kblocks = p.d / 64;
  kp.stages = stages;
  kp.a_rows = cand_a_rows(lq_pad, p.d);
  const bool argmax = p.out_argmax != nullptr || p.out_tokmax != nullptr;
  const size_t smem = 1024 + (size_t)2 * kp.a_rows * p.d * 2 + (size_t)stages * 128 * 128 * kps + 2048;
  dim3 grid(grid_x);
  cudaError_t e = cudaSuccess;
#define RS_CD_LAUNCH(BF, KPSV, AM)                                                                                          \
  {                                                                                                                         \
    e = cudaFuncSetAttribute(maxsim_cand_tc5_kernel<BF, KPSV, AM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) {                                                                                                 \
      maxsim_cand_tc5_kernel<BF, KPSV, AM><<<grid, kCdThreads, smem, stream>>>(map_q, map_d, kp);                          \
      e = cudaGetLastError();                                                                                               \
    }                                                                                                                       \
  }
#define RS_CD_DISPATCH(BF)                                            \
  {                                                                   \
    if (argmax) {                                                     \
      if (kps == 1) RS_CD_LAUNCH(BF, 1, true) else RS_CD_LAUNCH(BF, 2, true)   \
    } else {                                                          \
      if (kps == 1) RS_CD_LAUNCH(BF, 1, false) else RS_CD_LAUNCH(BF, 2, false) \
    }                                                                 \
  }
  if (dtype == 1) RS_CD_DISPATCH(true) else RS_CD_DISPATCH(false)
#undef RS_CD_DISPATCH
#undef RS_CD_LAUNCH
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return -3;
  }
  *launched = 1;
  return 0;
}

}  // namespace rs
