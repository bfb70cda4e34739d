// maxsim_cand_tc5.cu — per-query-candidate ColBERT MaxSim on tcgen05 tensor cores: the
// retrieve-then-rerank shape, where every query scores ITS OWN candidate documents.
//
// Reference: ColBERTReranker.rerank -> _colbert_rerank -> _compute_maxsim_scores
// (src/core/query/llm/rerankers.py:351-385, :215-265): one query [Lq, D] against the token
// embeddings of the documents retrieved for that query; per document S = Q . D^T (:247), max over
// document tokens (:250), weighted sum over query tokens (:255-261).
//
// Unlike the shared-candidate kernel (maxsim_tc5.cu) no document tile is reused by a second
// query, so the stage is HBM-bound: 2 * D bytes per (query, candidate) token at 32 flops per
// byte.  The kernel is therefore organised around the document stream:
//   * a work item is a (query, candidate) PAIR; the flattened pair list is cut into equal
//     contiguous spans, one per SM.  Pairs are query-major, so a CTA sees few query changes.
//   * a document is cut into CHUNKS of <= 128 tokens (equal parts, rounded up to 32).  A chunk is
//     fetched by TMA as 32-row boxes straight from the packed token buffer into a 5-stage
//     shared-memory ring (lanes of the producer warp issue the boxes of a chunk in parallel);
//     rows past the document's end belong to the next document (or are zero-filled past the
//     buffer) and are masked in the epilogue.
//   * one tcgen05.mma group per chunk: D[128 x n] = Qtile[128 x d] . chunk[n x d]^T with the query
//     zero-padded to 128 rows (rows >= lq_pad of the A tile are zeroed once; TMA refreshes only the
//     first lq_pad rows).  Four 128-column TMEM accumulators decouple the tensor pipe from the
//     epilogue.  The tensor pipe idles most of the time — it only has to keep up with HBM.
//   * epilogue warp q (TMEM lane quarter q, active when lq > 32 q) owns query tokens 32q..32q+31:
//     tcgen05.ld gives each thread its token's scores against 32 document tokens; running max over
//     the document's chunks, then w . max summed over the warp (and over the active warps in a
//     fixed order, so results are run-to-run identical).
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
constexpr int kCdStages = 5;     // chunk ring depth
constexpr int kCdSlots = 4;      // TMEM accumulators (128 columns each)
constexpr int kCdDescRing = 16;  // > stages + slots: a descriptor is dead before its entry is reused
constexpr int kCdTmemCols = 512;
constexpr int kCdEpiBar = 2;     // named barrier of the active epilogue warps

enum : int { kCdFirst = 1, kCdLast = 2, kCdNewQuery = 4, kCdEnd = 8 };

struct CandParams {
  const float* q_weight;       // [nq, lq] or null
  const int32_t* doc_offsets;  // [nd + 1]
  const int32_t* cand;         // [nq, nc] or null (candidate j of every query is document j)
  float* out;                  // [nq, nc]
  int64_t pairs;               // nq * nc
  int32_t nq, lq, lq_pad, nd, nc;
};

struct ChunkDesc {
  int32_t n;      // tokens fed to the MMA (multiple of 32; 0 for an empty document)
  int32_t valid;  // tokens that belong to the document
  int32_t flags;
  int32_t query;
  int64_t pair;   // output index
};

__device__ __forceinline__ float cd_reference_weight(int i, int lq) {
  return (lq > 2 && (i == 0 || i == lq - 1)) ? 0.f : 1.f;  // rerankers.py:255-261
}

__device__ __forceinline__ uint64_t cd_policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

template <bool BF16, int KH>
__global__ void __launch_bounds__(kCdThreads, 1)
    maxsim_cand_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                           const CandParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kKHBytes = 128 * 128;         // one K-half (64 elements) of a 128-row tile
  constexpr uint32_t kTileBytes = kKHBytes * KH;   // A tile and chunk stage alike
  constexpr uint32_t kBoxBytes = kCdBox * 128;

  const int64_t p_begin = p.pairs * blockIdx.x / gridDim.x;
  const int64_t p_end = p.pairs * (blockIdx.x + 1) / gridDim.x;
  if (p_begin >= p_end) return;  // uniform: nothing allocated yet

  // ---- shared memory carve-up (1024-byte aligned: SWIZZLE_128B atoms)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* smA = sm;                        // [2][KH][128 rows x 128 B]
  uint8_t* smB = smA + 2 * kTileBytes;      // [kCdStages][KH][128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + kCdStages * kTileBytes);
  uint64_t* a_full = bars;                       // 2
  uint64_t* a_empty = a_full + 2;                // 2
  uint64_t* b_full = a_empty + 2;                // kCdStages
  uint64_t* b_empty = b_full + kCdStages;        // kCdStages
  uint64_t* acc_full = b_empty + kCdStages;      // kCdSlots
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
    for (int s = 0; s < kCdStages; ++s) {
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
  // rows >= lq_pad of both query tiles stay zero for the whole kernel
  for (uint32_t i = threadIdx.x * 16; i < 2 * kTileBytes; i += kCdThreads * 16)
    *reinterpret_cast<uint4*>(smA + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to TMA / UMMA
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
    const uint32_t a_bytes = (uint32_t)p.lq_pad * 128u * KH;
    int64_t seq = 0;   // chunks issued so far
    int a_seq = -1;    // query switches so far - 1
    int cur_query = -1;
    // document bounds of pair (base + lane), fetched one batch of 32 pairs ahead of their use
    auto fetch = [&](int64_t pair, int& o0, int& o1) {
      o0 = o1 = 0;
      if (pair < p_end) {
        const int j = (int)(pair % p.nc);
        const int doc = p.cand ? __ldg(p.cand + pair) : j;
        if (doc >= 0 && doc < p.nd) {  // an index outside the collection scores as an empty document
          o0 = __ldg(p.doc_offsets + doc);
          o1 = __ldg(p.doc_offsets + doc + 1);
        }
      }
    };
    auto acquire = [&](int& s) {  // wait for the ring stage of chunk `seq`
      s = (int)(seq % kCdStages);
      const uint32_t ph = (uint32_t)(seq / kCdStages) & 1u;
      mbar_wait(&b_empty[s], ph ^ 1u);
    };
    int o0n, o1n;
    fetch(p_begin + lane, o0n, o1n);
    for (int64_t pb = p_begin; pb < p_end; pb += 32) {
      const int o0l = o0n, o1l = o1n;
      fetch(pb + 32 + lane, o0n, o1n);
      const int npairs = (int)min((int64_t)32, p_end - pb);
      for (int i = 0; i < npairs; ++i) {
        const int64_t pair = pb + i;
        const int o0 = __shfl_sync(0xFFFFFFFFu, o0l, i), o1 = __shfl_sync(0xFFFFFFFFu, o1l, i);
        const int query = (int)(pair / p.nc);
        const int ld = max(o1 - o0, 0);
        const int nch = max(1, (ld + kCdBN - 1) / kCdBN);
        const int per = (((ld + nch - 1) / nch) + kCdBox - 1) / kCdBox * kCdBox;  // chunk length, multiple of 32
        for (int c = 0; c < nch; ++c, ++seq) {
          const int valid = max(0, min(per, ld - c * per));
          const int n = (valid + kCdBox - 1) / kCdBox * kCdBox;
          int flags = (c == 0 ? kCdFirst : 0) | (c == nch - 1 ? kCdLast : 0);
          int s;
          acquire(s);
          if (query != cur_query) {  // warp-uniform: stage the new query's tokens in the other A tile
            flags |= kCdNewQuery;
            cur_query = query;
            ++a_seq;
            const int ab = a_seq & 1;
            if (lane == 0) {
              mbar_wait(&a_empty[ab], ((uint32_t)(a_seq >> 1) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&a_full[ab], a_bytes);
              for (int kh = 0; kh < KH; ++kh)
                tma_load_3d(smA + ab * kTileBytes + kh * kKHBytes, &map_q, kh * 64, 0, query, &a_full[ab], pol);
            }
          }
          if (lane == 0) {
            ChunkDesc& dd = desc[seq % kCdDescRing];
            dd.n = n;
            dd.valid = valid;
            dd.flags = flags;
            dd.query = query;
            dd.pair = pair;
            if (n > 0)
              mbar_arrive_expect_tx(&b_full[s], (uint32_t)n * 128u * KH);
            else
              mbar_arrive(&b_full[s]);
          }
          __syncwarp();
          const int nbox = n / kCdBox;
          if (lane < nbox * KH) {
            const int b = lane / KH, kh = lane % KH;
            tma_load_2d(smB + s * kTileBytes + kh * kKHBytes + b * kBoxBytes, &map_d, kh * 64, o0 + c * per + b * kCdBox,
                        &b_full[s], pol);
          }
        }
      }
    }
    {  // END marker travels the same way as a chunk
      int s;
      acquire(s);
      if (lane == 0) {
        ChunkDesc& dd = desc[seq % kCdDescRing];
        dd.n = 0;
        dd.valid = 0;
        dd.flags = kCdEnd;
        dd.query = -1;
        dd.pair = -1;
        mbar_arrive(&b_full[s]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int a_seq = -1;
      for (int64_t seq = 0;; ++seq) {
        const int s = (int)(seq % kCdStages);
        mbar_wait(&b_full[s], (uint32_t)(seq / kCdStages) & 1u);
        const ChunkDesc dd = desc[seq % kCdDescRing];
        const int slot = (int)(seq % kCdSlots);
        mbar_wait(&acc_empty[slot], ((uint32_t)(seq / kCdSlots) & 1u) ^ 1u);
        tc5_fence_after();
        if (dd.flags & kCdNewQuery) {
          if (a_seq >= 0) umma_commit(&a_empty[a_seq & 1]);  // free once every MMA of the previous query is done
          ++a_seq;
          mbar_wait(&a_full[a_seq & 1], (uint32_t)(a_seq >> 1) & 1u);
          tc5_fence_after();
        }
        if (dd.n > 0) {
          const uint32_t idesc = (1u << 4) | ((BF16 ? 1u : 0u) << 7) | ((BF16 ? 1u : 0u) << 10) |
                                 ((uint32_t)(dd.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
          const uint8_t* At = smA + (a_seq & 1) * kTileBytes;
#pragma unroll
          for (int kh = 0; kh < KH; ++kh) {
            const uint64_t da = umma_smem_desc_sw128(smem_u32(At + kh * kKHBytes));
            const uint64_t db = umma_smem_desc_sw128(smem_u32(smB + s * kTileBytes + kh * kKHBytes));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)  // 4 x UMMA_K(16 elements = 32 B) per 128-byte swizzle row
              umma_f16_ss(tmem_base + (uint32_t)slot * kCdBN, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc,
                          (kh | kk) != 0 ? 1u : 0u);
          }
        }
        acc_desc[slot] = dd;
        __threadfence_block();
        umma_commit(&acc_full[slot]);  // fires once the MMAs above (if any) have completed
        umma_commit(&b_empty[s]);
        if (dd.flags & kCdEnd) break;
      }
    }
  } else if (warp >= 4 && warp - 4 < nquad) {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp - 4;  // == warp % 4: TMEM lanes 32*quarter .. +31
    const int tok = quarter * 32 + lane;
    float w = 0.f;
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F;
    int docs_done = 0;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // max over the first `cnt` (1..32) of 32 columns
    auto consume = [&](const uint32_t (&v)[32], int cnt) {
      if (cnt >= 32) {
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
    for (int64_t seq = 0;; ++seq) {
      const int slot = (int)(seq % kCdSlots);
      mbar_wait(&acc_full[slot], (uint32_t)(seq / kCdSlots) & 1u);
      tc5_fence_after();
      const ChunkDesc dd = acc_desc[slot];
      if (dd.flags & kCdEnd) break;
      if (dd.flags & kCdNewQuery) {
        w = 0.f;
        if (tok < p.lq) w = p.q_weight ? __ldg(p.q_weight + (size_t)dd.query * p.lq + tok) : cd_reference_weight(tok, p.lq);
      }
      if (dd.flags & kCdFirst) m0 = m1 = -CUDART_INF_F;
      const uint32_t taddr = tlane + (uint32_t)slot * kCdBN;
      const int nblk = dd.n >> 5;  // 0..4, warp-uniform
      if (nblk > 0) {
        tmem_ld_32x32(taddr, va);
        tmem_ld_wait(va);
        if (nblk > 1) tmem_ld_32x32(taddr + 32, vb);
        consume(va, dd.valid);
        if (nblk > 1) {
          tmem_ld_wait(vb);
          if (nblk > 2) tmem_ld_32x32(taddr + 64, va);
          consume(vb, dd.valid - 32);
          if (nblk > 2) {
            tmem_ld_wait(va);
            if (nblk > 3) tmem_ld_32x32(taddr + 96, vb);
            consume(va, dd.valid - 64);
            if (nblk > 3) {
              tmem_ld_wait(vb);
              consume(vb, dd.valid - 96);
            }
          }
        }
      }
      // the accumulator is in registers (or was never written): hand the TMEM slot back
      tc5_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);
      if (dd.flags & kCdLast) {
        const float m = fmaxf(m0, m1);
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
bool tc5_maxsim_cand_supported(const Tc5State* s, int nq, int lq, int d, int nd, int nc, const int32_t* out_argmax) {
  if (!s || !tc5_has_encode(s)) return false;
  if (out_argmax != nullptr) return false;
  if (d != 64 && d != 128) return false;
  return lq >= 1 && lq <= 128 && nd >= 1 && nc >= 1 && nq >= 1;
}

int tc5_maxsim_cand(Tc5State* s, const MaxSimParams& p, int dtype, cudaStream_t stream, int* launched, std::string* err) {
  *launched = 0;
  const int lq_pad = p.lq <= 32 ? 32 : (p.lq <= 64 ? 64 : (p.lq <= 96 ? 96 : 128));
  const int nc = p.cand ? p.nc : p.nd;
  const long long pairs = (long long)p.nq * nc;
  const int num_sms = tc5_num_sms(s);
  const int grid_x = (int)(pairs < num_sms ? pairs : num_sms);

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
  kp.pairs = pairs;
  kp.nq = p.nq;
  kp.lq = p.lq;
  kp.lq_pad = lq_pad;
  kp.nd = p.nd;
  kp.nc = nc;
  const int kh = p.d / 64;
  const size_t smem = 1024 + (size_t)(2 + kCdStages) * 128 * 128 * kh + 1024;
  dim3 grid(grid_x);
  cudaError_t e = cudaSuccess;
#define RS_CD_LAUNCH(BF, KHV)                                                                                          \
  {                                                                                                                    \
    e = cudaFuncSetAttribute(maxsim_cand_tc5_kernel<BF, KHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) {                                                                                            \
      maxsim_cand_tc5_kernel<BF, KHV><<<grid, kCdThreads, smem, stream>>>(map_q, map_d, kp);                          \
      e = cudaGetLastError();                                                                                          \
    }                                                                                                                  \
  }
  if (dtype == 1) {
    if (kh == 1) RS_CD_LAUNCH(true, 1) else RS_CD_LAUNCH(true, 2)
  } else {
    if (kh == 1) RS_CD_LAUNCH(false, 1) else RS_CD_LAUNCH(false, 2)
  }
#undef RS_CD_LAUNCH
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return -3;
  }
  *launched = 1;
  return 0;
}

}  // namespace rs
