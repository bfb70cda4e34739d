// maxsim_mma.cu — general ColBERT MaxSim on warp-level tensor-core MMA (mma.sync m16n8k16),
// plus an exact-fp32 CUDA-core variant.
//
// Reference: ColBERTReranker._compute_maxsim_scores (src/core/query/llm/rerankers.py:215-265):
//   similarity = query_emb @ doc_emb.T   (:247)
//   max_sim    = similarity.max(dim=1)   (:250)
//   score      = max_sim[1:-1].sum()  if Lq > 2 else max_sim.sum()   (:255-261)
// The reference launches three kernels and one .item() sync per document; here one CTA
// keeps the query tokens resident in shared memory and streams documents through a
// cp.async double buffer, with row-max and the weighted query-token sum fused after the MMA,
// so the token-score matrix never leaves registers.
//
// This file is the GENERAL path: ragged documents, per-query candidate lists, any
// lq <= 128, any d % 16 == 0, optional argmax (for _explain_colbert_matches :489-492).
// The shared-candidate batched shape (batch_rerank_queries :583-593) has its own
// tcgen05/TMEM kernel in maxsim_tc5.cu.
#include <math_constants.h>

#include "common.cuh"
#include "kernels.h"

namespace rs {

constexpr int kMmaWarps = 4;
constexpr int kMmaThreads = kMmaWarps * 32;

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                                        uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// The reference's content-token rule as a weight (rerankers.py:255-261).
__device__ __forceinline__ float reference_weight(int i, int lq) {
  return (lq > 2 && (i == 0 || i == lq - 1)) ? 0.f : 1.f;
}

// MT m-tiles of 16 query tokens, NT n-tiles of 8 doc tokens per warp per stage.
template <typename T, int MT, int NT, bool ARGMAX>
__global__ void __launch_bounds__(kMmaThreads) maxsim_mma_kernel(const MaxSimParams p, int docs_per_cta) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int TOK = kMmaWarps * NT * 8;  // doc tokens per stage
  const int d = p.d;
  const int ldw = d + 8;                   // padded row length in elements (16 B pad: ldmatrix conflict-free)
  const int lq_pad = MT * 16;
  T* Qs = reinterpret_cast<T*>(smem);
  T* Ds = Qs + (size_t)lq_pad * ldw;  // 2 stages x TOK rows
  float* red_v = reinterpret_cast<float*>(Ds + (size_t)2 * TOK * ldw);  // [kMmaWarps][lq_pad]
  int* red_i = reinterpret_cast<int*>(red_v + kMmaWarps * lq_pad);      // [kMmaWarps][lq_pad] (ARGMAX)
  float* wts = reinterpret_cast<float*>(red_i + (ARGMAX ? kMmaWarps * lq_pad : 0));  // [lq_pad]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int qi = blockIdx.y;
  const int ndo = p.cand ? p.nc : p.nd;
  const int slot0 = blockIdx.x * docs_per_cta;
  const int slot1 = min(slot0 + docs_per_cta, ndo);
  const int vec_per_row = d >> 3;

  // ---- query tokens -> smem (rows >= lq zero-filled), weights
  {
    const T* qg = reinterpret_cast<const T*>(p.q) + (size_t)qi * p.lq * d;
    for (int v = tid; v < lq_pad * vec_per_row; v += kMmaThreads) {
      int r = v / vec_per_row, c = v - r * vec_per_row;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (r < p.lq) val = *reinterpret_cast<const uint4*>(qg + (size_t)r * d + c * 8);
      *reinterpret_cast<uint4*>(Qs + (size_t)r * ldw + c * 8) = val;
    }
    for (int i = tid; i < lq_pad; i += kMmaThreads) {
      float w = 0.f;
      if (i < p.lq) w = p.q_weight ? p.q_weight[(size_t)qi * p.lq + i] : reference_weight(i, p.lq);
      wts[i] = w;
    }
  }

  // first token row and length of the document in output slot `slot`; a candidate index outside the collection is
  // an empty document (score -inf), as in the tcgen05 candidate kernel
  auto doc_bounds = [&](int slot, int& beg, int& len) {
    const int dc = p.cand ? p.cand[(size_t)qi * p.nc + slot] : slot;
    beg = 0;
    len = 0;
    if (dc >= 0 && dc < p.nd) {
      beg = p.doc_offsets[dc];
      len = max(p.doc_offsets[dc + 1] - beg, 0);
    }
  };
  auto issue_tile = [&](int stage, int tok_begin, int ntok) {
    // copy ntok (<= TOK) token rows starting at global token row tok_begin into stage
    const T* src = reinterpret_cast<const T*>(p.doc_tokens) + (size_t)tok_begin * d;
    T* dst = Ds + (size_t)stage * TOK * ldw;
    for (int v = tid; v < ntok * vec_per_row; v += kMmaThreads) {
      int r = v / vec_per_row, c = v - r * vec_per_row;
      cp_async16(dst + (size_t)r * ldw + c * 8, src + (size_t)r * d + c * 8);
    }
  };

  // cursor over (slot, tile) pairs
  int cur_slot = slot0, cur_tile = 0, cur_beg = 0, cur_len = 0;
  if (cur_slot < slot1) {
    doc_bounds(cur_slot, cur_beg, cur_len);
    issue_tile(0, cur_beg, min(TOK, cur_len));
  }
  cp_async_commit();

  float mx[MT][2];
  int mi[MT][2];
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    mx[m][0] = mx[m][1] = -CUDART_INF_F;
    mi[m][0] = mi[m][1] = 0;
  }

  int stage = 0;
  while (cur_slot < slot1) {
    // ---- next (slot, tile)
    int nxt_slot = cur_slot, nxt_tile = cur_tile + 1, nxt_beg = cur_beg, nxt_len = cur_len;
    if (nxt_tile * TOK >= cur_len) {
      nxt_slot = cur_slot + 1;
      nxt_tile = 0;
      if (nxt_slot < slot1) doc_bounds(nxt_slot, nxt_beg, nxt_len);
    }
    if (nxt_slot < slot1) issue_tile(stage ^ 1, nxt_beg + nxt_tile * TOK, min(TOK, nxt_len - nxt_tile * TOK));
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();  // current stage (and Qs on the first pass) visible to all warps

    // ---- S tile = Q (lq_pad x d) . D_tile^T (d x TOK); this warp's columns: warp*NT*8 ..
    const int tok_in_doc0 = cur_tile * TOK + warp * NT * 8;  // doc-relative index of this warp's first column
    float acc[MT][NT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;

    if (tok_in_doc0 < cur_len) {  // warp-uniform: skip warps whose columns are all past the doc end
      const T* Dst = Ds + (size_t)stage * TOK * ldw + (size_t)(warp * NT * 8) * ldw;
      for (int ks = 0; ks < d; ks += 16) {
        uint32_t b[NT][2];
        if constexpr (NT == 1) {
          uint32_t addr = smem_u32(Dst + (size_t)(lane & 7) * ldw + ks + ((lane >> 3) & 1) * 8);
          ldsm_x2(addr, b[0][0], b[0][1]);
        } else {
#pragma unroll
          for (int n2 = 0; n2 < NT / 2; ++n2) {
            uint32_t addr =
                smem_u32(Dst + (size_t)(n2 * 16 + (lane & 7) + (lane >> 4) * 8) * ldw + ks + ((lane >> 3) & 1) * 8);
            ldsm_x4(addr, b[2 * n2][0], b[2 * n2][1], b[2 * n2 + 1][0], b[2 * n2 + 1][1]);
          }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          uint32_t a[4];
          uint32_t addr = smem_u32(Qs + (size_t)(m * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ldw + ks + (lane >> 4) * 8);
          ldsm_x4(addr, a[0], a[1], a[2], a[3]);
#pragma unroll
          for (int n = 0; n < NT; ++n) mma16816<T>(acc[m][n], a, b[n][0], b[n][1]);
        }
      }
      // ---- fused row-max over this warp's columns (columns past the doc end are ignored)
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int c0 = tok_in_doc0 + n * 8 + t4 * 2;  // doc-relative token index of acc[..][n][0]
#pragma unroll
        for (int m = 0; m < MT; ++m) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = c0 + (e & 1);
            const int h = e >> 1;
            float v = acc[m][n][e];
            if (col < cur_len && v > mx[m][h]) {  // strict '>' keeps the FIRST maximal token
              mx[m][h] = v;
              if (ARGMAX) mi[m][h] = col;
            }
          }
        }
      }
    }

    const bool doc_done = (nxt_slot != cur_slot);
    if (doc_done) {
      // quad reduce (lanes sharing g hold different columns of the same rows)
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v = mx[m][h];
          int ix = mi[m][h];
#pragma unroll
          for (int o = 1; o <= 2; o <<= 1) {
            float ov = __shfl_xor_sync(0xFFFFFFFFu, v, o);
            int oi = __shfl_xor_sync(0xFFFFFFFFu, ix, o);
            if (ov > v || (ARGMAX && ov == v && oi < ix)) {
              v = ov;
              ix = oi;
            }
          }
          if (t4 == 0) {
            red_v[warp * lq_pad + m * 16 + h * 8 + g] = v;
            if (ARGMAX) red_i[warp * lq_pad + m * 16 + h * 8 + g] = ix;
          }
          mx[m][h] = -CUDART_INF_F;
          mi[m][h] = 0;
        }
      __syncthreads();
      if (warp == 0) {
        float part = 0.f;
        for (int i = lane; i < lq_pad; i += 32) {
          float v = red_v[i];
          int ix = ARGMAX ? red_i[i] : 0;
#pragma unroll
          for (int w = 1; w < kMmaWarps; ++w) {
            float ov = red_v[w * lq_pad + i];
            int oi = ARGMAX ? red_i[w * lq_pad + i] : 0;
            if (ov > v || (ARGMAX && ov == v && oi < ix)) {
              v = ov;
              ix = oi;
            }
          }
          if (i < p.lq) {
            float w = wts[i];
            if (w != 0.f) part = fmaf(w, v, part);
            if (ARGMAX && p.out_argmax) p.out_argmax[((size_t)qi * ndo + cur_slot) * p.lq + i] = ix;
            if (ARGMAX && p.out_tokmax) p.out_tokmax[((size_t)qi * ndo + cur_slot) * p.lq + i] = v;
          }
        }
        part = warp_sum(part);
        if (lane == 0) p.out_scores[(size_t)qi * ndo + cur_slot] = part;
      }
    }
    __syncthreads();  // everyone done with `stage` (and red_*) before it is refilled
    stage ^= 1;
    cur_slot = nxt_slot;
    cur_tile = nxt_tile;
    cur_beg = nxt_beg;
    cur_len = nxt_len;
  }
  cp_async_wait<0>();
}

size_t maxsim_mma_smem_bytes_impl(int mt, int nt, int d, bool argmax) {
  const int ldw = d + 8, lq_pad = mt * 16, tok = kMmaWarps * nt * 8;
  return (size_t)lq_pad * ldw * 2 + (size_t)2 * tok * ldw * 2 + (size_t)kMmaWarps * lq_pad * 4 * (argmax ? 2 : 1) +
         (size_t)lq_pad * 4;
}

static void mma_config(int lq, int d, int& mt, int& nt) {
  int need = (lq + 15) / 16;
  mt = 1;
  while (mt < need) mt <<= 1;
  nt = (mt <= 2 && d <= 256) ? 4 : 1;
}

size_t maxsim_mma_smem_bytes(int lq, int d) {
  int mt, nt;
  mma_config(lq, d, mt, nt);
  return maxsim_mma_smem_bytes_impl(mt, nt, d, true);
}

template <typename T, int MT, int NT>
static cudaError_t launch_mma_cfg(const MaxSimParams& p, int num_sms, cudaStream_t stream) {
  const bool argmax = p.out_argmax != nullptr || p.out_tokmax != nullptr;
  const int ndo = p.cand ? p.nc : p.nd;
  const size_t smem = maxsim_mma_smem_bytes_impl(MT, NT, p.d, argmax);
  long long pairs = (long long)p.nq * ndo;
  int dpc = (int)(pairs / ((long long)num_sms * 8));
  dpc = dpc < 1 ? 1 : (dpc > 16 ? 16 : dpc);
  dim3 grid((ndo + dpc - 1) / dpc, p.nq);
  cudaError_t e;
  if (argmax) {
    e = cudaFuncSetAttribute(maxsim_mma_kernel<T, MT, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    maxsim_mma_kernel<T, MT, NT, true><<<grid, kMmaThreads, smem, stream>>>(p, dpc);
  } else {
    e = cudaFuncSetAttribute(maxsim_mma_kernel<T, MT, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    if (e != cudaSuccess) return e;
    maxsim_mma_kernel<T, MT, NT, false><<<grid, kMmaThreads, smem, stream>>>(p, dpc);
  }
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_mma_t(const MaxSimParams& p, int num_sms, cudaStream_t stream) {
  int mt, nt;
  mma_config(p.lq, p.d, mt, nt);
  if (mt == 1 && nt == 4) return launch_mma_cfg<T, 1, 4>(p, num_sms, stream);
  if (mt == 2 && nt == 4) return launch_mma_cfg<T, 2, 4>(p, num_sms, stream);
  if (mt == 1) return launch_mma_cfg<T, 1, 1>(p, num_sms, stream);
  if (mt == 2) return launch_mma_cfg<T, 2, 1>(p, num_sms, stream);
  if (mt == 4) return launch_mma_cfg<T, 4, 1>(p, num_sms, stream);
  return launch_mma_cfg<T, 8, 1>(p, num_sms, stream);
}

cudaError_t launch_maxsim_mma(const MaxSimParams& p, int dtype, int num_sms, cudaStream_t stream) {
  if (dtype == 0) return launch_mma_t<__half>(p, num_sms, stream);
  return launch_mma_t<__nv_bfloat16>(p, num_sms, stream);
}

// ================================================================================= fp32 SIMT
// Exact fp32 FMA path for RS_F32 inputs (the reference's CPU dtype; deployed sizes are tiny:
// <= 40 docs x 256 tokens, rerankers.py:32-33, mode_config.py).  One CTA per (query, doc).
// A lane owns ONE query token (blocks of 32 when lq > 32) and a warp 8 document tokens at a time: the query sits
// transposed in shared memory (Qt[e][i]: the lanes read consecutive words), a document token's values arrive as
// float4 loads of one address for the whole warp (a broadcast), every product is an fp32 FMA in ascending e, and the
// running (max, arg max) of a query token stays in its lane's registers — no shuffles, 32 FMAs per 12 loads.
// (Round 1 split d over the lanes and butterfly-reduced every (query token, doc token) pair: 475 us for BASELINE
// config 1, where the arithmetic is 74 MFMA.)
constexpr int kSimtThreads = 128;
constexpr int kSimtTokens = 8;

__global__ void __launch_bounds__(kSimtThreads) maxsim_simt_kernel(const MaxSimParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int d = p.d, lq = p.lq, lqp = (lq + 31) & ~31;
  float* Qt = reinterpret_cast<float*>(smem);           // [d][lqp]
  float* red_v = Qt + (size_t)lqp * d;                  // [4][lqp]
  int* red_i = reinterpret_cast<int*>(red_v + 4 * lqp);  // [4][lqp]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qi = blockIdx.y, slot = blockIdx.x;
  const int ndo = p.cand ? p.nc : p.nd;
  const int dc = p.cand ? p.cand[(size_t)qi * p.nc + slot] : slot;
  const bool in_range = dc >= 0 && dc < p.nd;  // outside the collection: an empty document, score -inf
  const int beg = in_range ? p.doc_offsets[dc] : 0, len = in_range ? max(p.doc_offsets[dc + 1] - beg, 0) : 0;
  const float* qg = reinterpret_cast<const float*>(p.q) + (size_t)qi * lq * d;
  for (int x = tid; x < lqp * d; x += kSimtThreads) {
    const int i = x / d, e = x - i * d;  // coalesced read of Q[i][e]
    Qt[(size_t)e * lqp + i] = i < lq ? qg[x] : 0.f;
  }
  __syncthreads();
  const float* dg = reinterpret_cast<const float*>(p.doc_tokens) + (size_t)beg * d;
  for (int ib = 0; ib < lqp; ib += 32) {
    float best = -CUDART_INF_F;
    int best_j = 0;
    const float* qcol = Qt + ib + lane;
    for (int j0 = warp * kSimtTokens; j0 < len; j0 += 4 * kSimtTokens) {
      float acc[kSimtTokens];
      const float4* rows[kSimtTokens];
#pragma unroll
      for (int t = 0; t < kSimtTokens; ++t) {
        acc[t] = 0.f;
        rows[t] = reinterpret_cast<const float4*>(dg + (size_t)min(j0 + t, len - 1) * d);  // past the end: a repeat
      }
      for (int e = 0; e < d; e += 4) {
        const float q0 = qcol[(size_t)e * lqp], q1 = qcol[(size_t)(e + 1) * lqp];
        const float q2 = qcol[(size_t)(e + 2) * lqp], q3 = qcol[(size_t)(e + 3) * lqp];
#pragma unroll
        for (int t = 0; t < kSimtTokens; ++t) {
          const float4 v = __ldg(rows[t] + (e >> 2));
          acc[t] = fmaf(q0, v.x, acc[t]);
          acc[t] = fmaf(q1, v.y, acc[t]);
          acc[t] = fmaf(q2, v.z, acc[t]);
          acc[t] = fmaf(q3, v.w, acc[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < kSimtTokens; ++t) {
        if (j0 + t < len && acc[t] > best) {  // tokens visited in increasing j: the first maximum is kept
          best = acc[t];
          best_j = j0 + t;
        }
      }
    }
    red_v[warp * lqp + ib + lane] = best;
    red_i[warp * lqp + ib + lane] = best_j;
  }
  __syncthreads();
  if (warp == 0) {
    float part = 0.f;
    for (int i = lane; i < lq; i += 32) {
      float v = red_v[i];
      int ix = red_i[i];
      for (int w = 1; w < 4; ++w) {
        float ov = red_v[w * lqp + i];
        int oi = red_i[w * lqp + i];
        if (ov > v || (ov == v && oi < ix)) {
          v = ov;
          ix = oi;
        }
      }
      float w = p.q_weight ? p.q_weight[(size_t)qi * lq + i] : reference_weight(i, lq);
      if (w != 0.f) part = fmaf(w, v, part);
      if (p.out_argmax) p.out_argmax[((size_t)qi * ndo + slot) * lq + i] = ix;
      if (p.out_tokmax) p.out_tokmax[((size_t)qi * ndo + slot) * lq + i] = v;
    }
    // fixed-order sum over lanes so the result is run-to-run deterministic
    part = warp_sum(part);
    if (lane == 0) p.out_scores[(size_t)qi * ndo + slot] = part;
  }
}

size_t maxsim_simt_smem_bytes(int lq, int d) {
  const size_t lqp = (size_t)((lq + 31) & ~31);
  return lqp * d * 4 + (size_t)8 * lqp * 4;
}

cudaError_t launch_maxsim_simt(const MaxSimParams& p, cudaStream_t stream) {
  const int ndo = p.cand ? p.nc : p.nd;
  const size_t smem = maxsim_simt_smem_bytes(p.lq, p.d);
  cudaError_t e = cudaFuncSetAttribute(maxsim_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(ndo, p.nq);
  maxsim_simt_kernel<<<grid, kSimtThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace rs
