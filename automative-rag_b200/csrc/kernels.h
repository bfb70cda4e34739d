// kernels.h — host-visible launchers of the sm_100a kernels (internal to the library).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rs {

// ------------------------------------------------------------------ dense single-query scan
struct ScanParams {
  const void* corpus;      // [n, d] fp16 | bf16
  const void* query;       // [d] same dtype
  const float* inv_norm;   // [n] or null
  const uint32_t* mask;    // bit-packed or null
  int64_t n;
  int32_t d;
  int32_t k;
  int32_t metric;          // 0 ip, 1 cosine
  int64_t id_base;
  uint64_t* ws_keys;       // [grid, k]
  unsigned* ticket;        // zero-initialised once, self-resetting
  float* out_scores;       // [k]
  int64_t* out_ids;        // [k]
  int32_t tile_rows;       // filled by the launcher
  int32_t stages;          // ring slots, filled by the launcher
  int32_t consumers;       // consumer warps, filled by the launcher (stages is a multiple of it)
  int32_t buf_cap;         // top-k buffer capacity, filled by the launcher
  int32_t buf_hw;          // its high-water mark (capacity - appends possible between two checks)
  int32_t rounds_per_check;  // consumer rounds between two buffer checks
  int32_t l2_policy;       // 0 evict_first (default), 1 normal, 2 evict_last
  unsigned long long* unit_counter;  // next unclaimed mask word of this launch; 0 on entry, reset by the merging CTA
  int32_t unit_words;      // smallest grab in mask words (32 rows each), filled by the launcher
  int32_t grab_max;        // largest grab in mask words (<= 32), filled by the launcher
  int32_t first_words;     // words per CTA handed out statically before the counter is used (0 = none)
  uint64_t* trace;         // diagnostics (rs_set_scan_trace): [grid][8] %globaltimer stamps, or null
  int32_t batch_max;       // filled by the launcher: ring slots claimed per producer pass, contiguous tiles (<= 8)
  int32_t gather_batch;    // same for gather tiles
  int32_t gather4;         // filled by the launcher: gather tiles go out as TMA tile::gather4 (row tensor map given)
};
int scan_tile_rows(int d);
size_t scan_smem_bytes(int d, int k);
// shared-memory plan of a scan launch: out[0..6] = tile_rows, consumers, stages, buf_cap, buf_hw, rounds_per_check,
// dynamic shared-memory bytes (host-only arithmetic; exported through rs_scan_plan for the CPU tests)
// chained: the plan of a launch inside a multi-query call (two CTAs per SM when scan_coresident_ok); then out has 8
// entries and out[7] = 1 when the co-resident plan applies
void scan_plan_query(int d, int k, int64_t* out, bool chained = false);
bool scan_coresident_ok(int d, int k);
// pdl: launch with programmatic stream serialization (only between consecutive scans of one call,
// whose inputs are all complete before the first launch; see dense_scan.cu)
// gather_map: tensor map of the corpus as [n][2d / 8] 8-byte elements, box = one row (tc5_encode_rows), or null:
// with a filter mask the passing rows of sparse mask words are then fetched four per TMA instruction.
bool scan_gather4_supported(int d);
cudaError_t launch_dense_scan(ScanParams p, int dtype, int num_sms, bool pdl, cudaStream_t stream,
                              const CUtensorMap* gather_map = nullptr, bool chained = false);

// ------------------------------------------------------------------ top-k list merge
cudaError_t launch_topk_merge(const float* scores, const int64_t* ids, int nlists, int nq, int k_in, int k_out,
                              int64_t score_list_stride, int64_t id_list_stride, float* out_scores, int64_t* out_ids,
                              cudaStream_t stream, const uint32_t* bound = nullptr, int bound_groups = 0,
                              const float* bound_scale = nullptr);

// ------------------------------------------------------------------ MaxSim
struct MaxSimParams {
  const void* q;               // [nq, lq, d]
  const float* q_weight;       // [nq, lq] or null (reference rule)
  const void* doc_tokens;      // [T, d]
  const int32_t* doc_offsets;  // [nd + 1]
  const int32_t* cand;         // null or [nq, nc]
  float* out_scores;           // [nq, ndo]   ndo = cand ? nc : nd
  int32_t* out_argmax;         // null or [nq, ndo, lq]
  float* out_tokmax;           // null or [nq, ndo, lq]: the per-query-token maxima themselves
  int64_t n_tokens;            // rows in doc_tokens
  int32_t nq, lq, d, nd, nc;
};
// warp-level mma.sync path: any lq <= 128, d % 16 == 0, ragged docs, candidate lists, argmax
cudaError_t launch_maxsim_mma(const MaxSimParams& p, int dtype, int num_sms, cudaStream_t stream);
size_t maxsim_mma_smem_bytes(int lq, int d);
// exact fp32 CUDA-core path (dtype RS_F32)
cudaError_t launch_maxsim_simt(const MaxSimParams& p, cudaStream_t stream);
size_t maxsim_simt_smem_bytes(int lq, int d);

// ------------------------------------------------------------------ rerank tail, filter mask
cudaError_t launch_rerank_postprocess(const float* scores, const float* other, int nq, int n, float w_a, float w_b,
                                      int top_k, int32_t* out_idx, float* out_scores, cudaStream_t stream);
cudaError_t launch_filter_mask(const int32_t* const* cols_dev, int nclauses, const int32_t* values_dev,
                               const int32_t* val_offsets_dev, const uint32_t* tombstone, int64_t n,
                               uint32_t* out_mask, int num_sms, cudaStream_t stream);

// global candidate ids -> local document indices of the owning rank (round-robin ownership), -1 elsewhere
cudaError_t launch_owned_candidates(const void* cand, int is_i64, int64_t n, int world, int rank, int64_t pool,
                                    int32_t* out, cudaStream_t stream);

// document list -> packed tokens in the compute dtype (ColBERTReranker._compute_maxsim_scores call shape)
cudaError_t launch_gather_docs(const void* const* ptrs_dev, const uint8_t* staged_dev, const int64_t* src_off_dev,
                               const int32_t* offsets_dev, int nd, int max_len, int d, int src_dtype, int dst_dtype,
                               void* dst, cudaStream_t stream);

}  // namespace rs
