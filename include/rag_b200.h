/*
 * rag_b200.h — C ABI of the B200-native retrieval-scoring engine.
 *
 * This is the drop-in boundary for the two data-parallel scoring stages of
 * jliang87/Automative-RAG.  Every entry point is `extern "C"`,
 * takes plain pointers and sizes, never throws, and returns an `int` status
 * (RS_OK == 0, negative on error; `rs_last_error` gives the text).  All data
 * pointers are DEVICE pointers owned by the caller unless the name ends in
 * `_host`; the engine owns only the opaque `rs_handle`.
 *
 * Reference interfaces each entry point replaces (paths relative to the
 * reference checkout):
 *
 *   rs_dense_topk / rs_dense_topk_host
 *       the arithmetic behind QdrantStore.similarity_search_with_score
 *       (src/core/query/retrieval/vectorstore.py:166-214), i.e. the Qdrant
 *       cosine / dot search that langchain_qdrant reaches through
 *       qdrant_client.query_points (vectorstore.py:192-196,202-205,209-212);
 *       collection created with Distance.COSINE (vectorstore.py:52-57,75-81).
 *   mask argument of rs_dense_topk
 *       the predicate QdrantStore._build_filter builds
 *       (vectorstore.py:216-276), evaluated to one bit per corpus row.
 *   rs_filter_mask
 *       evaluates that predicate on device over columnar metadata
 *       (payload fields indexed at vectorstore.py:89-122).
 *   rs_maxsim
 *       ColBERTReranker._compute_maxsim_scores
 *       (src/core/query/llm/rerankers.py:215-265): matmul :247, row-max :250,
 *       content-token sum :255-261.  `out_argmax` serves
 *       _explain_colbert_matches (rerankers.py:489-492).
 *   rs_maxsim_list
 *       the same function in its own call shape: one query, a Python list of per-document
 *       tensors in, a list of floats out (rerankers.py:215-217,244-263).
 *   rs_rerank_postprocess
 *       the sort / min-max / 0.8·colbert+0.2·bge blend / [:top_k] tail of
 *       ColBERTReranker.rerank (rerankers.py:302-343,377-380).
 *   rs_topk_merge
 *       has no reference counterpart (the reference is single-GPU); it merges
 *       per-shard top-k lists after the one all-gather of SURVEY.md §8e.
 *   rs_comm_*, rs_allgather_topk, rs_allgather, rs_allreduce_max_f32,
 *   rs_dense_topk_sharded_host, rs_owned_candidates
 *       the one exchange step per sharded stage (SURVEY.md §8b "rs_allgather_topk(comm, ...)",
 *       §8e) as the engine's own kernels over NVLink peer memory; no reference counterpart.
 *
 * Result order everywhere: score descending, ties by ascending id — a valid
 * refinement of the reference's stable `sorted(..., reverse=True)` for rerank
 * (rerankers.py:377-380) and of Qdrant's unspecified tie order.
 *
 * Threading: one handle per (process, device); calls on one handle must be
 * serialised by the caller.  Kernels are stream-ordered on the `stream`
 * argument (a cudaStream_t passed as void*; NULL = legacy default stream).
 * All launches of a handle share its workspaces: when consecutive calls name
 * DIFFERENT streams the engine orders the later call after the earlier one with
 * an event, so calls never overlap on the device whatever streams they use.  A
 * stream passed to the handle must stay alive until the next call on the handle.
 */
#ifndef RAG_B200_H_
#define RAG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_ABI_VERSION 2

typedef struct rs_handle rs_handle;

/* status codes */
enum {
  RS_OK = 0,
  RS_ERR_INVALID_ARG = -1,   /* NULL pointer, negative size, misaligned buffer ...      */
  RS_ERR_UNSUPPORTED = -2,   /* shape / dtype outside what the kernels implement          */
  RS_ERR_CUDA = -3,          /* a CUDA runtime / driver call failed                       */
  RS_ERR_NO_DEVICE = -4,     /* no sm_100 device visible: there is NO CPU fallback        */
  RS_ERR_NOMEM = -5,
  RS_ERR_COMM = -6           /* a peer's wire block could not be mapped (CUDA IPC / peer access)  */
};

/* element types of embedding buffers */
enum { RS_F16 = 0, RS_BF16 = 1, RS_F32 = 2 };

/* similarity metrics of the dense stage */
enum {
  RS_METRIC_IP = 0,      /* raw inner product                                            */
  RS_METRIC_COSINE = 1   /* <q/|q|, c/|c|>; |c| from inv_norm[] or assumed 1 when NULL    */
};

/* which MaxSim kernel family to use (RS_MAXSIM_AUTO picks by shape) */
enum { RS_MAXSIM_AUTO = 0, RS_MAXSIM_MMA = 1, RS_MAXSIM_TCGEN05 = 2, RS_MAXSIM_SIMT = 3, RS_MAXSIM_TCGEN05_CAND = 4 };

/* which dense kernel family to use (RS_DENSE_AUTO picks by nq) */
enum { RS_DENSE_AUTO = 0, RS_DENSE_SCAN = 1, RS_DENSE_TCGEN05 = 2 };

int rs_abi_version(void);

/* Create / destroy an engine handle bound to CUDA device `device`.
 * Fails with RS_ERR_NO_DEVICE when the device is absent or is not compute capability 10.x. */
int rs_create(int device, rs_handle** out);
int rs_destroy(rs_handle* h);

/* Text of the last error on this handle (or of the last failed rs_create when h == NULL). */
const char* rs_last_error(const rs_handle* h);

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t rs_launch_count(const rs_handle* h);

/* Force a kernel family (RS_*_AUTO restores shape-based choice). */
int rs_set_dense_impl(rs_handle* h, int impl);
int rs_set_maxsim_impl(rs_handle* h, int impl);
/*
 * Diagnostics: attach (or, with NULL, detach) a device buffer of 8 x num_sms x 8 uint64 into which
 * the single-query scan stamps %globaltimer at 8 points of every CTA's life (entry, barriers ready,
 * query loaded, first tile landed, last tile consumed, local top-k sorted, list published, merge
 * done); launch i writes block (i mod 8).  Used by scripts/scan_trace.py; off by default.
 */
int rs_set_scan_trace(rs_handle* h, uint64_t* trace_dev);
/*
 * Diagnostics, no device needed: the shared-memory plan the single-query scan uses for row length d and list length
 * k — out7 = { rows per tile, consumer warps, ring slots, top-k buffer capacity, its high-water mark, consumer
 * rounds between two buffer checks, dynamic shared-memory bytes }.  tests/test_abi.py checks its invariants for every
 * supported (d, k) on CPU.
 */
int rs_scan_plan(int32_t d, int32_t k, int64_t* out7);
/*
 * The plan of a launch inside a multi-query rs_dense_topk call (its launches are chained): out8[0..6] as
 * above, out8[7] = 1 when the CTA is sized for half an SM so that consecutive queries' CTAs share an SM and one query's
 * start-up and wind-down overlap its neighbour's streaming (d <= 1024 and k small enough for a 12-slot half ring),
 * 0 when the call uses the single-launch plan.
 */
int rs_scan_plan_chained(int32_t d, int32_t k, int64_t* out8);
/* Family used by the most recent rs_dense_topk / rs_maxsim call on this handle. */
int rs_last_dense_impl(const rs_handle* h);
/* Queries the most recent batched rs_dense_topk call with k > 128 re-ran through the single-query scan (usually 0). */
int rs_last_dense_redo(const rs_handle* h);
int rs_last_maxsim_impl(const rs_handle* h);

/*
 * Per-call statistics (the counterpart of the reference's per-request debug timing fields `search_time_ms` /
 * `docs_per_second`, src/services/system_service.py:336-372).  Off by default; with rs_set_profiling(h, 1) every
 * scoring entry point brackets its launches with CUDA events on the call's stream (two event records per call, no
 * synchronisation).  rs_last_call_stats waits for the most recent profiled call to finish and reports it.
 */
typedef struct rs_call_stats {
  int32_t entry;          /* RS_CALL_* of the call                                                          */
  int32_t kernel_family;  /* RS_DENSE_* / RS_MAXSIM_* the call used (0 for the others)                      */
  int32_t launches;       /* kernels launched by the call                                                   */
  int32_t queries;        /* nq of the call                                                                 */
  int64_t bytes_scanned;  /* algorithmic bytes: corpus rows (dense) / document tokens (MaxSim) the call reads */
  double flops;           /* 2*M*N*K of the call                                                             */
  float device_ms;        /* device time from the call's first launch to the end of its last                */
  float merge_ms;         /* of which the merge of the per-range lists (batched dense) or 0                 */
} rs_call_stats;
enum { RS_CALL_NONE = 0, RS_CALL_DENSE_TOPK = 1, RS_CALL_MAXSIM = 2, RS_CALL_TOPK_MERGE = 3, RS_CALL_ALLGATHER_TOPK = 4 };
int rs_set_profiling(rs_handle* h, int on);
int rs_last_call_stats(rs_handle* h, rs_call_stats* out);

/*
 * Exact brute-force top-k over a row-major corpus [n, d].
 *
 *   corpus      [n, d] dtype (RS_F16 | RS_BF16), 16-byte aligned, d % 8 == 0
 *   inv_norm    [n] fp32 1/|row| or NULL (rows already unit length / metric IP)
 *   queries     [nq, d] same dtype as corpus
 *   mask        bit-packed row filter, LSB-first: bit (i & 31) of word (i >> 5) set
 *               <=> row i passes.  NULL = every row passes.  Query j uses the words
 *               at mask + j * mask_stride_words (stride 0 = one mask for all queries).
 *   k           1 <= k <= 2048
 *   id_base     added to the local row index to form the returned id (shard offset)
 *   out_scores  [nq, k] fp32, descending; -inf where fewer than k rows pass
 *   out_ids     [nq, k] int64;  -1 where fewer than k rows pass
 *
 * Kernel families (rs_set_dense_impl; AUTO picks): a single-query scan launch per query (any k), or — from two
 * queries on — ONE batched tcgen05 pass for the whole call: k <= 128 directly; 128 < k <= 1024 with every corpus
 * range keeping its best <= 128 rows and an exactness check afterwards, in which case the call synchronises the
 * stream once and re-runs through the scan the queries whose answer a range could not hold (rs_last_dense_redo).
 */
int rs_dense_topk(rs_handle* h, const void* corpus, int64_t n, int32_t d, int32_t dtype,
                  const float* inv_norm, int32_t metric, const void* queries, int32_t nq,
                  const uint32_t* mask, int64_t mask_stride_words, int32_t k, int64_t id_base,
                  float* out_scores, int64_t* out_ids, void* stream);

/* Same search with HOST query / mask / output buffers (the call the Python adapter makes
 * per request): copies queries (+mask when given) host->device, runs rs_dense_topk, brings the
 * k results back and synchronises — all on `stream`, the CALLER's stream, so the search is
 * ordered after whatever the caller enqueued there before (rs_filter_mask writing mask_dev,
 * appends to the corpus, tombstone updates).  The corpus and inv_norm stay resident on the
 * device.  mask_host may be NULL; mask_dev is used when mask_host is NULL (either may be
 * NULL = no filter). */
int rs_dense_topk_host(rs_handle* h, const void* corpus, int64_t n, int32_t d, int32_t dtype,
                       const float* inv_norm, int32_t metric, const void* queries_host,
                       int32_t nq, const uint32_t* mask_host, const uint32_t* mask_dev,
                       int64_t mask_stride_words, int32_t k, int64_t id_base,
                       float* out_scores_host, int64_t* out_ids_host, void* stream);

/*
 * Merge `nlists` per-shard top-k lists into one.  Inputs laid out as the all-gather
 * leaves them: scores/ids [nlists, nq, k_in], list l starting at scores + l * score_list_stride
 * and ids + l * id_list_stride (strides in ELEMENTS; 0 = dense, i.e. nq * k_in), so the views
 * into the gathered wire buffer are consumed without a copy.  Entries with id < 0 are padding.
 * Lists are expected in descending score order (what rs_dense_topk writes); other orders give the
 * same result, only slower.
 * Output [nq, k_out] in (score desc, id asc) order, padded with (-inf, -1).
 * nlists * k_in <= 16384.
 */
int rs_topk_merge(rs_handle* h, const float* scores, const int64_t* ids, int32_t nlists,
                  int32_t nq, int32_t k_in, int32_t k_out, int64_t score_list_stride,
                  int64_t id_list_stride, float* out_scores, int64_t* out_ids, void* stream);

/*
 * ColBERT late-interaction MaxSim.
 *
 *   q            [nq, lq, d] dtype
 *   q_weight     [nq, lq] fp32 per-query-token weight, or NULL for the reference rule
 *                (rerankers.py:255-261): lq > 2 -> drop token 0 and token lq-1, else all ones
 *   doc_tokens   [n_tokens, d] dtype, packed;  doc i = rows doc_offsets[i] .. doc_offsets[i+1]
 *   n_tokens     rows in doc_tokens (>= doc_offsets[nd]); bounds the TMA tensor map
 *   doc_offsets  [nd + 1] int32 (device), non-decreasing; an empty document scores -inf
 *   cand         NULL: every query scores all nd docs (the reference's batch_rerank_queries
 *                shape, rerankers.py:583-593); else [nq, nc] int32 doc indices per query; an
 *                index outside [0, nd) is an empty document (score -inf) on every kernel
 *   out_scores   [nq, nd] (cand == NULL) or [nq, nc] fp32, in input order
 *   out_argmax   NULL, or int32 [nq, nd|nc, lq]: index within the doc of the max token (the first
 *                one on ties) — _explain_colbert_matches, rerankers.py:489-492
 *   out_tokmax   NULL, or fp32 [nq, nd|nc, lq]: max_j <Q[q,i,:], D[doc,j,:]> itself, the
 *                "similarity" the explanations report per query token (rerankers.py:493-501)
 *
 *   score(q, doc) = sum_i w[q,i] * max_j <Q[q,i,:], D[doc,j,:]>,  fp32 accumulation.
 *   dtype RS_F32 runs an exact-fp32 CUDA-core kernel (small shapes, deployed sizes).
 */
int rs_maxsim(rs_handle* h, const void* q, int32_t nq, int32_t lq, int32_t d, int32_t dtype,
              const float* q_weight, const void* doc_tokens, int64_t n_tokens,
              const int32_t* doc_offsets, int32_t nd, const int32_t* cand, int32_t nc,
              float* out_scores, int32_t* out_argmax, float* out_tokmax, void* stream);

/*
 * The same scoring in the exact call shape of ColBERTReranker._compute_maxsim_scores
 * (rerankers.py:215-265): ONE query against a LIST of separately allocated document matrices, scores back
 * in host memory.  q is [lq, d] and docs[i] is [doc_lens[i], d], all of element type src_dtype, each on
 * the host or on this handle's device (q_on_host / docs_on_host).  The engine stages the list in one upload
 * (raw tokens included when they are host-resident), packs and converts it to compute_dtype in one launch,
 * scores it with rs_maxsim (RS_F32 = the exact-fp32 kernel, the reference's CPU dtype; RS_F16 = what the
 * reference computes under autocast on CUDA) and returns after the nd scores are in out_scores_host.
 * q_weight_host: [lq] fp32 or NULL for the reference rule.  A zero-length document scores -inf.
 */
int rs_maxsim_list(rs_handle* h, const void* q, int32_t q_on_host, int32_t lq, int32_t d, int32_t src_dtype,
                   int32_t compute_dtype, const float* q_weight_host, const void* const* docs,
                   const int32_t* doc_lens, int32_t nd, int32_t docs_on_host, float* out_scores_host,
                   void* stream);

/*
 * Rerank tail (rerankers.py:302-343): per query row of `scores` [nq, n]:
 *   order by score desc (stable: ties keep input order);
 *   if other != NULL: min-max normalise scores over the row (all-equal -> 1.0), min-max
 *   normalise `other` the same way, blend w_a * a + w_b * b, re-sort (stable);
 *   write the first top_k (index into the input row, final score).
 *   out_idx [nq, top_k] int32, out_scores [nq, top_k] fp32.   n <= 16384.
 */
int rs_rerank_postprocess(rs_handle* h, const float* scores, const float* other, int32_t nq,
                          int32_t n, float w_a, float w_b, int32_t top_k, int32_t* out_idx,
                          float* out_scores, void* stream);

/*
 * Evaluate a _build_filter predicate (vectorstore.py:216-276) over columnar metadata into
 * the bit-packed mask rs_dense_topk consumes.  The predicate is an AND over `nclauses`
 * clauses; clause c tests int32 column `cols[c]` ([n], device; keyword fields are
 * dictionary-encoded by the host) against the value set
 * `values[val_offsets[c] .. val_offsets[c+1])` (OR within a clause == the nested
 * Filter(should=[MatchValue...]); a `year` Range(gte=v,lte=v) is the one-element set {v}).
 * `tombstone` (bit-packed, may be NULL) marks deleted rows, which never pass.
 *   cols         host array of `nclauses` device pointers
 *   values       host int32 array;  val_offsets host int32 [nclauses + 1]
 *   out_mask     device uint32 [ceil(n / 32)]
 */
int rs_filter_mask(rs_handle* h, const int32_t* const* cols, int32_t nclauses,
                   const int32_t* values, const int32_t* val_offsets, const uint32_t* tombstone,
                   int64_t n, uint32_t* out_mask, void* stream);

/*
 * ---- Multi-GPU exchange over NVLink / NVSwitch peer memory -------------------
 *
 * One process (or thread) per GPU, one handle per GPU.  Setup, once:
 *   1. every rank calls rs_comm_export: allocates this rank's WIRE BLOCK (2 x world slots of
 *      `slot_bytes`, plus flags) and fills a 128-byte handle describing it;
 *   2. the ranks exchange the handles by any host-side means (torch.distributed, MPI, a file ...);
 *   3. every rank calls rs_comm_open with all `world` handles in rank order: peers' blocks are
 *      mapped with CUDA IPC (ranks in other processes) or direct peer access (same process).
 * After that a collective is ONE kernel per rank: push the local contribution into every peer's
 * wire block over NVLink, raise a flag, wait for the peers' flags, consume out of local memory —
 * no NCCL, no host round trip.  Every rank must call the same collectives in the same order;
 * calls are stream-ordered like every other entry point.  A peer that never arrives makes the
 * kernel trap after 20 s (environment RS_COMM_TIMEOUT_MS; RS_ERR_CUDA on the next call) instead of
 * hanging the device.
 * slot_bytes bounds one rank's contribution per call: nq * k_in * 12 bytes for rs_allgather_topk.
 */
#define RS_COMM_HANDLE_BYTES 128
#define RS_COMM_MAX_WORLD 8

int rs_comm_export(rs_handle* h, int32_t world, int32_t rank, int64_t slot_bytes, void* out_handle /* [128] */);
int rs_comm_open(rs_handle* h, const void* handles /* [world][RS_COMM_HANDLE_BYTES], rank order */);
int rs_comm_close(rs_handle* h); /* unmaps the peers and frees the wire block; barrier across ranks first */
/* world == 0 when no exchange is open */
int rs_comm_info(const rs_handle* h, int32_t* world, int32_t* rank, int64_t* slot_bytes);

/*
 * The exchange of the dense stage: every rank contributes its LOCAL top-k lists
 * (local_scores / local_ids [nq, k_in], what rs_dense_topk wrote for its shard, ids already global
 * through id_base) and receives the merged global top-k_out of every query — gather and k-way
 * merge fused in one kernel.  Output order (score desc, id asc), identical on every rank and
 * bit-identical to rs_topk_merge over the gathered lists.  world * k_in <= 16384.
 */
int rs_allgather_topk(rs_handle* h, const float* local_scores, const int64_t* local_ids, int32_t nq,
                      int32_t k_in, int32_t k_out, float* out_scores, int64_t* out_ids, void* stream);

/* Plain all-gather: out [world][bytes] <- every rank's `local` (bytes % 16 == 0, 16-byte aligned). */
int rs_allgather(rs_handle* h, const void* local, int64_t bytes, void* out, void* stream);

/* Element-wise max over ranks of fp32 vectors — the exchange of the MaxSim stage when every
 * candidate is scored by the one rank that owns its token embeddings and all others contribute
 * -inf: out[i] = max over ranks of local[i]. */
int rs_allreduce_max_f32(rs_handle* h, const float* local, int64_t n, float* out, void* stream);
/*
 * Sharded MaxSim over per-query candidate lists (the reference's retrieve-then-rerank composition,
 * tests/test_retrieval.py:206-258 with rerankers.py:351-385, run on G GPUs): documents are owned round-robin
 * (owner = id % world, local index = id / world).  Maps n global candidate ids (int32, or int64 when cand_is_i64;
 * negative = padding; taken modulo `pool` first when pool > 0) to this rank's local document indices, -1 for
 * candidates another rank owns — which rs_maxsim scores -inf — so that rs_allreduce_max_f32 of the score blocks is
 * the merged result.  One launch on `stream`.
 */
int rs_owned_candidates(rs_handle* h, const void* cand, int32_t cand_is_i64, int64_t n, int32_t world, int32_t rank,
                        int64_t pool, int32_t* out_local, void* stream);

/* rs_dense_topk_host against a row-sharded corpus: host queries in, merged global (score, id)
 * pairs out in host memory on every rank; H2D, local scan, the fused exchange above and the
 * synchronise all on `stream`.  mask_dev covers this rank's rows. */
int rs_dense_topk_sharded_host(rs_handle* h, const void* corpus, int64_t n, int32_t d, int32_t dtype,
                               const float* inv_norm, int32_t metric, const void* queries_host,
                               int32_t nq, const uint32_t* mask_dev, int64_t mask_stride_words,
                               int32_t k, int64_t id_base, float* out_scores_host,
                               int64_t* out_ids_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RAG_B200_H_ */
