// tc5_host.h — host interface of the tcgen05/TMEM/TMA kernels (maxsim_tc5.cu, dense_tc5.cu).
// Owns the cuTensorMapEncodeTiled entry point and per-handle scratch for those kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "kernels.h"

namespace rs {

struct Tc5State;
Tc5State* tc5_create(int device, int num_sms);
void tc5_destroy(Tc5State* s);

// Shared-candidate MaxSim (reference batch_rerank_queries shape, rerankers.py:583-593).
bool tc5_maxsim_supported(const Tc5State* s, int nq, int lq, int d, int nd, const int32_t* cand,
                          const int32_t* out_argmax);
int tc5_maxsim(Tc5State* s, const MaxSimParams& p, int dtype, cudaStream_t stream, int* launched, std::string* err);

// Per-query-candidate MaxSim (reference rerank shape, rerankers.py:351-385); also serves shared candidates when
// there are too few query tokens for the kernel above (cand == null: candidate j of every query is document j).
bool tc5_maxsim_cand_supported(const Tc5State* s, int nq, int lq, int d, int nd, int nc);
int tc5_maxsim_cand(Tc5State* s, const MaxSimParams& p, int dtype, cudaStream_t stream, int* launched, std::string* err);

// shared helpers (defined in maxsim_tc5.cu)
bool tc5_has_encode(const Tc5State* s);
int tc5_num_sms(const Tc5State* s);
bool tc5_encode(const Tc5State* s, CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, std::string* err);

bool tc5_encode_rows(const Tc5State* s, CUtensorMap* map, const void* base, uint64_t n_rows, uint32_t row_bytes,
                     std::string* err);

// Batched dense top-k (GEMM + fused per-row top-k).
bool tc5_dense_supported(const Tc5State* s, int64_t n, int d, int nq, int k, const uint32_t* mask,
                         int64_t mask_stride_words, bool worthwhile = false);
int tc5_dense_topk(Tc5State* s, const void* corpus, int64_t n, int d, int dtype, const float* inv_norm, int metric,
                   const void* queries, int nq, const uint32_t* mask, int64_t mask_stride_words, int k, int64_t id_base, float* out_scores,
                   int64_t* out_ids, cudaStream_t stream, int* launched, std::string* err,
                   std::vector<int>* redo = nullptr);
// k <= 128: one pass, nothing to redo.  128 < k <= 1024: every corpus range keeps its best <= 128 rows, the merge takes
// the k best of all ranges' lists, and `redo` receives the queries for which a range may have held more of the answer
// than it could keep (the call synchronises the stream to learn that); the caller re-runs those through the scan.

}  // namespace rs
