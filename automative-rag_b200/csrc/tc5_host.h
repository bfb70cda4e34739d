// tc5_host.h — host interface of the tcgen05/TMEM/TMA kernels (maxsim_tc5.cu, dense_tc5.cu).
// Owns the cuTensorMapEncodeTiled entry point and per-handle scratch for those kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "kernels.h"

namespace rs {

struct Tc5State;
Tc5State* tc5_create(int device, int num_sms);
void tc5_destroy(Tc5State* s);

// Shared-candidate MaxSim (reference batch_rerank_queries shape, rerankers.py:583-593).
bool tc5_maxsim_supported(const Tc5State* s, int nq, int lq, int d, int nd, const int32_t* cand,
                          const int32_t* out_argmax);
int tc5_maxsim(Tc5State* s, const MaxSimParams& p, int dtype, cudaStream_t stream, int* launched, std::string* err);

// Batched dense top-k (GEMM + fused per-row top-k).
bool tc5_dense_supported(const Tc5State* s, int64_t n, int d, int nq, int k, const uint32_t* mask,
                         int64_t mask_stride_words);
int tc5_dense_topk(Tc5State* s, const void* corpus, int64_t n, int d, int dtype, const float* inv_norm, int metric,
                   const void* queries, int nq, const uint32_t* mask, int k, int64_t id_base, float* out_scores,
                   int64_t* out_ids, cudaStream_t stream, int* launched, std::string* err);

}  // namespace rs
