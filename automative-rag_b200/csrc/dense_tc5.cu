// dense_tc5.cu — batched dense top-k (many queries x corpus GEMM with a fused per-row top-k).
// Placeholder until the tcgen05 kernel lands: reports "unsupported" so rs_dense_topk uses the
// single-query scan for every query.
#include "tc5_host.h"

namespace rs {

bool tc5_dense_supported(const Tc5State*, int64_t, int, int, int, const uint32_t*, int64_t) { return false; }

int tc5_dense_topk(Tc5State*, const void*, int64_t, int, int, const float*, int, const void*, int, const uint32_t*, int,
                   int64_t, float*, int64_t*, cudaStream_t, int* launched, std::string* err) {
  *launched = 0;
  *err = "batched tcgen05 dense path not built";
  return -2;
}

}  // namespace rs
