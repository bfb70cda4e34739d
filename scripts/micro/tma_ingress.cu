// Micro-benchmark: how fast can every SM pull 128B-swizzled 2-D TMA tiles (64 x ROWS bf16 boxes) into shared
// memory, from an L2-resident footprint and from HBM?  One CTA per SM, one issuing thread, S-stage ring, no math.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingress tma_ingress.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../../automative-rag_b200/csrc/tc5.cuh"
using namespace rs;

__global__ void __launch_bounds__(128, 1) ingress(const __grid_constant__ CUtensorMap map, int rows_per_box, int kblocks,
                                                  int stages, long long boxes_per_cta, long long total_row_tiles, int same, int issuers) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t box_bytes = rows_per_box * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)stages * box_bytes);
  tma_prefetch_desc(&map);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // `issuers` threads (lanes of warp 0) each run their own ring over stages t, t + issuers, ...
  if (threadIdx.x < issuers) {
    const int t = threadIdx.x;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    const int my_stages = stages / issuers;
    sm += (size_t)t * my_stages * box_bytes;
    full += t * my_stages;
    stages = my_stages;
    boxes_per_cta /= issuers;
    long long issued = 0, done = 0;
    // CTA b walks row tiles b, b+grid, ... (same=0) or every CTA walks the same tiles (same=1: hot lines)
    auto coords = [&](long long i, int& kx, int& ry) {
      long long tile = (same ? i / kblocks : (blockIdx.x + (i / kblocks * issuers + t) * gridDim.x)) % total_row_tiles;
      kx = (int)(i % kblocks) * 64;
      ry = (int)tile * rows_per_box;
    };
    for (; issued < stages && issued < boxes_per_cta; ++issued) {
      int kx, ry; coords(issued, kx, ry);
      mbar_arrive_expect_tx(&full[issued % stages], box_bytes);
      tma_load_2d(sm + (issued % stages) * box_bytes, &map, kx, ry, &full[issued % stages], pol);
    }
    for (; done < boxes_per_cta; ++done) {
      const int s = (int)(done % stages);
      mbar_wait(&full[s], (uint32_t)((done / stages) & 1));
      if (issued < boxes_per_cta) {
        int kx, ry; coords(issued, kx, ry);
        mbar_arrive_expect_tx(&full[s], box_bytes);
        tma_load_2d(sm + s * box_bytes, &map, kx, ry, &full[s], pol);
        ++issued;
      }
    }
  }
}

int main(int argc, char** argv) {
  const long long rows = argc > 1 ? atoll(argv[1]) : 300000;   // tensor rows (x 256 B per row at d = 128)
  const int d = argc > 2 ? atoi(argv[2]) : 128;
  const int rows_per_box = argc > 3 ? atoi(argv[3]) : 256;
  const int stages = argc > 4 ? atoi(argv[4]) : 4;
  const int same = argc > 5 ? atoi(argv[5]) : 0;
  const int issuers = argc > 6 ? atoi(argv[6]) : 1;
  void* buf;
  cudaMalloc(&buf, (size_t)rows * d * 2);
  cudaMemset(buf, 1, (size_t)rows * d * 2);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)rows_per_box}, es[2] = {1, 1};
  CUresult r = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode failed %d\n", (int)r); return 1; }
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int kblocks = d / 64;
  const long long total_row_tiles = rows / rows_per_box;
  const long long boxes_per_cta = 20000LL * 256 / rows_per_box;
  const size_t smem = 1024 + (size_t)stages * rows_per_box * 128 + 256;
  cudaFuncSetAttribute(ingress, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 2; ++it) {
    cudaEventRecord(a);
    ingress<<<sms, 128, smem>>>(map, rows_per_box, kblocks, stages, boxes_per_cta, total_row_tiles, same, issuers);
    cudaEventRecord(b); cudaEventSynchronize(b);
  }
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)boxes_per_cta * rows_per_box * 128 * sms;
  printf("issuers=%d rows=%lld (%.0f MB) d=%d box=%dx64 stages=%d same=%d: %.2f ms, %.2f TB/s, %.1f B/clk/SM@1.9GHz  err=%s\n",
         issuers, rows, (double)rows * d * 2 / 1e6, d, rows_per_box, stages, same, ms, bytes / ms / 1e9,
         bytes / ms / 1e-3 / sms / 1.9e9, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
