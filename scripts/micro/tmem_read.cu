// Micro-benchmark: how fast can one SM move fp32 accumulators from TMEM to registers (tcgen05.ld)?
// This is the roof of every epilogue that has to look at each score once — MaxSim at d = 128 does 64 flops per
// accumulator byte, so the tensor pipe can only run as fast as TMEM can be drained.
// One CTA per SM, W warps (warp w reads lane quarter w % 4), each warp loops over the 512 columns with
// 32x32b.x32 or .x64 loads, waiting after every load ("dep") or after every second one ("pipelined").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read tmem_read.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../../automative-rag_b200/csrc/tc5.cuh"
using namespace rs;

__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_only() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: x32, wait after each load, 16 FMNMX3 per load (the MaxSim epilogue's inner loop)
// mode 1: x32, two loads in flight
// mode 2: x64, wait after each
// mode 3: x32, wait after each, NO math (pure drain)
__global__ void __launch_bounds__(512, 1) tmem_read(int mode, int iters, float* sink, long long* cycles) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  float m0 = -1e30f, m1 = -1e30f, m2 = -1e30f, m3 = -1e30f;
  const long long t0 = clock64();
  if (mode == 0 || mode == 3) {
    uint32_t v[32];
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
      for (int c = 0; c < 512; c += 32) {
        tmem_ld_32x32(base + c, v);
        tmem_ld_wait(v);
        if (mode == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            m0 = fmaxf(fmaxf(m0, __uint_as_float(v[i + 0])), __uint_as_float(v[i + 1]));
            m1 = fmaxf(fmaxf(m1, __uint_as_float(v[i + 2])), __uint_as_float(v[i + 3]));
            m2 = fmaxf(fmaxf(m2, __uint_as_float(v[i + 4])), __uint_as_float(v[i + 5]));
            m3 = fmaxf(fmaxf(m3, __uint_as_float(v[i + 6])), __uint_as_float(v[i + 7]));
          }
        } else {
          m0 = fmaxf(m0, __uint_as_float(v[lane]));
        }
      }
    }
  } else if (mode == 1) {
    uint32_t va[32], vb[32];
    for (int it = 0; it < iters; ++it) {
      tmem_ld_32x32(base, va);
#pragma unroll 1
      for (int c = 0; c < 512; c += 64) {
        tmem_ld_32x32(base + c + 32, vb);
        tmem_ld_wait(va);
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          m0 = fmaxf(fmaxf(m0, __uint_as_float(va[i + 0])), __uint_as_float(va[i + 1]));
          m1 = fmaxf(fmaxf(m1, __uint_as_float(va[i + 2])), __uint_as_float(va[i + 3]));
          m2 = fmaxf(fmaxf(m2, __uint_as_float(va[i + 4])), __uint_as_float(va[i + 5]));
          m3 = fmaxf(fmaxf(m3, __uint_as_float(va[i + 6])), __uint_as_float(va[i + 7]));
        }
        tmem_ld_32x32(base + ((c + 64) & 511), va);
        tmem_ld_wait(vb);
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          m0 = fmaxf(fmaxf(m0, __uint_as_float(vb[i + 0])), __uint_as_float(vb[i + 1]));
          m1 = fmaxf(fmaxf(m1, __uint_as_float(vb[i + 2])), __uint_as_float(vb[i + 3]));
          m2 = fmaxf(fmaxf(m2, __uint_as_float(vb[i + 4])), __uint_as_float(vb[i + 5]));
          m3 = fmaxf(fmaxf(m3, __uint_as_float(vb[i + 6])), __uint_as_float(vb[i + 7]));
        }
      }
      tmem_ld_wait(va);
    }
  } else {
    uint32_t v[64];
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
      for (int c = 0; c < 512; c += 64) {
        tmem_ld_32x64(base + c, v);
        tmem_wait_only();
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
          m0 = fmaxf(fmaxf(m0, __uint_as_float(v[i + 0])), __uint_as_float(v[i + 1]));
          m1 = fmaxf(fmaxf(m1, __uint_as_float(v[i + 2])), __uint_as_float(v[i + 3]));
          m2 = fmaxf(fmaxf(m2, __uint_as_float(v[i + 4])), __uint_as_float(v[i + 5]));
          m3 = fmaxf(fmaxf(m3, __uint_as_float(v[i + 6])), __uint_as_float(v[i + 7]));
        }
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  tc5_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc5_fence_after();
    tmem_dealloc(tmem_ptr, 512);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* sink;
  long long* cyc;
  cudaMalloc(&sink, (size_t)sms * 512 * 4);
  cudaMalloc(&cyc, sms * 8);
  const int iters = 2000;
  const char* names[4] = {"x32 wait-each + max", "x32 two in flight + max", "x64 wait-each + max", "x32 wait-each, no math"};
  printf("# TMEM -> register drain, one CTA per SM (%d SMs), 512 columns x %d passes per warp\n", sms, iters);
  for (int mode = 0; mode < 4; ++mode) {
    for (int warps = 4; warps <= 16; warps *= 2) {
      tmem_read<<<sms, warps * 32>>>(mode, 10, sink, cyc);
      cudaDeviceSynchronize();
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a);
      tmem_read<<<sms, warps * 32>>>(mode, iters, sink, cyc);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        return 1;
      }
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      long long c0;
      cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes_per_sm = (double)warps * iters * 512 * 32 * 4;
      printf("%-26s warps %2d : %8.3f ms  %7.1f B/clk/SM (SM 0: %lld clk)  %6.2f TB/s chip\n", names[mode], warps, ms,
             bytes_per_sm / (double)c0, c0, bytes_per_sm * sms / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
