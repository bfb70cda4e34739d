// common.cuh — device-side building blocks shared by the sm_100a kernels.
//   * mbarrier + TMA bulk-copy (cp.async.bulk) wrappers (inline PTX)
//   * order-preserving (score, row) -> u64 keys: larger key == better result
//   * CTA-level bitonic sort of u64 keys in shared memory (descending)
//   * fp16 / bf16 -> fp32 dot-product helpers
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rs {

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin until the phase with `parity` completes.  A protocol bug must surface as a CUDA error, not
// as a hung GPU: after ~2^31 cycles (about a second) of waiting the kernel traps.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0u && clock64() - t0 > (1ll << 31)) __trap();
  }
}

// L2 eviction policy for data that is streamed exactly once.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`.
// dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// 128-bit global load for data read exactly once: read-only path, no L1 allocation
__device__ __forceinline__ uint4 ld_stream(const uint4* ptr) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(ptr));
  return r;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// barrier + OR-reduction of a predicate over the participating threads
__device__ __forceinline__ bool named_bar_or(int id, int nthreads, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %3, 0;\n\t"
      "barrier.cta.red.or.pred p, %1, %2, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(id), "r"(nthreads), "r"((uint32_t)pred)
      : "memory");
  return r != 0;
}

// ----------------------------------------------------------------------------- result keys
// key = (orderable(score) << 32) | (0xFFFFFFFF - row).  Larger key <=> higher score, and on
// equal score the LOWER row wins: exactly the (score desc, id asc) total order of the ABI.
// key 0 is reserved for "empty slot" (no finite/inf score maps to orderable 0: NaNs are
// canonicalised to -inf first).
__device__ __forceinline__ uint32_t f32_orderable(float s) {
  if (s != s) s = -INFINITY;
  s += 0.f;  // -0.0 -> +0.0: the two compare equal everywhere else (numpy / torch / Python sorts), so they must tie here
  uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float orderable_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(f32_orderable(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - row);
}
__device__ __forceinline__ float key_score(uint64_t key) { return orderable_f32(static_cast<uint32_t>(key >> 32)); }
__device__ __forceinline__ uint32_t key_row(uint64_t key) { return 0xFFFFFFFFu - static_cast<uint32_t>(key); }

// ----------------------------------------------------------------------------- bitonic sort
// Sort `n_pow2` u64 keys in shared memory, DESCENDING, with `nthreads` threads (a multiple of 32)
// that all call this function (tid in [0, nthreads)), synchronising on named barrier `bar_id`.
//
// Compare-exchange stages whose stride is < 64 stay inside a 64-key chunk, so a warp runs all of
// them on its chunks with __syncwarp only; the CTA barrier is needed just for the strides >= 64.
// For 512 keys that is ~10 CTA barriers instead of 45 — the sort sits on the critical path of
// every top-k compaction and of the end-of-kernel merge.
__device__ __forceinline__ void bitonic_ce(uint64_t* keys, int lo, int stride, int size) {
  const int hi = lo + stride;
  const bool desc = ((lo & size) == 0);
  const uint64_t a = keys[lo], b = keys[hi];
  if ((a < b) == desc) {
    keys[lo] = b;
    keys[hi] = a;
  }
}
// strides stride_hi, stride_hi/2, ..., 1 of the merge step `size`, on every 64-key chunk, warp-locally
__device__ __forceinline__ void bitonic_local(uint64_t* keys, int n, int size, int stride_hi, int warp, int lane,
                                              int nwarps) {
  const int chunk_n = n < 64 ? n : 64;
  for (int base = warp * 64; base < n; base += nwarps * 64) {
    for (int stride = stride_hi; stride > 0; stride >>= 1) {
      __syncwarp();
      if (lane < (chunk_n >> 1)) bitonic_ce(keys, base + 2 * lane - (lane & (stride - 1)), stride, size);
    }
    __syncwarp();
  }
}
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* keys, int n_pow2, int tid, int nthreads, int bar_id) {
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  named_bar_sync(bar_id, nthreads);
  const int local_max = n_pow2 < 64 ? n_pow2 : 64;
  for (int size = 2; size <= local_max; size <<= 1) bitonic_local(keys, n_pow2, size, size >> 1, warp, lane, nwarps);
  for (int size = 128; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride >= 64; stride >>= 1) {
      named_bar_sync(bar_id, nthreads);
      for (int i = tid; i < (n_pow2 >> 1); i += nthreads) bitonic_ce(keys, 2 * i - (i & (stride - 1)), stride, size);
    }
    named_bar_sync(bar_id, nthreads);
    bitonic_local(keys, n_pow2, size, 32, warp, lane, nwarps);
  }
  named_bar_sync(bar_id, nthreads);
}

// ----------------------------------------------------------------------------- dot helpers
template <typename T>
struct Cvt;
template <>
struct Cvt<__half> {
  // two packed halves -> two floats
  static __device__ __forceinline__ float2 unpack(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
  }
  static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
};
template <>
struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float2 unpack(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
  }
  static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
};

// acc += <8 packed 16-bit elements in v, 8 packed 16-bit elements in q>, fp32 accumulation.
// sm_100 has a mixed-precision FMA (PTX fma.rn.f32.f16 / .bf16, SASS FHFMA with .H0/.H1 operand
// selects): a 16-bit x 16-bit product is exact in fp32, so this is bit-identical to converting both
// operands to fp32 first and costs one instruction per element instead of three.
template <typename T>
__device__ __forceinline__ float fma16(uint16_t a, uint16_t b, float c);
template <>
__device__ __forceinline__ float fma16<__half>(uint16_t a, uint16_t b, float c) {
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(a), "h"(b));
  return c;
}
template <>
__device__ __forceinline__ float fma16<__nv_bfloat16>(uint16_t a, uint16_t b, float c) {
  asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(c) : "h"(a), "h"(b));
  return c;
}
template <typename T>
__device__ __forceinline__ float dot2(uint32_t a, uint32_t b, float acc) {
  uint16_t al, ah, bl, bh;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(al), "=h"(ah) : "r"(a));
  asm("mov.b32 {%0, %1}, %2;" : "=h"(bl), "=h"(bh) : "r"(b));
  acc = fma16<T>(al, bl, acc);
  return fma16<T>(ah, bh, acc);
}
template <typename T>
__device__ __forceinline__ float dot8(const uint4& v, const uint4& q, float acc) {
  acc = dot2<T>(v.x, q.x, acc);
  acc = dot2<T>(v.y, q.y, acc);
  acc = dot2<T>(v.z, q.z, acc);
  return dot2<T>(v.w, q.w, acc);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

}  // namespace rs
