// common.cuh — shared device helpers for the k-mer clustering engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kc {

constexpr uint32_t kFullMask = 0xFFFFFFFFu;
constexpr uint32_t kSentinel = 0xFFFFFFFFu;  // never a packed k-mer (21^7 < 2^31) nor a protein
constexpr int kNumSM = 148;                  // B200: 2 dies x 74 SMs

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// streaming (read-once) loads: keep them out of L1 so the tables that are re-used stay there
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ld_stream_u32x2(const uint2* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u32x4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// inclusive warp scan
__device__ __forceinline__ uint32_t warp_scan_incl(uint32_t v) {
  const uint32_t l = lane_id();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(kFullMask, v, o);
    if (l >= (uint32_t)o) v += t;
  }
  return v;
}

__host__ __device__ __forceinline__ uint32_t next_pow2_u32(uint32_t v) {
  if (v <= 1) return 1;
  --v;
  v |= v >> 1;
  v |= v >> 2;
  v |= v >> 4;
  v |= v >> 8;
  v |= v >> 16;
  return v + 1;
}

__host__ __device__ constexpr uint32_t pow21(int k) {
  uint32_t v = 1;
  for (int i = 0; i < k; ++i) v *= 21u;
  return v;
}

}  // namespace kc
