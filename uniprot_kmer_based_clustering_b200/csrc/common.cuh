// common.cuh — shared device helpers for the k-mer clustering engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kc {

constexpr uint32_t kFullMask = 0xFFFFFFFFu;
constexpr uint32_t kSentinel = 0xFFFFFFFFu;  // never a packed k-mer (21^7 < 2^31) nor a protein
constexpr int kNumSM = 148;                  // B200: 2 dies x 74 SMs

// Rows are owned in blocks of 64 consecutive rows of the pair order ("bins": the entry bins of the
// partitioned index, bucket.cuh).  Sharded build / scoring: bin_owner[row >> 6] names the rank that
// owns the row; null = everything is this engine's.
constexpr uint32_t kBinRowsLog = 6;
constexpr uint32_t kBinRows = 1u << kBinRowsLog;
struct RowOwner {
  const uint8_t* bin_owner;
  uint32_t me;
  __device__ __forceinline__ bool mine(uint32_t row) const {
    return !bin_owner || bin_owner[row >> kBinRowsLog] == me;
  }
};

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

// first index i in [0, n) with a[i] >= key (a ascending); all 32 lanes call it, result uniform.
// 32 samples per round: two dependent loads for n <= 1024.
__device__ __forceinline__ uint32_t warp_lower_bound(const uint32_t* __restrict__ a, uint32_t n, uint32_t key) {
  const uint32_t lane = lane_id();
  uint32_t lo = 0, hi = n;
  while (hi - lo > 32u) {
    const uint32_t step = (hi - lo + 31u) >> 5;
    const uint32_t i = lo + lane * step;
    const bool lt = i < hi && a[i] < key;
    const uint32_t c = __popc(__ballot_sync(kFullMask, lt));
    const uint32_t nlo = c == 0 ? lo : lo + (c - 1u) * step + 1u;
    const uint32_t nhi = min(hi, lo + c * step);
    lo = nlo;
    hi = nhi;
  }
  const uint32_t i = lo + lane;
  const bool lt = i < hi && a[i] < key;
  return lo + __popc(__ballot_sync(kFullMask, lt));
}

// sub-warp groups of G lanes (G = 8 or 32): one group works on one row of a sliced pass
template <int G>
__device__ __forceinline__ uint32_t group_mask() {
  return G == 32 ? kFullMask : (((1u << (G & 31)) - 1u) << ((lane_id() / G) * G));
}
template <int G>
__device__ __forceinline__ unsigned long long group_sum64(unsigned long long v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(group_mask<G>(), v, o);
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

// ---- counter-based position sampler (kc_sample_position, include/kc_b200.h) ----------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t sample_key(unsigned long long seed, uint32_t protein) {
  return mix32((uint32_t)seed ^ mix32(protein ^ (uint32_t)(seed >> 32) ^ 0x9E3779B9u));
}
// bijection of [0, n): balanced 4-round Feistel over 2*half bits, cycle-walked into range
__host__ __device__ __forceinline__ uint32_t sample_perm(uint32_t key, uint32_t n, uint32_t x) {
  uint32_t half = 1;
  while ((1ull << (2 * half)) < (unsigned long long)n) ++half;
  const uint32_t mask = (1u << half) - 1u;
  uint32_t y = x;
  do {
    uint32_t L = y >> half, R = y & mask;
#pragma unroll
    for (uint32_t round = 0; round < 4; ++round) {
      const uint32_t t = L ^ (mix32(R ^ key ^ (round * 0x9E3779B9u)) & mask);
      L = R;
      R = t;
    }
    y = (L << half) | R;
  } while (y >= n);
  return y;
}

__host__ __device__ constexpr uint32_t pow21(int k) {
  uint32_t v = 1;
  for (int i = 0; i < k; ++i) v *= 21u;
  return v;
}

}  // namespace kc
