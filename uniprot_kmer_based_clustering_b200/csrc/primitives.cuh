// primitives.cuh — hand-written device-wide exclusive scan and stable LSD radix sort.
// Utility passes only (prefix sums over the vocabulary / rows, final edge ordering); the
// hot kernels live in extract.cuh, index.cuh and pairs.cuh.
#pragma once
#include "common.cuh"

namespace kc {

// ----------------------------------------------------------------------------------------
// Exclusive scan.  in(i) -> uint64 addend, out(i, exclusive_prefix, addend).
// Three launches: per-tile reduce, single-block scan of the tile sums, per-tile apply.
// Tile = 256 threads x 16 rounds; each warp owns 512 consecutive items so every load is a
// coalesced 128 B line and the scan order is the index order.
// ----------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanTile = 4096;

__device__ __forceinline__ unsigned long long warp_scan_incl64(unsigned long long v) {
  const uint32_t l = lane_id();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(kFullMask, v, o);
    if (l >= (uint32_t)o) v += t;
  }
  return v;
}

template <class InF>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(InF in, uint64_t n,
                                                                   unsigned long long* tile_sums) {
  __shared__ unsigned long long ws[kScanThreads / 32];
  const uint32_t w = threadIdx.x >> 5, l = lane_id();
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + w * 512u;
  unsigned long long s = 0;
#pragma unroll 4
  for (int r = 0; r < 16; ++r) {
    const uint64_t i = base + r * 32 + l;
    if (i < n) s += in(i);
  }
  s = warp_sum64(s);
  if (l == 0) ws[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < kScanThreads / 32; ++i) t += ws[i];
    tile_sums[blockIdx.x] = t;
  }
}

// in-place exclusive scan of the tile sums by one block; writes the grand total
__global__ void __launch_bounds__(1024) scan_tilesums_kernel(unsigned long long* sums, uint32_t nb,
                                                              unsigned long long* total) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry;
  const uint32_t t = threadIdx.x, w = t >> 5, l = t & 31;
  if (t == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += 1024) {
    const uint32_t i = base + t;
    const unsigned long long v = i < nb ? sums[i] : 0ull;
    const unsigned long long incl = warp_scan_incl64(v);
    if (l == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      const unsigned long long x = wsum[l];
      const unsigned long long xs = warp_scan_incl64(x);
      wsum[l] = xs - x;
    }
    __syncthreads();
    const unsigned long long excl = incl - v + wsum[w] + carry;
    if (i < nb) sums[i] = excl;
    __syncthreads();
    if (t == 1023) carry = excl + v;
    __syncthreads();
  }
  if (t == 0) *total = carry;
}

template <class InF, class OutF>
__global__ void __launch_bounds__(kScanThreads)
    scan_apply_kernel(InF in, OutF out, uint64_t n, const unsigned long long* tile_offs) {
  __shared__ unsigned long long ws[kScanThreads / 32];
  const uint32_t w = threadIdx.x >> 5, l = lane_id();
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + w * 512u;
  unsigned long long s = 0;
#pragma unroll 4
  for (int r = 0; r < 16; ++r) {
    const uint64_t i = base + r * 32 + l;
    if (i < n) s += in(i);
  }
  s = warp_sum64(s);
  if (l == 0) ws[w] = s;
  __syncthreads();
  unsigned long long running = tile_offs[blockIdx.x];
  for (uint32_t i = 0; i < w; ++i) running += ws[i];
  for (int r = 0; r < 16; ++r) {
    const uint64_t i = base + r * 32 + l;
    const unsigned long long v = i < n ? (unsigned long long)in(i) : 0ull;
    const unsigned long long incl = warp_scan_incl64(v);
    if (i < n) out(i, running + incl - v, v);
    running += __shfl_sync(kFullMask, incl, 31);
  }
}

struct ScanScratch {
  unsigned long long* tile_sums = nullptr;  // capacity in tiles
  unsigned long long* total = nullptr;      // device scalar
  uint64_t cap_tiles = 0;
};

// returns number of launches (3); *total (device) holds the grand total afterwards
template <class InF, class OutF>
inline int exclusive_scan(InF in, OutF out, uint64_t n, const ScanScratch& sc, cudaStream_t st) {
  const uint32_t nb = (uint32_t)((n + kScanTile - 1) / kScanTile);
  if (nb == 0) {
    cudaMemsetAsync(sc.total, 0, 8, st);
    return 0;
  }
  scan_reduce_kernel<<<nb, kScanThreads, 0, st>>>(in, n, sc.tile_sums);
  scan_tilesums_kernel<<<1, 1024, 0, st>>>(sc.tile_sums, nb, sc.total);
  scan_apply_kernel<<<nb, kScanThreads, 0, st>>>(in, out, n, sc.tile_sums);
  return 3;
}

// ----------------------------------------------------------------------------------------
// Stable LSD radix sort of (u64 key, u64 value) pairs, 8 bits per pass.
// ----------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsTile = 2048;

__global__ void __launch_bounds__(kRsThreads)
    rs_hist_kernel(const unsigned long long* keys, uint64_t n, int shift, uint32_t* hist, uint32_t nb) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kRsTile;
#pragma unroll
  for (int j = 0; j < kRsTile / kRsThreads; ++j) {
    const uint64_t i = base + j * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kRsThreads)
    rs_scatter_kernel(const unsigned long long* keys_in, const unsigned long long* vals_in,
                      unsigned long long* keys_out, unsigned long long* vals_out, uint64_t n, int shift,
                      const uint32_t* hist_scanned, uint32_t nb) {
  __shared__ uint32_t cnt[8][256];
  __shared__ uint32_t base[8][256];
  const uint32_t t = threadIdx.x, w = t >> 5, l = t & 31;
  for (int i = t; i < 8 * 256; i += kRsThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const uint64_t tile = (uint64_t)blockIdx.x * kRsTile + w * 256u;
  unsigned long long k[8];
  uint32_t rank[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint64_t i = tile + r * 32 + l;
    const bool valid = i < n;
    k[r] = valid ? keys_in[i] : ~0ull;
    const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(kFullMask, valid ? d : 256u + l);
    const uint32_t rin = __popc(peers & lanemask_lt());
    rank[r] = valid ? cnt[w][d] + rin : 0u;
    __syncwarp();
    if (valid && rin == 0) cnt[w][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {
    uint32_t running = hist_scanned[(uint64_t)t * nb + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) {
      base[ww][t] = running;
      running += cnt[ww][t];
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint64_t i = tile + r * 32 + l;
    if (i < n) {
      const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
      const uint32_t pos = base[w][d] + rank[r];
      keys_out[pos] = k[r];
      vals_out[pos] = vals_in[i];
    }
  }
}

struct U32In {
  const uint32_t* p;
  __device__ unsigned long long operator()(uint64_t i) const { return p[i]; }
};
struct U32ExclOut {
  uint32_t* p;
  __device__ void operator()(uint64_t i, unsigned long long excl, unsigned long long) const {
    p[i] = (uint32_t)excl;
  }
};

// Sorts n pairs by the key bits listed in `passes` (each entry = shift of an 8-bit digit,
// least significant first).  Result ends in (keys_a, vals_a) if the number of passes is even,
// else in (keys_b, vals_b); returns 1 in *in_b in the latter case.  hist holds 256*nb u32.
inline int radix_sort_pairs(unsigned long long* keys_a, unsigned long long* vals_a,
                            unsigned long long* keys_b, unsigned long long* vals_b, uint64_t n,
                            const int* passes, int n_passes, uint32_t* hist, const ScanScratch& sc,
                            cudaStream_t st, int* in_b) {
  int launches = 0;
  const uint32_t nb = (uint32_t)((n + kRsTile - 1) / kRsTile);
  *in_b = 0;
  if (nb == 0) return 0;
  for (int p = 0; p < n_passes; ++p) {
    rs_hist_kernel<<<nb, kRsThreads, 0, st>>>(keys_a, n, passes[p], hist, nb);
    launches += 1 + exclusive_scan(U32In{hist}, U32ExclOut{hist}, 256ull * nb, sc, st);
    rs_scatter_kernel<<<nb, kRsThreads, 0, st>>>(keys_a, vals_a, keys_b, vals_b, n, passes[p], hist, nb);
    launches += 1;
    unsigned long long* tk = keys_a;
    keys_a = keys_b;
    keys_b = tk;
    unsigned long long* tv = vals_a;
    vals_a = vals_b;
    vals_b = tv;
    *in_b ^= 1;
  }
  return launches;
}

}  // namespace kc
