// extract.cuh — K1/K2/K3: residue encode, k-mer extraction, per-protein dedup, census marks.
//
// Replaces (reference root relative):
//   amino_acid_to_bits  src/protein.rs:49-54   -> 256-entry LUT in shared memory
//   create_five_mer     src/protein.rs:29-37   -> base-21 pack, first residue most significant
//   Protein::new        src/protein.rs:107-132 -> kmers_per_position_kernel (get_five_mers)
//   sort(); dedup()     src/main.rs:100-102    -> bitonic sort + adjacent-diff in shared memory
//   merge_sort census   src/main.rs:23-48,103-116 -> two presence bitmaps over the 21^k universe:
//                       2 bits of state per k-mer (see census_mark)
#pragma once
#include "common.cuh"

namespace kc {

__constant__ uint8_t c_residue_lut[256];

// ---- partitioned index (bucket.cuh): the extract kernels append every (distinct k-mer, row)
// incidence to the bucket its k-mer hashes to
// (the slot size is a run-time choice: 4096 records, or 8192 when a bucket overflowed; bucket.cuh)
__host__ __device__ __forceinline__ uint32_t kmer_bucket_hash(uint32_t kmer) { return kmer * 0x9E3779B1u; }

// where the extract kernels append the incidences (rec == nullptr: the universe-table build).
// Sharded build (multi-GPU, every rank owns some row blocks of the pair triangle): a rank
// only needs the k-mers its own rows hold, with ALL their holders.  `filter` is a bitmap over a
// hash of the k-mers of the rank's own rows (kmer_filter_build_kernel); the incidences of the
// other rows are appended only when their k-mer passes it (false positives only cost work).
struct BucketScatter {
  uint2* rec;          // n_buckets slots of `cap` {k-mer, row} records
  uint32_t* cursor;    // records appended per bucket (may pass cap: overflow, the build retries / falls back)
  uint32_t n_buckets;
  uint32_t cap;
  const uint32_t* filter;  // null: every incidence is kept
  uint32_t filter_mask;    // filter bits - 1 (power of two)
  RowOwner owner;
  __device__ __forceinline__ static uint32_t filter_hash(uint32_t kmer) {
    uint32_t h = kmer * 0x85EBCA6Bu;
    return h ^ (h >> 13);
  }
  // does the k-mer of a foreign row pass the filter of this rank's k-mers?
  __device__ __forceinline__ bool passes(uint32_t kmer) const {
    const uint32_t f = filter_hash(kmer) & filter_mask;
    return (__ldg(filter + (f >> 5)) >> (f & 31u)) & 1u;
  }
  // two steps so that a caller can keep several reservations (L2 atomics) in flight
  __device__ __forceinline__ unsigned long long reserve_kept(uint32_t kmer) const {  // (filter already applied)
    const uint32_t b = __umulhi(kmer_bucket_hash(kmer), n_buckets);
    const uint32_t pos = atomicAdd(&cursor[b], 1u);
    return pos < cap ? (unsigned long long)b * cap + pos : ~0ull;
  }
  __device__ __forceinline__ unsigned long long reserve(uint32_t kmer, uint32_t row) const {
    if (filter && !owner.mine(row) && !passes(kmer)) return ~0ull;
    return reserve_kept(kmer);
  }
  __device__ __forceinline__ void store(unsigned long long at, uint32_t kmer, uint32_t row) const {
    if (at != ~0ull) rec[at] = make_uint2(kmer, row);
  }
  __device__ __forceinline__ void put(uint32_t kmer, uint32_t row) const { store(reserve(kmer, row), kmer, row); }
};

// ---- census: two bits of state per k-mer, interleaved in one word (16 k-mers per u32):
// bit 2j = "held by >= 1 protein", bit 2j+1 = "held by >= 2 proteins".  The first holder sets
// the low bit; every later holder (a different protein, because the caller has deduplicated
// within the protein) finds it set and sets the high bit.  Both bits live in the same 32-byte
// sector, so one incidence costs one random sector.  The plain pre-read may be stale; it only
// skips work that is idempotent.
// PREREAD: worth it when k-mers are hot (small universe: many holders per k-mer, most marks are
// no-ops); with a sparse universe (k=7) it only adds a dependent round trip.
template <bool PREREAD>
__device__ __forceinline__ void census_mark(uint32_t kmer, uint32_t* __restrict__ seen) {
  const uint32_t w = kmer >> 4, sh = (kmer & 15u) * 2u;
  const uint32_t lo = 1u << sh, hi = 2u << sh;
  if (PREREAD && (seen[w] & hi)) return;
  const uint32_t old = atomicOr(&seen[w], lo);
  if ((old & lo) && !(old & hi)) atomicOr(&seen[w], hi);
}

// L2 blocking: the k-mer universe is cut into n_slices slices of 2^slice_shift k-mers.  While a
// row's sorted distinct k-mers are written out, ksplit[q * n + r] records where slice q starts
// in row r (ksplit[n_slices * n + r] = number of distinct k-mers), so that every later pass
// over one slice reads one contiguous run of each row without searching.
__device__ __forceinline__ void record_slice_starts(uint32_t* __restrict__ ksplit, uint32_t n, uint32_t r,
                                                    uint32_t q_from, uint32_t q_to, uint32_t j) {
  if (!ksplit) return;  // the partitioned index (bucket.cuh) does not slice the rows
  for (uint32_t q = q_from; q <= q_to; ++q) ksplit[(size_t)q * n + r] = j;
}

template <int K>
__device__ __forceinline__ uint32_t pack_kmer(const uint8_t* codes) {
  uint32_t v = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) v = v * 21u + codes[j];
  return v;
}

// stage `len` residues of one protein as codes into shared memory (word loads, any alignment)
__device__ __forceinline__ void stage_codes(const uint8_t* __restrict__ res, uint32_t pstart, uint32_t len,
                                            uint8_t* codes, const uint8_t* lut, uint32_t tid,
                                            uint32_t nthreads) {
  const uint32_t a0 = pstart & ~3u;
  const uint32_t words = (pstart + len - a0 + 3u) >> 2;
  const uint32_t* rw = reinterpret_cast<const uint32_t*>(res + a0);
  for (uint32_t wi = tid; wi < words; wi += nthreads) {
    const uint32_t x = ld_stream_u32(rw + wi);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int idx = (int)(a0 + 4u * wi + b) - (int)pstart;
      if (idx >= 0 && (uint32_t)idx < len) codes[idx] = lut[(x >> (8 * b)) & 255u];
    }
  }
}

// warp-synchronous bitonic sort of np2 (power of two) keys in shared memory
__device__ __forceinline__ void warp_bitonic(uint32_t* keys, uint32_t np2, uint32_t lane) {
  for (uint32_t size = 2; size <= np2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t t = lane; t < (np2 >> 1); t += 32) {
        const uint32_t i = ((t & ~(stride - 1u)) << 1) | (t & (stride - 1u));
        const uint32_t j = i | stride;
        const bool up = (i & size) == 0;
        const uint32_t a = keys[i], b = keys[j];
        if ((a > b) == up) {
          keys[i] = b;
          keys[j] = a;
        }
      }
      __syncwarp();
    }
  }
}

// In-register bitonic sort of 32*E keys held in shared memory (blocked layout: lane l owns
// keys[l*E .. l*E+E)).  Strides below E are compare-exchanges between a lane's own registers,
// strides of E and more are one shuffle per register (15 of the 45 stages at E = 16).  About
// 2.5x fewer instructions than the shared-memory network above and no bank conflicts.
template <int E>
__device__ __forceinline__ void warp_sort_blocked(uint32_t* keys, uint32_t lane) {
  uint32_t a[E];
#pragma unroll
  for (int j = 0; j < E; j += 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(keys + lane * E + j);
    a[j] = v.x;
    a[j + 1] = v.y;
    a[j + 2] = v.z;
    a[j + 3] = v.w;
  }
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
    const bool up_lane = ((lane * E) & size) == 0;  // direction when it depends on the lane
#pragma unroll
    for (int d = size >> 1; d > 0; d >>= 1) {
      if (d >= E) {
        const bool takemin = ((lane & (d / E)) == 0) == up_lane;
#pragma unroll
        for (int j = 0; j < E; ++j) {
          const uint32_t p = __shfl_xor_sync(kFullMask, a[j], d / E);
          a[j] = takemin ? min(a[j], p) : max(a[j], p);
        }
      } else {
#pragma unroll
        for (int j = 0; j < E; ++j) {
          if ((j & d) == 0) {
            const bool up = size < E ? ((j & size) == 0) : up_lane;
            const uint32_t lo = min(a[j], a[j | d]), hi = max(a[j], a[j | d]);
            a[j] = up ? lo : hi;
            a[j | d] = up ? hi : lo;
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < E; j += 4)
    *reinterpret_cast<uint4*>(keys + lane * E + j) = make_uint4(a[j], a[j + 1], a[j + 2], a[j + 3]);
}

// block-wide bitonic sort (keys may live in shared or global memory)
__device__ __forceinline__ void block_bitonic(uint32_t* keys, uint32_t np2) {
  for (uint32_t size = 2; size <= np2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        const uint32_t i = ((t & ~(stride - 1u)) << 1) | (t & (stride - 1u));
        const uint32_t j = i | stride;
        const bool up = (i & size) == 0;
        const uint32_t a = keys[i], b = keys[j];
        if ((a > b) == up) {
          keys[i] = b;
          keys[j] = a;
        }
      }
      __syncthreads();
    }
  }
}

constexpr uint32_t kWarpMaxPos = 1024;    // proteins with <= this many positions: one warp
constexpr uint32_t kBlockMaxPos = 32768;  // <= this many: one CTA, keys in shared memory
constexpr int kExtractWarps = 8;

// ---------------------------------------------------------------------------------------
// K2 (short proteins): one warp per protein.  Writes the protein's sorted distinct k-mers to
// pk[pstart .. pstart+ndist).
// ---------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kExtractWarps * 32)
    extract_dedup_warp_kernel(const uint8_t* __restrict__ res, const uint32_t* __restrict__ pstart,
                              const uint32_t* __restrict__ plen, uint32_t n, uint32_t* __restrict__ pk,
                              uint32_t* __restrict__ ndist, uint32_t slice_shift, uint32_t n_slices,
                              uint32_t* __restrict__ ksplit, uint32_t sample_every, unsigned long long sample_seed,
                              const uint32_t* __restrict__ orig_of, unsigned long long* __restrict__ n_incid,
                              BucketScatter scatter, uint32_t min_pos) {
  __shared__ uint8_t s_lut[256];
  __shared__ __align__(16) uint32_t s_keys[kExtractWarps][kWarpMaxPos];
  __shared__ __align__(4) uint8_t s_codes[kExtractWarps][kWarpMaxPos + 8];
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
  s_lut[threadIdx.x] = c_residue_lut[threadIdx.x];
  __syncthreads();
  uint32_t* keys = s_keys[w];
  uint8_t* codes = s_codes[w];
  unsigned long long incid = 0;
  const uint32_t nwarps = gridDim.x * kExtractWarps;
  for (uint32_t r = blockIdx.x * kExtractWarps + w; r < n; r += nwarps) {
    const uint32_t len = plen[r];
    if (len < (uint32_t)K) {
      if (lane == 0) ndist[r] = 0;
      if (ksplit)
        for (uint32_t q = lane; q <= n_slices; q += 32) ksplit[(size_t)q * n + r] = 0;
      continue;
    }
    const uint32_t npos_all = len - K + 1;
    if (npos_all > kWarpMaxPos) continue;  // handled by the block kernels
    if (npos_all <= min_pos) continue;     // handled by extract_scatter_warp_kernel
    // optional subsampling (Protein::new_with_rand_fivemers): floor(positions / d) distinct positions
    const uint32_t npos = sample_every > 1 ? npos_all / sample_every : npos_all;
    if (npos == 0) {
      if (lane == 0) ndist[r] = 0;
      if (ksplit)
        for (uint32_t q = lane; q <= n_slices; q += 32) ksplit[(size_t)q * n + r] = 0;
      continue;
    }
    const uint32_t ps = pstart[r];
    stage_codes(res, ps, len, codes, s_lut, lane, 32);
    __syncwarp();
    const uint32_t np2 = npos <= 128 ? 128u : next_pow2_u32(npos);
    if (sample_every > 1) {
      const uint32_t skey = sample_key(sample_seed, orig_of ? orig_of[r] : r);
      for (uint32_t i = lane; i < np2; i += 32)
        keys[i] = i < npos ? pack_kmer<K>(codes + sample_perm(skey, npos_all, i)) : kSentinel;
    } else {
      for (uint32_t i = lane; i < np2; i += 32) keys[i] = i < npos ? pack_kmer<K>(codes + i) : kSentinel;
    }
    __syncwarp();
    if (np2 == 128) warp_sort_blocked<4>(keys, lane);
    else if (np2 == 256) warp_sort_blocked<8>(keys, lane);
    else if (np2 == 512) warp_sort_blocked<16>(keys, lane);
    else warp_sort_blocked<32>(keys, lane);
    __syncwarp();
    uint32_t base = 0;
    for (uint32_t c = 0; c < npos; c += 128) {  // four 32-key chunks at a time: four bucket reservations in flight
      uint32_t v[4], j[4];
      bool first[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t i = c + 32 * u + lane;
        v[u] = i < npos ? keys[i] : kSentinel;
        first[u] = i < npos && (i == 0 || v[u] != keys[i - 1]);
        const uint32_t m = __ballot_sync(kFullMask, first[u]);
        j[u] = base + __popc(m & lanemask_lt());
        base += __popc(m);
      }
      unsigned long long at[4];
      if (scatter.rec) {
#pragma unroll
        for (int u = 0; u < 4; ++u) at[u] = first[u] ? scatter.reserve(v[u], r) : ~0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (first[u]) {
          const uint32_t i = c + 32 * u + lane;
          pk[ps + j[u]] = v[u];
          record_slice_starts(ksplit, n, r, i == 0 ? 0u : (keys[i - 1] >> slice_shift) + 1u, v[u] >> slice_shift, j[u]);
        }
      }
      if (scatter.rec) {
#pragma unroll
        for (int u = 0; u < 4; ++u) scatter.store(at[u], v[u], r);
      }
    }
    if (lane == 0) ndist[r] = base;
    if (ksplit)
      for (uint32_t q = (keys[npos - 1] >> slice_shift) + 1u + lane; q <= n_slices; q += 32)
        ksplit[(size_t)q * n + r] = base;
    incid += base;
    __syncwarp();
  }
  if (lane == 0 && incid) atomicAdd(n_incid, incid);
}

// ---------------------------------------------------------------------------------------
// K2 (long proteins): one CTA per listed protein.  GLOBAL_KEYS = keys in a global scratch
// slice (proteins beyond the shared-memory capacity), else in dynamic shared memory.
// ---------------------------------------------------------------------------------------
template <int K, bool GLOBAL_KEYS>
__global__ void __launch_bounds__(512)
    extract_dedup_block_kernel(const uint8_t* __restrict__ res, const uint32_t* __restrict__ pstart,
                               const uint32_t* __restrict__ plen, const uint32_t* __restrict__ list,
                               const unsigned long long* __restrict__ scratch_off,
                               uint32_t* __restrict__ scratch, uint32_t* __restrict__ pk,
                               uint32_t* __restrict__ ndist, uint32_t n, uint32_t slice_shift, uint32_t n_slices,
                               uint32_t* __restrict__ ksplit, uint32_t sample_every, unsigned long long sample_seed,
                               const uint32_t* __restrict__ orig_of, unsigned long long* __restrict__ n_incid,
                               BucketScatter scatter) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  __shared__ uint8_t s_lut[256];
  __shared__ uint32_t s_wcnt[32];
  if (threadIdx.x < 256) s_lut[threadIdx.x] = c_residue_lut[threadIdx.x];
  const uint32_t r = list[blockIdx.x];
  const uint32_t len = plen[r], ps = pstart[r];
  const uint32_t npos_all = len - K + 1;
  const uint32_t npos = sample_every > 1 ? npos_all / sample_every : npos_all;  // >= 1: long proteins only
  const uint32_t skey = sample_key(sample_seed, orig_of ? orig_of[r] : r);
  const uint32_t np2 = next_pow2_u32(npos);
  uint32_t* keys;
  __syncthreads();
  if (GLOBAL_KEYS) {
    keys = scratch + scratch_off[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) {
      uint32_t v = kSentinel;
      if (i < npos) {
        const uint32_t pos = sample_every > 1 ? sample_perm(skey, npos_all, i) : i;
        v = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) v = v * 21u + s_lut[res[ps + pos + j]];
      }
      keys[i] = v;
    }
  } else {
    keys = reinterpret_cast<uint32_t*>(dyn_smem);
    uint8_t* codes = dyn_smem + (size_t)next_pow2_u32(npos_all) * 4;
    stage_codes(res, ps, len, codes, s_lut, threadIdx.x, blockDim.x);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x)
      keys[i] = i < npos ? pack_kmer<K>(codes + (sample_every > 1 ? sample_perm(skey, npos_all, i) : i)) : kSentinel;
  }
  __syncthreads();
  block_bitonic(keys, np2);
  // two-pass compaction: every warp owns a contiguous slice
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t slice = (((npos + nw - 1) / nw) + 31u) & ~31u;
  const uint32_t lo = min(npos, w * slice), hi = min(npos, lo + slice);
  uint32_t cnt = 0;
  for (uint32_t c = lo; c < hi; c += 32) {
    const uint32_t i = c + lane;
    const bool first = i < hi && (i == 0 || keys[i] != keys[i - 1]);
    cnt += __popc(__ballot_sync(kFullMask, first));
  }
  if (lane == 0) s_wcnt[w] = cnt;
  __syncthreads();
  uint32_t base = 0, total = 0;
  for (uint32_t i = 0; i < nw; ++i) {
    if (i < w) base += s_wcnt[i];
    total += s_wcnt[i];
  }
  for (uint32_t c = lo; c < hi; c += 32) {
    const uint32_t i = c + lane;
    const uint32_t v = i < hi ? keys[i] : kSentinel;
    const bool first = i < hi && (i == 0 || v != keys[i - 1]);
    const uint32_t m = __ballot_sync(kFullMask, first);
    if (first) {
      const uint32_t j = base + __popc(m & lanemask_lt());
      pk[ps + j] = v;
      record_slice_starts(ksplit, n, r, i == 0 ? 0u : (keys[i - 1] >> slice_shift) + 1u, v >> slice_shift, j);
      if (scatter.rec) scatter.put(v, r);
    }
    base += __popc(m);
  }
  if (ksplit)
    for (uint32_t q = (keys[npos - 1] >> slice_shift) + 1u + threadIdx.x; q <= n_slices; q += blockDim.x)
      ksplit[(size_t)q * n + r] = total;
  if (threadIdx.x == 0) {
    ndist[r] = total;
    atomicAdd(n_incid, (unsigned long long)total);
  }
}

// ---------------------------------------------------------------------------------------
// K2 for the partitioned index (bucket.cuh), proteins of <= kHashMaxPos positions: one warp per
// protein, dedup through a per-warp shared-memory hash set instead of a sort (the buckets do not
// need the row's k-mers in order), every newly seen k-mer appended to its bucket right away.
// Nothing but ndist is written per row: the sorted distinct k-mers (pk) are only needed by the
// canonical view, which runs extract_dedup_warp_kernel on demand.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kHashSlots = 1024;
constexpr uint32_t kHashMaxPos = 716;  // load factor <= 0.7
constexpr int kXsWarps = 8;

template <int K>
__global__ void __launch_bounds__(kXsWarps * 32)
    extract_scatter_warp_kernel(const uint8_t* __restrict__ res, const uint32_t* __restrict__ pstart,
                                const uint32_t* __restrict__ plen, uint32_t row_begin, uint32_t n,
                                uint32_t* __restrict__ ndist,
                                uint32_t sample_every, unsigned long long sample_seed,
                                const uint32_t* __restrict__ orig_of, unsigned long long* __restrict__ n_incid,
                                BucketScatter scatter) {
  __shared__ uint8_t s_lut[256];
  __shared__ __align__(16) uint32_t s_tab[kXsWarps][kHashSlots];
  __shared__ __align__(4) uint8_t s_codes[kXsWarps][kHashMaxPos + K + 8];
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
  s_lut[threadIdx.x] = c_residue_lut[threadIdx.x];
  __syncthreads();
  uint32_t* tab = s_tab[w];
  uint8_t* codes = s_codes[w];
  unsigned long long incid = 0;
  const uint32_t nwarps = gridDim.x * kXsWarps;
  for (uint32_t r = row_begin + blockIdx.x * kXsWarps + w; r < n; r += nwarps) {  // rows [row_begin, n)
    const uint32_t len = plen[r];
    if (len < (uint32_t)K) {
      if (lane == 0) ndist[r] = 0;
      continue;
    }
    const uint32_t npos_all = len - K + 1;
    if (npos_all > kHashMaxPos) continue;  // sorted by the warp / block kernels
    const uint32_t npos = sample_every > 1 ? npos_all / sample_every : npos_all;
    if (npos == 0) {
      if (lane == 0) ndist[r] = 0;
      continue;
    }
    const uint32_t ps = pstart[r];
#pragma unroll
    for (int i = 0; i < (int)(kHashSlots / 128); ++i)
      *reinterpret_cast<uint4*>(tab + i * 128 + lane * 4) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
    stage_codes(res, ps, len, codes, s_lut, lane, 32);
    __syncwarp();
    const uint32_t skey = sample_every > 1 ? sample_key(sample_seed, orig_of ? orig_of[r] : r) : 0u;
    uint32_t fresh = 0;
    // Sharded build, a row of another rank: most of its k-mers fail the filter of this rank's k-mers (81 % at
    // 8 ranks), so the filter is probed FIRST (four L2 loads in flight) and only the survivors enter the hash
    // set; ndist of such a row counts its kept k-mers (it only sizes the row's entry capacity).
    const bool foreign = scatter.rec && scatter.filter && !scatter.owner.mine(r);
    for (uint32_t c = 0; c < npos; c += 128) {  // four k-mers per lane: four bucket reservations in flight
      uint32_t v[4];
      bool first[4], act[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t i = c + 32 * u + lane;
        first[u] = false;
        act[u] = i < npos;
        v[u] = act[u] ? pack_kmer<K>(codes + (sample_every > 1 ? sample_perm(skey, npos_all, i) : i)) : 0u;
      }
      if (foreign) {
#pragma unroll
        for (int u = 0; u < 4; ++u) act[u] = act[u] && scatter.passes(v[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (act[u]) {
          uint32_t h = (v[u] * 2654435761u) >> 22;
          for (;;) {
            const uint32_t cur = tab[h];
            if (cur == v[u]) break;
            if (cur == kSentinel) {
              const uint32_t old = atomicCAS(&tab[h], kSentinel, v[u]);
              if (old == kSentinel) {
                first[u] = true;
                break;
              }
              if (old == v[u]) break;
            }
            h = (h + 1u) & (kHashSlots - 1u);
          }
        }
      }
      unsigned long long at[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        at[u] = !first[u] ? ~0ull : foreign ? scatter.reserve_kept(v[u]) : scatter.reserve(v[u], r);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        scatter.store(at[u], v[u], r);
        fresh += first[u];
      }
    }
    fresh = warp_sum(fresh);
    if (lane == 0) ndist[r] = fresh;
    incid += fresh;
    __syncwarp();
  }
  if (lane == 0 && incid) atomicAdd(n_incid, incid);
}

// The same for proteins of kHashMaxPos+1 .. kCtaHashMaxPos positions: one CTA (4 warps) per listed
// protein, a 4096-slot hash set in shared memory.  (Longer ones go through the sorting block
// kernels: with skewed lengths up to 2 000 residues those were 24 % of the step, profiles/.)
constexpr uint32_t kCtaHashSlots = 4096;
constexpr uint32_t kCtaHashMaxPos = 2800;  // load factor <= 0.7
constexpr int kXcThreads = 128;

template <int K>
__global__ void __launch_bounds__(kXcThreads)
    extract_scatter_cta_kernel(const uint8_t* __restrict__ res, const uint32_t* __restrict__ pstart,
                               const uint32_t* __restrict__ plen, const uint32_t* __restrict__ list, uint32_t n_list,
                               uint32_t* __restrict__ ndist, uint32_t sample_every, unsigned long long sample_seed,
                               const uint32_t* __restrict__ orig_of, unsigned long long* __restrict__ n_incid,
                               BucketScatter scatter) {
  __shared__ uint8_t s_lut[256];
  __shared__ __align__(16) uint32_t s_tab[kCtaHashSlots];
  __shared__ __align__(4) uint8_t s_codes[kCtaHashMaxPos + K + 8];
  __shared__ uint32_t s_fresh;
  const uint32_t tid = threadIdx.x;
  for (uint32_t i = tid; i < 256; i += kXcThreads) s_lut[i] = c_residue_lut[i];
  unsigned long long incid = 0;
  for (uint32_t li = blockIdx.x; li < n_list; li += gridDim.x) {
    const uint32_t r = list[li];
    const uint32_t len = plen[r], ps = pstart[r];
    const uint32_t npos_all = len - K + 1;
    const uint32_t npos = sample_every > 1 ? npos_all / sample_every : npos_all;
    __syncthreads();  // the previous row is done with the table and the codes
    for (uint32_t i = tid * 4; i < kCtaHashSlots; i += kXcThreads * 4)
      *reinterpret_cast<uint4*>(s_tab + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
    if (tid == 0) s_fresh = 0;
    stage_codes(res, ps, len, s_codes, s_lut, tid, kXcThreads);
    __syncthreads();
    const uint32_t skey = sample_every > 1 ? sample_key(sample_seed, orig_of ? orig_of[r] : r) : 0u;
    uint32_t fresh = 0;
    for (uint32_t c = 0; c < npos; c += 4 * kXcThreads) {  // four k-mers per thread: four reservations in flight
      uint32_t v[4];
      bool first[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t i = c + u * kXcThreads + tid;
        first[u] = false;
        v[u] = 0;
        if (i < npos) {
          v[u] = pack_kmer<K>(s_codes + (sample_every > 1 ? sample_perm(skey, npos_all, i) : i));
          uint32_t h = (v[u] * 2654435761u) >> 20;
          for (;;) {
            const uint32_t cur = s_tab[h];
            if (cur == v[u]) break;
            if (cur == kSentinel) {
              const uint32_t old = atomicCAS(&s_tab[h], kSentinel, v[u]);
              if (old == kSentinel) {
                first[u] = true;
                break;
              }
              if (old == v[u]) break;
            }
            h = (h + 1u) & (kCtaHashSlots - 1u);
          }
        }
      }
      unsigned long long at[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) at[u] = first[u] ? scatter.reserve(v[u], r) : ~0ull;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        scatter.store(at[u], v[u], r);
        fresh += first[u];
      }
    }
    fresh = warp_sum(fresh);
    if (lane_id() == 0 && fresh) atomicAdd(&s_fresh, fresh);
    __syncthreads();
    if (tid == 0) {
      ndist[r] = s_fresh;
      incid += s_fresh;
    }
  }
  if (tid == 0 && incid) atomicAdd(n_incid, incid);
}

// Sharded build: mark (a hash of) every k-mer of this rank's rows in the filter bitmap.
// One warp per row, any length; duplicates are harmless.
template <int K>
__global__ void __launch_bounds__(256)
    kmer_filter_build_kernel(const uint8_t* __restrict__ res, const uint32_t* __restrict__ pstart,
                             const uint32_t* __restrict__ plen, uint32_t row_begin, uint32_t n, RowOwner owner,
                             uint32_t sample_every, unsigned long long sample_seed,
                             const uint32_t* __restrict__ orig_of, uint32_t* __restrict__ filter,
                             uint32_t filter_mask) {
  __shared__ uint8_t s_lut[256];
  s_lut[threadIdx.x] = c_residue_lut[threadIdx.x];
  __syncthreads();
  const uint32_t lane = lane_id();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = row_begin + gw; r < n; r += nw) {  // rows [row_begin, n)
    if (!owner.mine(r)) continue;
    const uint32_t len = plen[r], ps = pstart[r];
    if (len < (uint32_t)K) continue;
    const uint32_t npos_all = len - K + 1;
    const uint32_t npos = sample_every > 1 ? npos_all / sample_every : npos_all;
    const uint32_t skey = sample_every > 1 ? sample_key(sample_seed, orig_of ? orig_of[r] : r) : 0u;
    for (uint32_t i = lane; i < npos; i += 32) {
      const uint32_t pos = sample_every > 1 ? sample_perm(skey, npos_all, i) : i;
      uint32_t v = 0;
#pragma unroll
      for (int j = 0; j < K; ++j) v = v * 21u + s_lut[res[ps + pos + j]];
      const uint32_t f = BucketScatter::filter_hash(v) & filter_mask;
      atomicOr(&filter[f >> 5], 1u << (f & 31u));
    }
  }
}

// ---------------------------------------------------------------------------------------
// K3: census pass over one k-mer slice.  L2 blocking: the census state of the whole 21^k
// universe (450 MB at k=7) does not fit the 126 MB L2, so the random read-modify-writes are
// done slice by slice; lo[r]..hi[r] is the run of row r that falls into the slice.
// G lanes work on one row.
// ---------------------------------------------------------------------------------------
template <int G, bool PREREAD>
__global__ void __launch_bounds__(256)
    census_pass_kernel(const uint32_t* __restrict__ pk, const uint32_t* __restrict__ pstart,
                       const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, uint32_t n,
                       uint32_t* __restrict__ seen) {
  const uint32_t gl = lane_id() % G;
  const uint32_t gg = (blockIdx.x * blockDim.x + threadIdx.x) / G, ng = (gridDim.x * blockDim.x) / G;
  for (uint32_t r = gg; r < n; r += ng) {
    const uint32_t i0 = lo[r], i1 = hi[r], ps = pstart[r];
    for (uint32_t i = i0 + gl; i < i1; i += G) census_mark<PREREAD>(pk[ps + i], seen);
  }
}

// ---------------------------------------------------------------------------------------
// K1: Protein::new / get_five_mers — one packed k-mer per start position, proteins back to
// back in INPUT order.  A CTA stages a 4096-residue tile plus a (K-1)-residue halo in shared
// memory with 16-byte loads; the tile's protein boundaries are staged next to it.
// ---------------------------------------------------------------------------------------
constexpr int kTileRes = 4096;
constexpr int kTileMaxProt = 1024;

template <int K>
__global__ void __launch_bounds__(256)
    kmers_per_position_kernel(const uint8_t* __restrict__ res, uint64_t n_res,
                              const unsigned long long* __restrict__ off,
                              const unsigned long long* __restrict__ kpos_off, uint32_t n_prot,
                              uint32_t* __restrict__ out) {
  __shared__ uint8_t s_lut[256];
  __shared__ __align__(16) uint8_t s_codes[kTileRes + 16];
  __shared__ unsigned long long s_off[kTileMaxProt + 1];
  __shared__ uint32_t s_p0, s_np;
  const uint32_t t = threadIdx.x;
  s_lut[t] = c_residue_lut[t];
  const uint64_t tile0 = (uint64_t)blockIdx.x * kTileRes;
  if (t == 0) {
    // first protein whose end is beyond tile0: upper_bound(off, tile0) - 1
    uint32_t lo = 0, hi = n_prot;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (off[mid + 1] <= tile0) lo = mid + 1; else hi = mid;
    }
    s_p0 = lo;
    // number of proteins starting before the tile end
    const uint64_t tend = tile0 + kTileRes;
    uint32_t lo2 = lo, hi2 = n_prot;
    while (lo2 < hi2) {
      const uint32_t mid = (lo2 + hi2) >> 1;
      if (off[mid] < tend) lo2 = mid + 1; else hi2 = mid;
    }
    s_np = lo2 - lo;
  }
  __syncthreads();
  {
    const uint4 x = ld_stream_u32x4(reinterpret_cast<const uint4*>(res + tile0) + t);  // buffer is padded
    const uint32_t wv[4] = {x.x, x.y, x.z, x.w};
    uint32_t packed[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      packed[q] = (uint32_t)s_lut[wv[q] & 255u] | ((uint32_t)s_lut[(wv[q] >> 8) & 255u] << 8) |
                  ((uint32_t)s_lut[(wv[q] >> 16) & 255u] << 16) | ((uint32_t)s_lut[wv[q] >> 24] << 24);
    }
    reinterpret_cast<uint4*>(s_codes)[t] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    if (t < K - 1) {
      const uint64_t g = tile0 + kTileRes + t;
      s_codes[kTileRes + t] = g < n_res ? s_lut[res[g]] : (uint8_t)20;
    }
  }
  const uint32_t p0 = s_p0, np = s_np;
  const bool local = np <= (uint32_t)kTileMaxProt;
  if (local)
    for (uint32_t i = t; i <= np; i += 256) s_off[i] = off[p0 + i];
  __syncthreads();
#pragma unroll 4
  for (int j = 0; j < kTileRes / 256; ++j) {
    const uint32_t li = j * 256 + t;
    const uint64_t g = tile0 + li;
    if (g >= n_res) break;
    // protein holding residue g
    uint32_t lo = 0, hi = np;
    if (local) {
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_off[mid + 1] <= g) lo = mid + 1; else hi = mid;
      }
    } else {
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (off[p0 + mid + 1] <= g) lo = mid + 1; else hi = mid;
      }
    }
    const uint64_t pbeg = local ? s_off[lo] : off[p0 + lo];
    const uint64_t pend = local ? s_off[lo + 1] : off[p0 + lo + 1];
    if (g + K <= pend) out[kpos_off[p0 + lo] + (g - pbeg)] = pack_kmer<K>(s_codes + li);
  }
}

// K1 in subsampling mode: the x-th sampled k-mer of every protein (input order), x ascending
template <int K>
__global__ void __launch_bounds__(256)
    kmers_sampled_kernel(const uint8_t* __restrict__ res, const unsigned long long* __restrict__ off,
                         const unsigned long long* __restrict__ kpos_off, uint32_t n_prot, uint32_t sample_every,
                         unsigned long long sample_seed, uint32_t* __restrict__ out) {
  __shared__ uint8_t s_lut[256];
  s_lut[threadIdx.x] = c_residue_lut[threadIdx.x];
  __syncthreads();
  const uint32_t lane = lane_id();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t p = gw; p < n_prot; p += nw) {
    const unsigned long long b = off[p], len = off[p + 1] - b;
    if (len < (unsigned long long)K) continue;
    const uint32_t npos_all = (uint32_t)(len - K + 1), m = npos_all / sample_every;
    const uint32_t skey = sample_key(sample_seed, p);
    for (uint32_t x = lane; x < m; x += 32) {
      const uint32_t pos = sample_perm(skey, npos_all, x);
      uint32_t v = 0;
#pragma unroll
      for (int j = 0; j < K; ++j) v = v * 21u + s_lut[res[b + pos + j]];
      out[kpos_off[p] + x] = v;
    }
  }
}

}  // namespace kc
