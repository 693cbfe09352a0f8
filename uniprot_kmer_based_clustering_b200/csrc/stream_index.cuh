// stream_index.cuh — K1..K5 as a STREAMING, STABLE two-level partition of (k-mer, row) records,
// followed by one shared-memory pass per bucket.  The default index build.
//
// Replaces (reference root relative):
//   Protein::new / create_five_mer           src/protein.rs:29-37,107-132   (sx_l1_* : rolling base-21 pack)
//   sort(); dedup() + merge_sort census      src/main.rs:23-48,100-116      (duplicates of a row meet in the
//                                                                            bucket pass, where they are adjacent)
//   split unique/repeated + Mphf::new x2     src/main.rs:127-147
//   remove_unique_five_mers + modify_hash_five_mer + kmer_freq   src/protein.rs:151-174, src/main.rs:182-193
//   times_kmer_visited / triangular layout   src/graph/vertex.rs:92-136
//
// Why it is shaped like this (profiles/r1_q_full_1m.md): the previous build deduplicated every protein in
// a per-warp hash set and appended each incidence through a per-bucket atomic cursor (344 M L2 atomics and
// 344 M scattered 8-byte stores: 9.9 ms, nothing saturated), then hash-grouped and rank-sorted (O(f^2)) every
// bucket.  Here nothing is appended through atomics and nothing is rank-sorted:
//   1. The residue stream is walked in the pair order, tile by tile (8 192 residues per CTA step, 16-byte
//      loads, rolling k-mer per thread).  Every position yields a record {k-mer, row}; the record's bucket is
//      the top b1+b2 bits of a bijective hash of the k-mer.
//   2. Level 1 and level 2 are STABLE counting-sort passes (count, scan, scatter) with the tile staged in
//      shared memory and ranked by warp ballots (no shared-memory atomics in the scatter), written out as
//      runs of consecutive records per digit.  Stability keeps every bucket sorted by row.  Level 1 runs in
//      PASSES (one per upload chunk while the stream is still crossing PCIe, else one): see SxPlan.
//   3. One WARP per bucket (~330 records; one CTA for the few larger ones): stable LSD radix sort of the bucket
//      on the REMAINING hash bits in shared memory.  Because the hash is a bijection, equal keys are equal
//      k-mers; because the passes are stable, the holders of a k-mer come out ascending and the duplicates
//      of a row are adjacent.  Census, ids, postings, suffixes and bin-local runs then fall out of ballots and
//      carried counters; no hash table, no O(f^2).
//   4. A bucket that does not fit shared memory (a k-mer with thousands of holders) takes the same steps
//      through global scratch with one CTA (sx_huge_kernel): no fallback build, no retry.
// The outputs (postings, vocabulary, entry bins, run records) are the ones bucket.cuh documents; the entry
// bins are split by rows_finalize_kernel as before.
#pragma once
#include "bucket.cuh"
#include "common.cuh"
#include "extract.cuh"
#include "index.cuh"

namespace kc {

__host__ __device__ __forceinline__ uint32_t sx_hash(uint32_t kmer) { return kmer * 0x9E3779B1u; }  // odd: a bijection of u32

constexpr uint32_t kNoDigit = 0xFFFFFFFFu;
// Tile = 8 192 records per CTA step in both levels.  Level 1: 512 threads x 16 positions, two CTAs per SM
// (its phases are barrier-separated: a second CTA fills the gaps).  Level 2: 1 024 threads x 8 records, one
// CTA per SM (16 records per thread do not fit 64 registers next to their digits, peers and ranks; measured:
// level 1 3.9 vs 5.0 ms, level 2 4.2 vs 3.7 ms for the two shapes, profiles/r2_history.md).
constexpr uint32_t kSxTile = 8192;
constexpr int kL1Threads = 512, kL1V = 16;
constexpr int kL2Threads = 1024, kL2V = 8;

// Level 1 runs in PASSES over consecutive tile ranges of the residue stream (one pass when the stream is
// resident; one per upload chunk when it is still crossing PCIe, so that the level-1 sort of chunk u runs
// while chunk u + 1 arrives).  Every pass is a complete counting sort of its own records into its own region
// of the level-1 output (pass-major, regions back to back); a level-1 partition is then the concatenation of
// its segments over the passes, which is again the stream order: level 2 reads it through a segment table.
constexpr int kSxMaxPass = 8;
struct SxPlan {
  uint32_t b1, b2;           // digit bits of the two levels (<= 10 each)
  uint32_t r;                // remaining hash bits: the in-bucket sort key
  uint32_t g1;               // level-1 chunks = CTAs of the level-1 kernels (per pass)
  uint32_t c2;               // level-2 chunks per level-1 partition
  uint32_t ballots;          // 1: warp peers by one ballot per digit bit, 0: by match.any (A/B switch)
  uint32_t n_pass;           // level-1 passes (1 .. kSxMaxPass)
  uint32_t pass_tile[kSxMaxPass + 1];  // pass u = tiles [pass_tile[u], pass_tile[u + 1])
  __host__ __device__ uint32_t d1() const { return 1u << b1; }
  __host__ __device__ uint32_t d2() const { return 1u << b2; }
  __host__ __device__ uint32_t n_buckets() const { return 1u << (b1 + b2); }
  // level-1 histogram of pass u: h1[u * h1_stride() + d * g1 + chunk]; the slot behind the last one holds the
  // end of the pass's region (= the start of the next pass's)
  __host__ __device__ uint32_t h1_stride() const { return d1() * g1 + 1u; }
  __host__ __device__ uint32_t pass_tiles_per_chunk(uint32_t u) const {
    return (pass_tile[u + 1] - pass_tile[u] + g1 - 1u) / g1;
  }
};

// dynamic shared memory of the two scatter kernels for D digits:
// staged records | warp counters (u16; the level-1 kernel first keeps the tile's residue codes there) |
// dstart[D + 1] | goff[D] | lut[256] | scan scratch
__host__ __device__ constexpr size_t sx_cnt_bytes(uint32_t D, uint32_t warps) {
  return (size_t)warps * D * 2 > (size_t)kSxTile + 64 ? (size_t)warps * D * 2 : (size_t)kSxTile + 64;
}
__host__ __device__ constexpr size_t sx_scatter_smem(uint32_t D, uint32_t warps) {
  return (size_t)kSxTile * 8 + sx_cnt_bytes(D, warps) + ((size_t)D + 4) * 4 + (size_t)D * 4 + 256 + 64 * 4;
}

// block-wide exclusive scan of one u32 per thread (THREADS <= 1024); returns the exclusive prefix,
// *total = block sum.  wsum: u32[33].  Barriers inside; wsum may be reused after the call returns.
template <int THREADS>
__device__ __forceinline__ uint32_t sx_block_scan(uint32_t v, uint32_t* __restrict__ wsum, uint32_t* total) {
  constexpr int WARPS = THREADS / 32;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t incl = warp_scan_incl(v);
  __syncthreads();  // wsum free
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t x = lane < (uint32_t)WARPS ? wsum[lane] : 0u;
    const uint32_t xs = warp_scan_incl(x);
    __syncwarp();
    wsum[lane] = xs - x;
    if (lane == 31) wsum[32] = xs;
  }
  __syncthreads();
  *total = wsum[32];
  return incl - v + wsum[warp];
}

// ---------------------------------------------------------------------------------------
// Stable rank of one tile.  Items are warp-striped: item j of lane l of warp w is tile element
// w * (V*32) + j * 32 + l, and that element order is the order the ranks preserve.
// digit[j] = digit of the item or kNoDigit.  Returns pos[j] = index of the item in the tile
// sorted by digit; dstart[d] = first index of digit d, dstart[D] = valid items.
// cnt: u16[WARPS][D]; dstart: u32[D + 1]; wsum: u32[33].  All threads call it (barriers inside).
// ---------------------------------------------------------------------------------------
template <int THREADS, int V>
__device__ __forceinline__ void sx_tile_rank(const uint32_t (&digit)[V], uint32_t (&pos)[V], uint32_t logd,
                                             uint32_t rounds, bool ballots, uint16_t* __restrict__ cnt,
                                             uint32_t* __restrict__ dstart, uint32_t* __restrict__ wsum) {
  constexpr int WARPS = THREADS / 32;
  const uint32_t D = 1u << logd;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(cnt);
    for (uint32_t i = tid; i < (uint32_t)WARPS * D / 2; i += THREADS) z[i] = 0;
  }
  // lanes of the warp that hold the same digit, for every round up front: the matches are independent and
  // pipeline; the counter updates below are the only serial chain
  uint32_t peers[V];
  if (ballots) {
#pragma unroll
    for (int j = 0; j < V; ++j) peers[j] = (uint32_t)j < rounds ? __ballot_sync(kFullMask, digit[j] != kNoDigit) : 0u;
    for (uint32_t bit = 0; bit < logd; ++bit) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if ((uint32_t)j < rounds) {
          const bool one = (digit[j] >> bit) & 1u;
          const uint32_t bm = __ballot_sync(kFullMask, one);
          peers[j] &= one ? bm : ~bm;
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j)
      peers[j] = (uint32_t)j < rounds
                     ? __match_any_sync(kFullMask, digit[j] != kNoDigit ? digit[j] : (0x80000000u | lane))
                     : 0u;
  }
  __syncthreads();
  uint16_t* cw = cnt + warp * D;
  // (one shared-memory atomic per distinct digit and round + a shuffle, which lets the rounds pipeline, was
  // measured slower than this load / store / __syncwarp chain: profiles/r2_history.md)
  const uint32_t lt_mask = lanemask_lt();
#pragma unroll
  for (int j = 0; j < V; ++j) {
    if ((uint32_t)j < rounds) {  // uniform
      const uint32_t d = digit[j];
      const bool valid = d != kNoDigit;
      const uint32_t rin = __popc(peers[j] & lt_mask);
      const uint32_t before = valid ? cw[d] : 0u;
      __syncwarp();
      if (valid && rin == 0) cw[d] = (uint16_t)(before + __popc(peers[j]));
      __syncwarp();
      pos[j] = before + rin;
    }
  }
  __syncthreads();
  // per digit: exclusive prefix over the warps (in place) and the digit's total
  {  // two digits (one 32-bit word of 16-bit counters) per step: the packed halves never carry (<= tile size)
    uint32_t* c32 = reinterpret_cast<uint32_t*>(cnt);
    const uint32_t W = D >> 1;
    for (uint32_t wd = tid; wd < W; wd += THREADS) {
      uint32_t run = 0;
#pragma unroll 4
      for (int w = 0; w < WARPS; ++w) {
        const uint32_t t = c32[(uint32_t)w * W + wd];
        c32[(uint32_t)w * W + wd] = run;
        run += t;
      }
      dstart[2u * wd] = run & 0xFFFFu;
      dstart[2u * wd + 1u] = run >> 16;
    }
  }
  __syncthreads();
  // exclusive scan of the totals: thread t owns the digits [t * dpt, (t + 1) * dpt)
  const uint32_t dpt = (D + THREADS - 1) / THREADS;
  uint32_t mine = 0;
  for (uint32_t q = 0; q < dpt; ++q) {
    const uint32_t d = tid * dpt + q;
    if (d < D) mine += dstart[d];
  }
  uint32_t total;
  uint32_t run = sx_block_scan<THREADS>(mine, wsum, &total);
  for (uint32_t q = 0; q < dpt; ++q) {
    const uint32_t d = tid * dpt + q;
    if (d < D) {
      const uint32_t t = dstart[d];
      dstart[d] = run;
      run += t;
    }
  }
  if (tid == 0) dstart[D] = total;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < V; ++j)
    if ((uint32_t)j < rounds && digit[j] != kNoDigit) pos[j] += dstart[digit[j]] + cw[digit[j]];
}

// ---------------------------------------------------------------------------------------
// Level 1: the records of one tile of the residue stream (pair order).
// soff[n + 1]: row r = residues [soff[r], soff[r + 1]) of `res`.  Thread t of the CTA owns the 16
// consecutive positions t0 + 16 t ..; it rolls the k-mer along them and hands every position to
// `sink(j, kmer | kSentinel, row)`.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sx_pow21(int e) {
  uint32_t v = 1;
  for (int i = 0; i < e; ++i) v *= 21u;
  return v;
}

// largest r in [lo, hi] with soff[r] <= x (soff[lo] <= x is given)
__device__ __forceinline__ uint32_t sx_row_of(const uint32_t* __restrict__ soff, uint32_t lo, uint32_t hi, uint32_t x) {
  while (lo < hi) {
    const uint32_t mid = (lo + hi + 1u) >> 1;
    if (__ldg(soff + mid) <= x) lo = mid; else hi = mid - 1u;
  }
  return lo;
}

template <int K, int P, class Sink>
__device__ __forceinline__ void sx_tile_positions(const uint8_t* __restrict__ res, uint32_t R,
                                                  const uint32_t* __restrict__ soff, uint32_t t0,
                                                  uint32_t row_lo, uint32_t row_hi,
                                                  uint8_t* __restrict__ s_codes, const uint8_t* __restrict__ s_lut,
                                                  Sink sink) {
  static_assert(P == 8 || P == 16, "8 or 16 positions per thread");
  constexpr int NW = P / 4;  // 32-bit words of residues per thread
  const uint32_t tid = threadIdx.x;
  // stage the tile's residue codes (+ halo): one 8- or 16-byte load per thread, `res` is padded with zeros
  {
    uint32_t w[NW];
    if (P == 16) {
      const uint4 x = *reinterpret_cast<const uint4*>(res + (size_t)t0 + tid * 16u);
      w[0] = x.x, w[1] = x.y, w[NW - 2] = x.z, w[NW - 1] = x.w;
    } else {
      const uint2 x = *reinterpret_cast<const uint2*>(res + (size_t)t0 + tid * 8u);
      w[0] = x.x, w[1] = x.y;
    }
    uint32_t o[NW];
#pragma unroll
    for (int q = 0; q < NW; ++q)
      o[q] = (uint32_t)s_lut[w[q] & 255u] | ((uint32_t)s_lut[(w[q] >> 8) & 255u] << 8) |
             ((uint32_t)s_lut[(w[q] >> 16) & 255u] << 16) | ((uint32_t)s_lut[w[q] >> 24] << 24);
    if (P == 16) *reinterpret_cast<uint4*>(s_codes + tid * 16u) = make_uint4(o[0], o[1], o[NW - 2], o[NW - 1]);
    else *reinterpret_cast<uint2*>(s_codes + tid * 8u) = make_uint2(o[0], o[1]);
    if (tid < 16) s_codes[kSxTile + tid] = s_lut[res[(size_t)t0 + kSxTile + tid]];
  }
  __syncthreads();
  const uint32_t i0 = t0 + tid * (uint32_t)P;
  uint32_t row = 0, next = 0;
  if (i0 < R) {
    row = sx_row_of(soff, row_lo, row_hi, i0);
    next = __ldg(soff + row + 1);
  }
  uint32_t c[P + K - 1];
  {
    uint32_t w[NW + 2];
#pragma unroll
    for (int q = 0; q < NW + 2; q += 2) {
      const uint2 a = *reinterpret_cast<const uint2*>(s_codes + tid * (uint32_t)P + 4u * q);
      w[q] = a.x;
      w[q + 1] = a.y;
    }
#pragma unroll
    for (int q = 0; q < P + K - 1; ++q) c[q] = (w[q >> 2] >> (8 * (q & 3))) & 255u;
  }
  const uint32_t top = sx_pow21(K - 1);
  uint32_t km = 0;
#pragma unroll
  for (int q = 0; q < K; ++q) km = km * 21u + c[q];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const uint32_t i = i0 + (uint32_t)j;
    bool valid = i < R;
    if (valid) {
      while (i >= next) {  // next row (empty rows: several steps)
        ++row;
        next = __ldg(soff + row + 1);
      }
      valid = i + K <= next;
    }
    sink(j, valid ? km : kSentinel, row);
    if (j < P - 1) km = (km - c[j] * top) * 21u + c[j + K];
  }
}

// Sharded build (multi-GPU, owner computes: every rank owns row blocks of the pair triangle and needs the
// k-mers its own rows hold, with ALL their holders): the records of foreign rows are kept only when their
// k-mer passes the filter of the rank's own k-mers (kmer_filter_build_kernel; false positives only cost work).
struct SxKeep {
  uint32_t* filter;        // null: keep everything.  A Bloom filter blocked to one 32-bit word: two bits per
  uint32_t word_mask;      // k-mer, one probe (one 32-byte sector) per foreign position
  RowOwner owner;
  uint16_t* keepmask;      // 16 positions per entry: the count pass records its decisions, the scatter pass
                           // reads them back instead of probing again
  __device__ __forceinline__ static uint32_t word_of(uint32_t kmer, uint32_t word_mask, uint32_t* need) {
    uint32_t h = kmer * 0x85EBCA6Bu;
    h ^= h >> 15;
    h *= 0xC2B2AE35u;
    h ^= h >> 13;
    *need = (1u << (h & 31u)) | (1u << ((h >> 5) & 31u));
    return (h >> 10) & word_mask;
  }
  __device__ __forceinline__ bool operator()(uint32_t kmer, uint32_t row) const {
    if (!filter || owner.mine(row)) return true;
    uint32_t need;
    const uint32_t w = word_of(kmer, word_mask, &need);
    return (__ldg(filter + w) & need) == need;
  }
  __device__ __forceinline__ void add(uint32_t kmer) const {
    uint32_t need;
    const uint32_t w = word_of(kmer, word_mask, &need);
    if ((filter[w] & need) != need) atomicOr(&filter[w], need);
  }
};

// record slot inside a warp's segment (32 * V records) of the staging buffer: position p = V * lane + j of the
// blocked phase, read back as p = 32 * j' + lane' by the striped phase; the XOR keeps both free of bank
// conflicts (8-byte slots: 16 per bank row; V = 8 or 16)
__device__ __forceinline__ uint32_t sx_swz(uint32_t p) { return p ^ ((p >> 4) & 15u); }

// sharded build: the filter of the k-mers the rank's own rows hold.  Tiles [tile_lo, tile_hi) of the stream.
template <int K>
__global__ void __launch_bounds__(kL1Threads, 2)
    sx_filter_build_kernel(const uint8_t* __restrict__ res, uint32_t R, const uint32_t* __restrict__ soff,
                           const uint32_t* __restrict__ tile_row, uint32_t tile_lo, uint32_t tile_hi, SxKeep keep) {
  __shared__ __align__(16) uint8_t s_codes[kSxTile + 64];
  __shared__ uint8_t s_lut[256];
  if (threadIdx.x < 256) s_lut[threadIdx.x] = c_residue_lut[threadIdx.x];
  __syncthreads();
  for (uint32_t t = tile_lo + blockIdx.x; t < tile_hi; t += gridDim.x) {
    const unsigned long long t0 = (unsigned long long)t * kSxTile;
    if (t0 >= R) break;
    sx_tile_positions<K, kL1V>(res, R, soff, (uint32_t)t0, tile_row[t], tile_row[t + 1u], s_codes, s_lut,
                               [&](int, uint32_t km, uint32_t row) {
                                 if (km != kSentinel && keep.owner.mine(row)) keep.add(km);
                               });
    __syncthreads();
  }
}

// level-1 count of one pass: digit histogram of every chunk.  hist[pass * stride + d * g1 + chunk]
template <int K, bool SHARDED>
__global__ void __launch_bounds__(kL1Threads, 2)
    sx_l1_count_kernel(const uint8_t* __restrict__ res, uint32_t R, const uint32_t* __restrict__ soff,
                       const uint32_t* __restrict__ tile_row, SxPlan plan, uint32_t pass, SxKeep keep,
                       uint32_t* __restrict__ hist) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(dyn_smem);  // [D1]
  uint8_t* s_codes = dyn_smem + (size_t)plan.d1() * 4;       // [tile + 64]
  uint8_t* s_lut = s_codes + kSxTile + 64;
  const uint32_t tid = threadIdx.x, D = plan.d1(), sh = 32u - plan.b1;
  for (uint32_t d = tid; d < D; d += kL1Threads) s_hist[d] = 0;
  if (tid < 256) s_lut[tid] = c_residue_lut[tid];
  __syncthreads();
  const uint32_t chunk = blockIdx.x;
  const uint32_t tpc = plan.pass_tiles_per_chunk(pass), tile_end = plan.pass_tile[pass + 1];
  const uint32_t tile_lo = plan.pass_tile[pass] + chunk * tpc;
  hist += (size_t)pass * plan.h1_stride();
  for (uint32_t t = 0; t < tpc && tile_lo + t < tile_end; ++t) {
    const unsigned long long t0 = (unsigned long long)(tile_lo + t) * kSxTile;
    if (t0 >= R) break;
    uint32_t kept = 0;
    sx_tile_positions<K, kL1V>(res, R, soff, (uint32_t)t0, tile_row[tile_lo + t], tile_row[tile_lo + t + 1u], s_codes, s_lut,
                         [&](int j, uint32_t km, uint32_t row) {
                           if (km != kSentinel && (!SHARDED || keep(km, row))) {
                             atomicAdd(&s_hist[sx_hash(km) >> sh], 1u);
                             if (SHARDED) kept |= 1u << j;
                           }
                         });
    if (SHARDED) keep.keepmask[(size_t)(tile_lo + t) * kL1Threads + tid] = (uint16_t)kept;
    __syncthreads();
  }
  for (uint32_t d = tid; d < D; d += kL1Threads) hist[(size_t)d * plan.g1 + chunk] = s_hist[d];
}

// level-1 scatter of one pass: hist_scanned[pass * stride + d * g1 + chunk] = where this chunk's records of
// digit d start in `out`
template <int K, bool SHARDED>
__global__ void __launch_bounds__(kL1Threads, 2)
    sx_l1_scatter_kernel(const uint8_t* __restrict__ res, uint32_t R, const uint32_t* __restrict__ soff,
                         const uint32_t* __restrict__ tile_row, SxPlan plan, uint32_t pass, SxKeep keep,
                         const uint32_t* __restrict__ hist_scanned, uint2* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  const uint32_t D = plan.d1(), sh = 32u - plan.b1;
  uint2* s_buf = reinterpret_cast<uint2*>(dyn_smem);                          // [tile]
  uint16_t* s_cnt = reinterpret_cast<uint16_t*>(dyn_smem + (size_t)kSxTile * 8);  // [warps][D]  (codes first)
  uint8_t* s_codes = reinterpret_cast<uint8_t*>(s_cnt);
  uint32_t* s_dstart = reinterpret_cast<uint32_t*>(dyn_smem + (size_t)kSxTile * 8 + sx_cnt_bytes(D, kL1Threads / 32));  // [D + 1]
  uint32_t* s_goff = s_dstart + D + 4;                                       // [D]
  uint8_t* s_lut = reinterpret_cast<uint8_t*>(s_goff + D);                   // [256]
  uint32_t* s_wsum = reinterpret_cast<uint32_t*>(s_lut + 256);               // [33]
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t chunk = blockIdx.x;
  hist_scanned += (size_t)pass * plan.h1_stride();
  for (uint32_t d = tid; d < D; d += kL1Threads) s_goff[d] = hist_scanned[(size_t)d * plan.g1 + chunk];
  if (tid < 256) s_lut[tid] = c_residue_lut[tid];
  __syncthreads();
  const uint32_t tpc = plan.pass_tiles_per_chunk(pass), tile_end = plan.pass_tile[pass + 1];
  const uint32_t tile_lo = plan.pass_tile[pass] + chunk * tpc;
  uint2* seg = s_buf + warp * (32u * kL1V);
  for (uint32_t t = 0; t < tpc && tile_lo + t < tile_end; ++t) {
    const unsigned long long t0 = (unsigned long long)(tile_lo + t) * kSxTile;
    if (t0 >= R) break;
    const uint32_t kept = SHARDED ? keep.keepmask[(size_t)(tile_lo + t) * kL1Threads + tid] : 0xFFFFu;
    sx_tile_positions<K, kL1V>(res, R, soff, (uint32_t)t0, tile_row[tile_lo + t], tile_row[tile_lo + t + 1u], s_codes, s_lut,
                         [&](int j, uint32_t km, uint32_t row) {
                           if (SHARDED && !((kept >> j) & 1u)) km = kSentinel;
                           seg[sx_swz(lane * (uint32_t)kL1V + (uint32_t)j)] = make_uint2(km, row);
                         });
    __syncthreads();  // the codes live where the rank counters are zeroed next
    // Sharded build: most records of a tile are dropped (foreign rows whose k-mer is not in the filter).
    // Compact the warp's segment first (in place, order kept) so that the ranking below only runs the
    // rounds that hold records; the other builds lose 2 % of the positions (row ends): not worth the pass.
    uint32_t rounds = kL1V;
    if (SHARDED) {
      uint32_t cntw = 0;
#pragma unroll
      for (int j = 0; j < kL1V; ++j) {
        const uint2 v = seg[sx_swz((uint32_t)j * 32u + lane)];
        const uint32_t m = __ballot_sync(kFullMask, v.x != kSentinel);
        __syncwarp();
        if (v.x != kSentinel) seg[cntw + __popc(m & lanemask_lt())] = v;
        cntw += __popc(m);
        __syncwarp();
      }
      rounds = (cntw + 31u) >> 5;
      // (records beyond cntw in the last round are stale copies: mark them)
      if (rounds * 32u > cntw && cntw + lane < rounds * 32u) seg[cntw + lane] = make_uint2(kSentinel, 0u);
      __syncwarp();
    }
    uint32_t km[kL1V], rw[kL1V], digit[kL1V], pos[kL1V];
#pragma unroll
    for (int j = 0; j < kL1V; ++j) {
      uint2 v = make_uint2(kSentinel, 0u);
      if (!SHARDED || (uint32_t)j < rounds) v = seg[SHARDED ? (uint32_t)j * 32u + lane : sx_swz((uint32_t)j * 32u + lane)];
      km[j] = v.x;
      rw[j] = v.y;
      digit[j] = v.x == kSentinel ? kNoDigit : sx_hash(v.x) >> sh;
    }
    sx_tile_rank<kL1Threads, kL1V>(digit, pos, plan.b1, rounds, plan.ballots != 0, s_cnt, s_dstart, s_wsum);
#pragma unroll
    for (int j = 0; j < kL1V; ++j)
      if (digit[j] != kNoDigit) s_buf[pos[j]] = make_uint2(km[j], rw[j]);
    __syncthreads();
    const uint32_t n_valid = s_dstart[D];
    for (uint32_t i = tid; i < n_valid; i += kL1Threads) {
      const uint2 v = s_buf[i];
      const uint32_t d = sx_hash(v.x) >> sh;
      out[(size_t)s_goff[d] + (i - s_dstart[d])] = v;
    }
    __syncthreads();
    for (uint32_t d = tid; d < D; d += kL1Threads) s_goff[d] += s_dstart[d + 1] - s_dstart[d];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Level 2: level-1 partition p = the concatenation, over the level-1 passes u, of the segments
// [h1[u * stride + p * g1], h1[u * stride + (p + 1) * g1]) of `in` (one segment when the stream was resident).
// That VIRTUAL record range is cut into c2 chunks; one CTA per chunk.
// hist[(p * d2 + d) * c2 + chunk]: the global exclusive scan of that array is the final position
// of every (partition, digit, chunk) run, and its stride-c2 samples are the bucket starts.
// ---------------------------------------------------------------------------------------
// segment table of partition p in shared memory: s_vs[u] = virtual start of segment u (s_vs[n_pass] = the
// partition's records), s_pb[u] = where segment u starts in `in`.  Returns this CTA's virtual range.
__device__ __forceinline__ void sx_l2_chunk(const uint32_t* __restrict__ h1, const SxPlan& plan, uint32_t p,
                                            uint32_t c, uint32_t* __restrict__ s_vs, uint32_t* __restrict__ s_pb,
                                            uint32_t& beg, uint32_t& end) {
  if (threadIdx.x == 0) {
    uint32_t lo[kSxMaxPass], hi[kSxMaxPass];
#pragma unroll
    for (int u = 0; u < kSxMaxPass; ++u) {
      if ((uint32_t)u < plan.n_pass) {
        const uint32_t* h = h1 + (size_t)u * plan.h1_stride();
        lo[u] = h[(size_t)p * plan.g1];
        hi[u] = h[(size_t)(p + 1u) * plan.g1];
      }
    }
    uint32_t run = 0;
#pragma unroll
    for (int u = 0; u < kSxMaxPass; ++u) {
      if ((uint32_t)u < plan.n_pass) {
        s_vs[u] = run;
        s_pb[u] = lo[u];
        run += hi[u] - lo[u];
      }
    }
    s_vs[plan.n_pass] = run;
  }
  __syncthreads();
  const uint32_t total = s_vs[plan.n_pass];
  const uint32_t per = (total + plan.c2 - 1u) / plan.c2;
  beg = min(total, c * per);
  end = min(total, beg + per);
}
// segment of virtual record i
__device__ __forceinline__ uint32_t sx_l2_seg_of(const uint32_t* __restrict__ s_vs, uint32_t n_pass, uint32_t i) {
  uint32_t s = 0;
#pragma unroll
  for (int u = 1; u < kSxMaxPass; ++u) s += ((uint32_t)u < n_pass && i >= s_vs[u]) ? 1u : 0u;
  return s;
}

__global__ void __launch_bounds__(kL2Threads)
    sx_l2_count_kernel(const uint2* __restrict__ in, const uint32_t* __restrict__ h1, SxPlan plan,
                       uint32_t* __restrict__ hist) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  __shared__ uint32_t s_vs[kSxMaxPass + 1], s_pb[kSxMaxPass];
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(dyn_smem);
  const uint32_t tid = threadIdx.x, D = plan.d2(), sh = 32u - plan.b1 - plan.b2, mask = D - 1u;
  const uint32_t p = blockIdx.x / plan.c2, c = blockIdx.x % plan.c2;
  for (uint32_t d = tid; d < D; d += kL2Threads) s_hist[d] = 0;
  uint32_t beg, end;
  sx_l2_chunk(h1, plan, p, c, s_vs, s_pb, beg, end);  // (barrier inside)
  if (plan.n_pass == 1) {  // one segment: the plain strided loop (its loads batch)
    const uint2* seg = in + s_pb[0];
    for (uint32_t i = beg + tid; i < end; i += kL2Threads)
      atomicAdd(&s_hist[(sx_hash(ld_stream_u32(&seg[i].x)) >> sh) & mask], 1u);
  } else {
    for (uint32_t u = 0; u < plan.n_pass; ++u) {
      const uint32_t a = max(beg, s_vs[u]), b = min(end, s_vs[u + 1]);
      if (a >= b) continue;
      const uint2* seg = in + s_pb[u];
      for (uint32_t i = a - s_vs[u] + tid; i < b - s_vs[u]; i += kL2Threads)
        atomicAdd(&s_hist[(sx_hash(ld_stream_u32(&seg[i].x)) >> sh) & mask], 1u);
    }
  }
  __syncthreads();
  for (uint32_t d = tid; d < D; d += kL2Threads) hist[((size_t)p * D + d) * plan.c2 + c] = s_hist[d];
}

__global__ void __launch_bounds__(kL2Threads, 1)
    sx_l2_scatter_kernel(const uint2* __restrict__ in, const uint32_t* __restrict__ h1, SxPlan plan,
                         const uint32_t* __restrict__ hist_scanned, uint2* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  __shared__ uint32_t s_vs[kSxMaxPass + 1], s_pb[kSxMaxPass];
  const uint32_t D = plan.d2(), sh = 32u - plan.b1 - plan.b2, mask = D - 1u;
  uint2* s_buf = reinterpret_cast<uint2*>(dyn_smem);
  uint16_t* s_cnt = reinterpret_cast<uint16_t*>(dyn_smem + (size_t)kSxTile * 8);
  uint32_t* s_dstart = reinterpret_cast<uint32_t*>(dyn_smem + (size_t)kSxTile * 8 + sx_cnt_bytes(D, kL2Threads / 32));
  uint32_t* s_goff = s_dstart + D + 4;
  uint32_t* s_wsum = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s_goff + D) + 256);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t p = blockIdx.x / plan.c2, c = blockIdx.x % plan.c2;
  uint32_t beg, end;
  sx_l2_chunk(h1, plan, p, c, s_vs, s_pb, beg, end);
  if (beg >= end) return;
  for (uint32_t d = tid; d < D; d += kL2Threads) s_goff[d] = hist_scanned[((size_t)p * D + d) * plan.c2 + c];
  __syncthreads();
  for (uint32_t t0 = beg; t0 < end; t0 += kSxTile) {
    uint32_t km[kL2V], rw[kL2V], digit[kL2V], pos[kL2V];
    // the warp's 32 * kL2V virtual records usually lie in one segment: one offset for all of them
    const uint32_t w0 = t0 + warp * (32u * kL2V);
    uint32_t delta = s_pb[0];
    bool one_seg = true;
    if (plan.n_pass > 1 && w0 < end) {
      const uint32_t sa = sx_l2_seg_of(s_vs, plan.n_pass, w0);
      const uint32_t sb = sx_l2_seg_of(s_vs, plan.n_pass, min(end, w0 + 32u * kL2V) - 1u);
      one_seg = sa == sb;
      delta = s_pb[sa] - s_vs[sa];
    }
    uint32_t pi[kL2V];  // physical index (mod 2^32) of the thread's records
#pragma unroll
    for (int j = 0; j < kL2V; ++j) pi[j] = w0 + (uint32_t)j * 32u + lane + delta;
    if (!one_seg) {  // (rare, warp-uniform: the range straddles a segment border)
#pragma unroll
      for (int j = 0; j < kL2V; ++j) {
        const uint32_t i = w0 + (uint32_t)j * 32u + lane;
        if (i < end) {
          const uint32_t sg = sx_l2_seg_of(s_vs, plan.n_pass, i);
          pi[j] = s_pb[sg] + (i - s_vs[sg]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kL2V; ++j) {  // (no control flow between the loads: they are issued back to back)
      const uint32_t i = w0 + (uint32_t)j * 32u + lane;
      uint2 v = make_uint2(kSentinel, 0u);
      if (i < end) v = ld_stream_u32x2(in + pi[j]);
      km[j] = v.x;
      rw[j] = v.y;
      digit[j] = v.x == kSentinel ? kNoDigit : (sx_hash(v.x) >> sh) & mask;
    }
    sx_tile_rank<kL2Threads, kL2V>(digit, pos, plan.b2, kL2V, plan.ballots != 0, s_cnt, s_dstart, s_wsum);
#pragma unroll
    for (int j = 0; j < kL2V; ++j)
      if (digit[j] != kNoDigit) s_buf[pos[j]] = make_uint2(km[j], rw[j]);
    __syncthreads();
    const uint32_t n_valid = s_dstart[D];
    for (uint32_t i = tid; i < n_valid; i += kL2Threads) {
      const uint2 v = s_buf[i];
      const uint32_t d = (sx_hash(v.x) >> sh) & mask;
      out[(size_t)s_goff[d] + (i - s_dstart[d])] = v;
    }
    __syncthreads();
    for (uint32_t d = tid; d < D; d += kL2Threads) s_goff[d] += s_dstart[d + 1] - s_dstart[d];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// The bucket pass.  bucket b = records [bstart(b), bstart(b + 1)) of `rec`, sorted by row (stable
// partition), bstart(b) = h2[b * c2], h2[n_buckets * c2] = all records.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sx_self_score(const uint8_t* __restrict__ ss3, uint32_t kmer, int k);

struct SxBucketArgs {
  const uint2* rec;           // the level-2 output
  uint2* scratch;             // the level-1 output (free again): sx_huge_kernel's ping-pong space
  const uint32_t* h2;         // scanned level-2 histogram
  SxPlan plan;
  const uint32_t* first_after;  // cross-class mode, else null
  int k;
  uint32_t* col;              // postings
  uint4* entries;             // entry bins (bin_region)
  const uint32_t* rowcap_prefix;
  uint32_t* bin_cursor;
  uint32_t* vocab;
  uint32_t* freq;
  uint8_t* selfscore;
  const uint8_t* ss3;         // BLOSUM62 self-score of every residue triple (21^3 entries)
  BucketGlobals* g;
  unsigned long long* n_incid;  // (row, k-mer) incidences after the per-row dedup
  uint32_t* mid_list;         // buckets beyond a warp's shared memory: one CTA each (sx_bucket_kernel)
  uint32_t* mid_cnt;
  uint32_t* huge_list;        // buckets beyond a CTA's shared memory: global-memory path (sx_huge_kernel)
  uint32_t* huge_cnt;
  uint32_t mid_cap;           // records the CTA kernel takes (4096; 512 in the tests)
  RowOwner owner;             // sharded build: the counters are owned by the rank of a k-mer's FIRST holder,
                              // the run records and the pair work by the rank of the row (null: all mine)
};
// Postings and ids are placed by CAPACITY, with no reservation: bucket b = records [beg, end) writes its
// postings at col[beg ..) (it has at most end - beg of them) and numbers its repeated k-mers beg / 2 + local
// id (at most (end - beg) / 2 of them, and floor(beg / 2) + floor(n / 2) <= floor(end / 2)).  Both spaces have
// holes; every consumer goes through entries / ids that carry absolute positions.  The build is thereby
// deterministic, and a million buckets do not queue up on two atomic counters.

// What one (row, repeated k-mer) incidence needs once the holders of its k-mer lie sorted in `rows`
// (rows[c] for c in [c, ge): the holder itself and the holders after it): its posting, its entry
// {row | self-score << 24, id, postings suffix of the holders to pair with through the hash kernels}, and,
// for the first holder of a run of >= 2 holders inside one 64-row bin, the run record for pairs_tile_kernel.
// The suffix starts behind the row's bin (bin-local partners go to the tiles) and, in cross-class mode,
// not before the first holder of a later class block.  Same conventions as bucket_build_kernel.
struct SxEmit {
  uint4 ent;
  unsigned long long rmask;  // != 0: the run record's mask
  uint32_t len;              // partners left to the hash kernels
};
template <class Rows>
__device__ __forceinline__ SxEmit sx_emit(const Rows& rows, uint32_t c, uint32_t ge, bool head, uint32_t row,
                                          uint32_t ss, uint32_t id, uint32_t post /* of c */,
                                          const uint32_t* __restrict__ first_after) {
  SxEmit o;
  const uint32_t bin = row >> kBinRowsLog;
  unsigned long long mask = 1ull << (row & (kBinRows - 1u));
  uint32_t a = c + 1u;
  while (a < ge) {
    const uint32_t x = rows(a);
    if ((x >> kBinRowsLog) != bin) break;
    mask |= 1ull << (x & (kBinRows - 1u));
    ++a;
  }
  const bool leader = a > c + 1u && (head || (rows(c - 1u) >> kBinRowsLog) != bin);
  if (first_after) {
    const uint32_t target = first_after[row];
    uint32_t lo = a, hi = ge;  // first holder >= target in [a, ge)
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (rows(mid) < target) lo = mid + 1u; else hi = mid;
    }
    a = lo;
  }
  o.len = ge - a;
  const uint2 sf = o.len == 1u ? make_uint2(rows(a), kSentinel) : make_uint2(post + (a - c), post + (ge - c));
  o.ent = make_uint4(row | (ss << 24), id, sf.x, sf.y);
  o.rmask = leader ? mask : 0ull;
  return o;
}

// ---------------------------------------------------------------------------------------
// One WARP per bucket (the common case: the plan aims at ~330 records per bucket, a warp takes WCAP).
// No block barrier anywhere: sort, dedup, group scans and emission are warp-synchronous over the
// warp's private slice of shared memory; round u = elements [32 u, 32 u + 32).
//   sort    stable LSD on the remaining hash bits, <= 7 bits per pass: ranks by match.any in a first
//           sweep (kept packed in registers), placement in a second
//   dedup   adjacent duplicates (one row, one k-mer) dropped, in place
//   back    next k-mer head / next run head behind every record (rounds descending, carried)
//   emit    rounds ascending with carried prefix counts: postings, vocabulary, entries, run records
// Buckets beyond WCAP go to the CTA kernel's list, beyond its capacity to the global-memory path.
// ---------------------------------------------------------------------------------------
constexpr int kWbWarps = 8;
template <uint32_t WCAP>
__host__ __device__ constexpr size_t sx_wb_warp_bytes() { return (size_t)WCAP * 16 + 128 * 2 + 16; }
constexpr size_t kWbTableBytes = 448;  // BLOSUM62 self-score of every residue pair (441 entries)

// self-score of a k-mer from the CTA's table of residue pairs: k = 7 -> 2 + 2 + 2 + 1 digits, k = 5 -> 2 + 2 + 1
// (a lone digit d is looked up as the pair (0, d) = 9 + score(d): residue C pads it)
__device__ __forceinline__ uint32_t sx_self_score2(const uint8_t* __restrict__ ss2, uint32_t kmer, int k) {
  const uint32_t q1 = kmer / 441u, r1 = kmer - q1 * 441u;
  const uint32_t q2 = q1 / 441u, r2 = q1 - q2 * 441u;
  if (k == 5) return (uint32_t)ss2[r1] + ss2[r2] + ss2[q2] - 9u;
  const uint32_t q3 = q2 / 441u, r3 = q2 - q3 * 441u;
  return (uint32_t)ss2[r1] + ss2[r2] + ss2[r3] + ss2[q3] - 9u;
}

template <bool CROSS, uint32_t WCAP>
__global__ void __launch_bounds__(kWbWarps * 32) sx_warp_bucket_kernel(SxBucketArgs A) {
  constexpr int RMAX = WCAP / 32;  // rounds
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  uint8_t* s_ss2 = dyn_smem;
  for (uint32_t i = threadIdx.x; i < 441u; i += kWbWarps * 32) s_ss2[i] = (uint8_t)kmer_self_score(i, 2);
  __syncthreads();
  uint8_t* base = dyn_smem + kWbTableBytes + (size_t)warp * sx_wb_warp_bytes<WCAP>();
  uint2* X = reinterpret_cast<uint2*>(base);
  uint2* Y = X + WCAP;
  uint16_t* cnt = reinterpret_cast<uint16_t*>(Y + WCAP);  // [128]
  const uint32_t n_buckets = A.plan.n_buckets(), c2 = A.plan.c2, r = A.plan.r;
  const uint32_t n_pass = (r + 6u) / 7u, pass_bits = (r + n_pass - 1u) / n_pass, dmask = (1u << pass_bits) - 1u;
  const uint32_t keyshift = 32u - r;  // the key = the low r bits of the hash: pass p takes bits [p * pass_bits, ..)
  (void)keyshift;
  const uint32_t lt = lanemask_lt(), gt = ~lt & ~(1u << lane);
  const uint32_t gw = blockIdx.x * kWbWarps + warp, nw = gridDim.x * kWbWarps;
  unsigned long long multi = 0, work = 0;
  uint32_t n_distinct = 0, n_rep = 0, nnz_t = 0, n_kept = 0, max_bucket = 0;

  for (uint32_t b = gw; b < n_buckets; b += nw) {
    const uint32_t beg = __ldg(A.h2 + (size_t)b * c2), end = __ldg(A.h2 + (size_t)(b + 1u) * c2);
    const uint32_t nrec = end - beg;
    if (nrec == 0) continue;
    max_bucket = max(max_bucket, nrec);
    if (nrec > WCAP) {
      if (lane == 0) {
        if (nrec > A.mid_cap) A.huge_list[atomicAdd(A.huge_cnt, 1u)] = b;
        else A.mid_list[atomicAdd(A.mid_cnt, 1u)] = b;
      }
      continue;
    }
    const uint32_t R = (nrec + 31u) >> 5;
    __syncwarp();
    for (uint32_t u = 0; u < R; ++u) {
      const uint32_t i = u * 32u + lane;
      if (i < nrec) X[i] = ld_stream_u32x2(A.rec + beg + i);
    }
    if (b + nw < n_buckets) {  // the warp's next bucket: pull it into L2 now (one 128-byte line per lane)
      const uint32_t nbeg = __ldg(A.h2 + (size_t)(b + nw) * c2), nend = __ldg(A.h2 + (size_t)(b + nw + 1u) * c2);
      const uint8_t* pf = reinterpret_cast<const uint8_t*>(A.rec + nbeg) + lane * 128u;
      if (nbeg + lane * 16u < nend) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
    }
    // ---- stable LSD sort on the low r hash bits
    for (uint32_t p = 0; p < n_pass; ++p) {
      const uint32_t sh = p * pass_bits;
      reinterpret_cast<uint32_t*>(cnt)[lane] = 0;
      reinterpret_cast<uint32_t*>(cnt)[lane + 32u] = 0;
      __syncwarp();
      uint32_t rk[RMAX / 2 > 0 ? RMAX / 2 : 1];  // two 16-bit ranks per register
#pragma unroll
      for (int q = 0; q < (RMAX / 2 > 0 ? RMAX / 2 : 1); ++q) rk[q] = 0;
      // the matches of every round up front (independent: they pipeline), then the serial counter chain
      uint32_t peers[RMAX];
#pragma unroll
      for (int u = 0; u < RMAX; ++u) {
        peers[u] = 0;
        if ((uint32_t)u < R) {  // uniform
          const uint32_t i = (uint32_t)u * 32u + lane;
          const uint32_t d = i < nrec ? (sx_hash(X[i].x) >> sh) & dmask : (0x80000000u | lane);
          peers[u] = __match_any_sync(kFullMask, d);
        }
      }
#pragma unroll
      for (int u = 0; u < RMAX; ++u) {
        if ((uint32_t)u < R) {
          const uint32_t i = (uint32_t)u * 32u + lane;
          const bool valid = i < nrec;
          const uint32_t d = valid ? (sx_hash(X[i].x) >> sh) & dmask : 0u;
          const uint32_t rin = __popc(peers[u] & lt);
          const uint32_t before = valid ? cnt[d] : 0u;
          __syncwarp();
          if (valid && rin == 0) cnt[d] = (uint16_t)(before + __popc(peers[u]));
          __syncwarp();
          rk[u >> 1] |= (before + rin) << (16 * (u & 1));
        }
      }
      {  // digit totals -> starts: lane l owns digits 4 l .. 4 l + 3
        const uint2 t = reinterpret_cast<const uint2*>(cnt)[lane];
        const uint32_t c0 = t.x & 0xFFFFu, c1 = t.x >> 16, c2_ = t.y & 0xFFFFu, c3 = t.y >> 16;
        const uint32_t sum = c0 + c1 + c2_ + c3;
        const uint32_t ex = warp_scan_incl(sum) - sum;
        __syncwarp();
        reinterpret_cast<uint2*>(cnt)[lane] = make_uint2(ex | ((ex + c0) << 16), (ex + c0 + c1) | ((ex + c0 + c1 + c2_) << 16));
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < RMAX; ++u) {
        if ((uint32_t)u < R) {
          const uint32_t i = (uint32_t)u * 32u + lane;
          if (i < nrec) {
            const uint2 v = X[i];
            const uint32_t d = (sx_hash(v.x) >> sh) & dmask;
            Y[cnt[d] + ((rk[u >> 1] >> (16 * (u & 1))) & 0xFFFFu)] = v;
          }
        }
      }
      __syncwarp();
      uint2* t = X;
      X = Y;
      Y = t;
    }
    // ---- X = sorted.  Drop the duplicates of a row (adjacent now), compact in place
    uint32_t nk = 0;
    {
      uint2 carry = make_uint2(kSentinel, kSentinel);
      for (uint32_t u = 0; u < R; ++u) {
        const uint32_t i = u * 32u + lane;
        const uint2 v = i < nrec ? X[i] : make_uint2(kSentinel, kSentinel);
        uint2 pv;
        pv.x = __shfl_up_sync(kFullMask, v.x, 1);
        pv.y = __shfl_up_sync(kFullMask, v.y, 1);
        if (lane == 0) pv = carry;
        carry.x = __shfl_sync(kFullMask, v.x, 31);
        carry.y = __shfl_sync(kFullMask, v.y, 31);
        const bool keep = i < nrec && !(v.x == pv.x && v.y == pv.y);
        const uint32_t m = __ballot_sync(kFullMask, keep);
        __syncwarp();
        if (keep) X[nk + __popc(m & lt)] = v;
        nk += __popc(m);
        n_kept += keep && A.owner.mine(v.y);
        __syncwarp();
      }
    }
    // ---- backward: next k-mer head and next run head (new k-mer or new 64-row bin) behind every record
    uint32_t* meta = reinterpret_cast<uint32_t*>(Y);
    const uint32_t G = (nk + 31u) >> 5;
    {
      uint32_t carry_h = nk, carry_rh = nk;
      for (uint32_t u = G; u-- > 0;) {
        const uint32_t c = u * 32u + lane, c0 = u * 32u;
        const bool in = c < nk;
        const uint2 v = in ? X[c] : make_uint2(kSentinel, kSentinel);
        uint2 pv;
        pv.x = __shfl_up_sync(kFullMask, v.x, 1);
        pv.y = __shfl_up_sync(kFullMask, v.y, 1);
        if (lane == 0) pv = c > 0 ? X[c - 1u] : make_uint2(kSentinel, kSentinel);
        const bool head = in && v.x != pv.x;
        const bool rh = in && (head || (v.y >> kBinRowsLog) != (pv.y >> kBinRowsLog));
        const uint32_t mh = __ballot_sync(kFullMask, head), mrh = __ballot_sync(kFullMask, rh);
        const uint32_t ge = (mh & gt) ? c0 + (__ffs(mh & gt) - 1u) : carry_h;
        const uint32_t gr = (mrh & gt) ? c0 + (__ffs(mrh & gt) - 1u) : carry_rh;
        if (in) meta[c] = ge | (gr << 16);
        if (mh) carry_h = c0 + (__ffs(mh) - 1u);
        if (mrh) carry_rh = c0 + (__ffs(mrh) - 1u);
        n_distinct += head && A.owner.mine(v.y);
      }
    }
    __syncwarp();
    // ---- forward: postings, vocabulary, entries, run records
    const uint32_t col_base = beg, id_base = beg >> 1;
    uint32_t carryP = 0, carryI = 0;
    for (uint32_t u = 0; u < G; ++u) {
      const uint32_t c = u * 32u + lane, c0 = u * 32u;
      const bool in = c < nk;
      const uint2 v = in ? X[c] : make_uint2(kSentinel, kSentinel);
      uint2 pv;
      pv.x = __shfl_up_sync(kFullMask, v.x, 1);
      pv.y = __shfl_up_sync(kFullMask, v.y, 1);
      if (lane == 0) pv = c > 0 ? X[c - 1u] : make_uint2(kSentinel, kSentinel);
      const uint32_t mt = in ? meta[c] : 0u;
      const uint32_t ge = mt & 0xFFFFu, gr = mt >> 16;
      const bool head = in && v.x != pv.x;
      const bool rep = in && !(head && ge == c + 1u);
      const bool rh = in && (head || (v.y >> kBinRowsLog) != (pv.y >> kBinRowsLog));
      const uint32_t mrep = __ballot_sync(kFullMask, rep), mreph = __ballot_sync(kFullMask, rep && head);
      if (!mrep) continue;  // a round of singletons
      const uint32_t P = carryP + __popc(mrep & lt), I = carryI + __popc(mreph & lt);
      carryP += __popc(mrep);
      carryI += __popc(mreph);
      // the rows' bins: adjacent lanes of one bin share a reservation, issued first (and the bin's region
      // start is fetched now) so that both latencies hide behind the rest of the round
      const uint32_t bin = rep ? v.y >> kBinRowsLog : kSentinel;
      const uint32_t r0 = v.y & ~(kBinRows - 1u);
      size_t region = 0;
      if (rep) region = bin_region(A.rowcap_prefix, r0);
      const bool mine = rep && A.owner.mine(v.y);
      const bool leader = mine && rh && gr - c >= 2u;
      const uint32_t pbin = __shfl_up_sync(kFullMask, bin, 1);
      const bool seghead = rep && (lane == 0 || pbin != bin);
      const uint32_t mseg = __ballot_sync(kFullMask, seghead), mlead = __ballot_sync(kFullMask, leader);
      uint32_t s0 = lane, range = 0;
      if (rep) {
        s0 = 31u - __clz(mseg & ~gt);
        const uint32_t e0 = (mseg & gt) ? (uint32_t)__ffs(mseg & gt) - 1u : 32u;
        range = (e0 >= 32u ? 0xFFFFFFFFu : ((1u << e0) - 1u)) & ~((1u << s0) - 1u);
      }
      uint32_t rbase = 0;
      if (rep && lane == s0) rbase = atomicAdd(&A.bin_cursor[bin], (uint32_t)(__popc(mrep & range) + __popc(mlead & range)));
      // run masks: OR of the row bits over the lanes [c, min(gr, end of the round)), by a suffix OR-scan that
      // stops at run heads; the leader of a run that continues into the next round adds the rest itself
      unsigned long long rmask = 0;
      if (mlead) {  // uniform
        const uint32_t mrh = __ballot_sync(kFullMask, rh);
        uint32_t lo = 0, hi = 0;
        if (in) {
          const uint32_t bit = v.y & (kBinRows - 1u);
          lo = bit < 32u ? 1u << bit : 0u;
          hi = bit < 32u ? 0u : 1u << (bit - 32u);
        }
        // lanes until my run's end inside the round: up to the next run head (exclusive)
        const uint32_t next_rh = (mrh & gt) ? (uint32_t)__ffs(mrh & gt) - 1u : 32u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t tlo = __shfl_down_sync(kFullMask, lo, o), thi = __shfl_down_sync(kFullMask, hi, o);
          if (lane + (uint32_t)o < next_rh) {
            lo |= tlo;
            hi |= thi;
          }
        }
        if (leader) {
          rmask = ((unsigned long long)hi << 32) | lo;
          for (uint32_t x = c0 + 32u; x < gr; ++x) rmask |= 1ull << (X[x].y & (kBinRows - 1u));
        }
      }
      uint4 ent = make_uint4(0, 0, 0, 0);
      if (rep) {
        const uint32_t row = v.y;
        const uint32_t post = col_base + P;
        const uint32_t lid = I - (head ? 0u : 1u);
        const uint32_t ss = sx_self_score2(s_ss2, v.x, A.k);
        A.col[post] = row;
        if (head) {
          const uint32_t f = ge - c;
          A.vocab[id_base + lid] = v.x;
          A.freq[id_base + lid] = f;
          A.selfscore[id_base + lid] = (uint8_t)ss;
          if (mine) {  // the first holder's rank owns the k-mer's totals
            ++n_rep;
            nnz_t += f;
            multi += (unsigned long long)f * (f - 1u) / 2u;
          }
        }
        uint32_t a = gr;
        if (CROSS) {
          const uint32_t target = A.first_after[row];
          uint32_t lo = a, hi = ge;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (X[mid].y < target) lo = mid + 1u; else hi = mid;
          }
          a = lo;
        }
        const uint32_t len = ge - a;
        if (mine) work += len;
        const uint2 sf = len == 1u ? make_uint2(X[a].y, kSentinel) : make_uint2(post + (a - c), post + (ge - c));
        ent = make_uint4(row | (ss << 24), id_base + lid, sf.x, sf.y);
      }
      rbase = __shfl_sync(kFullMask, rbase, s0);
      if (rep) {
        const uint32_t mine_before = range & lt;
        const uint32_t at = rbase + __popc(mrep & mine_before) + __popc(mlead & mine_before);
        uint4* d = A.entries + region + at;
        d[0] = ent;
        if (leader)
          d[1] = make_uint4(0x80000000u | (ent.x >> 24), r0 >> kBinRowsLog, (uint32_t)rmask, (uint32_t)(rmask >> 32));
      }
    }
  }
  const unsigned long long n_dist64 = warp_sum64(n_distinct), n_rep64 = warp_sum64(n_rep), nnz64 = warp_sum64(nnz_t),
                           kept64 = warp_sum64(n_kept);
  multi = warp_sum64(multi);
  work = warp_sum64(work);
  if (lane == 0) {
    if (n_dist64) atomicAdd(&A.g->n_distinct, n_dist64);
    if (n_rep64) atomicAdd(&A.g->n_repeated, n_rep64);
    if (nnz64) atomicAdd(&A.g->nnz, nnz64);
    if (multi) atomicAdd(&A.g->multi_total, multi);
    if (work) atomicAdd(&A.g->work_total, work);
    if (kept64) atomicAdd(A.n_incid, kept64);
    atomicMax(&A.g->max_bucket, max_bucket);
  }
}

template <uint32_t CAP>
__host__ __device__ constexpr size_t sx_bucket_cnt_bytes() {
  return (size_t)(CAP / 8 / 32) * 256 * 2 > 4 * 132 * 4 ? (size_t)(CAP / 8 / 32) * 256 * 2 : (size_t)4 * 132 * 4;
}
template <uint32_t CAP>
__host__ __device__ constexpr size_t sx_bucket_smem() {
  // records | two sort buffers (later: the compacted, sorted records) | warp counters (later: the block
  // arrays) | dstart | scan scratch
  return (size_t)CAP * 8 + (size_t)CAP * 8 + sx_bucket_cnt_bytes<CAP>() + 260 * 4 + 40 * 4;
}

// The phases after the sort work on "blocks" of 32 consecutive elements (one warp, one round): what a lane
// needs from its own block comes from ballots, what it needs from the other blocks from four small arrays
// (one entry per block) that warp 0 scans.  nb <= 128 blocks, 4 per lane.
__device__ __forceinline__ void sx_scan_blocks(uint32_t nb, uint32_t* __restrict__ sum, uint32_t* __restrict__ fh,
                                               uint32_t* __restrict__ frh, uint32_t* __restrict__ lh,
                                               uint32_t* __restrict__ total) {
  const uint32_t lane = lane_id();
  uint32_t v[4];
  // exclusive sum (packed halves: both stay below 2^16)
  uint32_t acc = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    v[q] = 4u * lane + q < nb ? sum[4u * lane + q] : 0u;
    acc += v[q];
  }
  uint32_t ex = warp_scan_incl(acc) - acc;
  const uint32_t tot = __shfl_sync(kFullMask, ex + acc, 31);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (4u * lane + q < nb) sum[4u * lane + q] = ex;
    ex += v[q];
  }
  if (lane == 0) *total = tot;
  if (!fh) return;
  // fh / frh: smallest position in the LATER blocks (exclusive suffix min); lh: largest in the EARLIER ones
  for (int which = 0; which < 2; ++which) {
    uint32_t* arr = which ? frh : fh;
    uint32_t m = kSentinel;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = 4u * lane + q < nb ? arr[4u * lane + q] : kSentinel;
      m = min(m, v[q]);
    }
    uint32_t sfx = m;  // inclusive suffix min over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_down_sync(kFullMask, sfx, o);
      if (lane + (uint32_t)o < 32u) sfx = min(sfx, t);
    }
    uint32_t after = __shfl_down_sync(kFullMask, sfx, 1);
    if (lane == 31u) after = kSentinel;
#pragma unroll
    for (int q = 3; q >= 0; --q) {
      if (4u * lane + q < nb) arr[4u * lane + q] = after;
      after = min(after, v[q]);
    }
  }
  {
    uint32_t m = 0;  // positions are stored + 1 (0 = none)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = 4u * lane + q < nb ? lh[4u * lane + q] : 0u;
      m = max(m, v[q]);
    }
    uint32_t pfx = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFullMask, pfx, o);
      if (lane >= (uint32_t)o) pfx = max(pfx, t);
    }
    uint32_t before = __shfl_up_sync(kFullMask, pfx, 1);
    if (lane == 0) before = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (4u * lane + q < nb) lh[4u * lane + q] = before;
      before = max(before, v[q]);
    }
  }
}

template <bool CROSS, uint32_t CAP>
__global__ void __launch_bounds__(CAP / 8, CAP == 4096 ? 2 : 4) sx_bucket_kernel(SxBucketArgs A) {
  constexpr int THREADS = CAP / 8;
  constexpr int V = 8;
  constexpr int WARPS = THREADS / 32;
  constexpr uint32_t IDXBITS = CAP == 4096 ? 12 : (CAP == 512 ? 9 : 13);
  static_assert((1u << IDXBITS) == CAP, "CAP must be 512, 4096 or 8192");
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  uint2* s_rec = reinterpret_cast<uint2*>(dyn_smem);                       // [CAP] as loaded (row order)
  uint32_t* s_a = reinterpret_cast<uint32_t*>(dyn_smem + (size_t)CAP * 8);   // [CAP] sort ping
  uint32_t* s_b = s_a + CAP;                                               // [CAP] sort pong
  uint2* s_cmp = reinterpret_cast<uint2*>(s_a);                            // [CAP] sorted + deduplicated records
  uint16_t* s_cnt = reinterpret_cast<uint16_t*>(dyn_smem + (size_t)CAP * 16);  // [WARPS][256]
  uint32_t* s_dstart = reinterpret_cast<uint32_t*>(dyn_smem + (size_t)CAP * 16 + sx_bucket_cnt_bytes<CAP>());  // [257]
  uint32_t* s_wsum = s_dstart + 260;                                       // [33]
  uint32_t* s_bsum = reinterpret_cast<uint32_t*>(s_cnt);                   // [132] per block: counts -> prefix (the counters are dead)
  uint32_t* s_bfh = s_bsum + 132;                                          // first head -> next head after the block
  uint32_t* s_bfrh = s_bfh + 132;                                          // the same for run heads
  uint32_t* s_blh = s_bfrh + 132;                                          // last head + 1 -> last head before the block
  __shared__ uint32_t s_total;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t c2 = A.plan.c2, r = A.plan.r;
  const uint32_t n_mid = *A.mid_cnt;
  const uint32_t keymask = r >= 32u ? 0xFFFFFFFFu : ((1u << r) - 1u);
  const uint32_t n_pass = (r + 7u) / 8u, pass_bits = (r + n_pass - 1u) / n_pass;
  const bool ballots = false;  // (8-bit digits: match.any measured faster here than eight ballots per round)
  const uint32_t lt = lanemask_lt(), gt = ~lt & ~(1u << lane);
  unsigned long long multi = 0, work = 0;
  uint32_t n_distinct = 0, n_rep = 0, nnz_t = 0, n_kept = 0, max_bucket = 0;

  for (uint32_t mi = blockIdx.x; mi < n_mid; mi += gridDim.x) {
    __syncthreads();
    const uint32_t b = A.mid_list[mi];
    const uint32_t beg = A.h2[(size_t)b * c2], end = A.h2[(size_t)(b + 1u) * c2];
    const uint32_t nrec = end - beg;  // (0 < nrec <= CAP: the warp kernel sorted the buckets into the lists)
    max_bucket = max(max_bucket, nrec);
    // every warp takes `rounds` x 32 consecutive records: all warps are busy whatever the bucket's fill
    const uint32_t rounds = (nrec + THREADS - 1u) / THREADS, seg = rounds * 32u;
    // ---- load (warp-striped) and sort by the remaining hash bits: stable LSD
    uint32_t item[V], digit[V], pos[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t i = warp * seg + (uint32_t)j * 32u + lane;
      item[j] = kSentinel;
      if ((uint32_t)j < rounds && i < nrec) {
        const uint2 v = ld_stream_u32x2(A.rec + beg + i);
        s_rec[i] = v;
        item[j] = ((sx_hash(v.x) & keymask) << IDXBITS) | i;
      }
    }
    uint32_t* src = s_a;
    uint32_t* dst = s_b;
    for (uint32_t p = 0; p < n_pass; ++p) {
#pragma unroll
      for (int j = 0; j < V; ++j)
        digit[j] = item[j] == kSentinel ? kNoDigit : (item[j] >> (IDXBITS + pass_bits * p)) & ((1u << pass_bits) - 1u);
      sx_tile_rank<THREADS, V>(digit, pos, pass_bits, rounds, ballots, s_cnt, s_dstart, s_wsum);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if ((uint32_t)j < rounds && digit[j] != kNoDigit) dst[pos[j]] = item[j];
      __syncthreads();
      if (p + 1u < n_pass) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const uint32_t i = warp * seg + (uint32_t)j * 32u + lane;
          item[j] = (uint32_t)j < rounds && i < nrec ? dst[i] : kSentinel;
        }
      }
      uint32_t* t = src;
      src = dst;
      dst = t;
    }
    // ---- src = sorted items.  From here on element q = u * THREADS + tid (CTA-striped): round u of warp w is
    // block u * WARPS + w.  Drop the duplicates of a row (adjacent now) and compact.
    const uint32_t nb = rounds * WARPS;
    uint32_t km[V], rw[V];
    uint32_t keep = 0;
#pragma unroll
    for (int u = 0; u < V; ++u) {
      if ((uint32_t)u >= rounds) break;
      const uint32_t q = (uint32_t)u * THREADS + tid;
      uint2 v = make_uint2(kSentinel, kSentinel);
      if (q < nrec) v = s_rec[src[q] & (CAP - 1u)];
      uint2 pv;
      pv.x = __shfl_up_sync(kFullMask, v.x, 1);
      pv.y = __shfl_up_sync(kFullMask, v.y, 1);
      if (lane == 0) pv = q > 0 && q - 1u < nrec ? s_rec[src[q - 1u] & (CAP - 1u)] : make_uint2(kSentinel, kSentinel);
      const bool k1 = q < nrec && !(v.x == pv.x && v.y == pv.y);
      const uint32_t m = __ballot_sync(kFullMask, k1);
      if (lane == 0) s_bsum[(uint32_t)u * WARPS + warp] = __popc(m);
      if (k1) keep |= 1u << u;
      n_kept += k1 && A.owner.mine(v.y);
      km[u] = v.x;
      rw[u] = __popc(m & lt);  // (the row moves to rw[] below; until then: the offset inside the block)
      pos[u] = v.y;
    }
    __syncthreads();  // (every read of src and s_rec is done)
    if (warp == 0) sx_scan_blocks(nb, s_bsum, nullptr, nullptr, nullptr, &s_total);
    __syncthreads();
    const uint32_t nk = s_total;
#pragma unroll
    for (int u = 0; u < V; ++u) {
      if ((uint32_t)u >= rounds) break;
      if ((keep >> u) & 1u) s_cmp[s_bsum[(uint32_t)u * WARPS + warp] + rw[u]] = make_uint2(km[u], pos[u]);
    }
    __syncthreads();
    // ---- groups: heads, run heads (a new k-mer or a new 64-row bin), repeated k-mers; per block for the scans
    const uint32_t grounds = (nk + THREADS - 1u) / THREADS, gnb = grounds * WARPS;
#pragma unroll
    for (int u = 0; u < V; ++u) {
      if ((uint32_t)u >= grounds) break;
      const uint32_t c = (uint32_t)u * THREADS + tid;
      const bool in = c < nk;
      const uint2 v = in ? s_cmp[c] : make_uint2(kSentinel, kSentinel);
      const uint2 pv = in && c > 0 ? s_cmp[c - 1u] : make_uint2(kSentinel, kSentinel);
      const uint32_t nx = c + 1u < nk ? s_cmp[c + 1u].x : kSentinel;
      const bool head = in && v.x != pv.x, headn = nx != v.x;
      const bool rep = in && !(head && headn);
      const bool rh = in && (head || (v.y >> kBinRowsLog) != (pv.y >> kBinRowsLog));
      const uint32_t mh = __ballot_sync(kFullMask, head), mrh = __ballot_sync(kFullMask, rh);
      const uint32_t mrep = __ballot_sync(kFullMask, rep), mreph = __ballot_sync(kFullMask, rep && head);
      if (lane == 0) {
        const uint32_t blk = (uint32_t)u * WARPS + warp, c0 = c;
        s_bsum[blk] = __popc(mrep) | (__popc(mreph) << 16);
        s_bfh[blk] = mh ? c0 + (__ffs(mh) - 1u) : kSentinel;
        s_bfrh[blk] = mrh ? c0 + (__ffs(mrh) - 1u) : kSentinel;
        s_blh[blk] = mh ? c0 + (31u - __clz(mh)) + 1u : 0u;
      }
      n_distinct += head && A.owner.mine(v.y);
    }
    __syncthreads();
    if (warp == 0) sx_scan_blocks(gnb, s_bsum, s_bfh, s_bfrh, s_blh, &s_total);
    __syncthreads();
    const uint32_t nnz_b = s_total & 0xFFFFu, nrep_b = s_total >> 16;
    if (nnz_b == 0) continue;  // only singletons
    (void)nrep_b;
    const uint32_t col_base = beg, id_base = beg >> 1;
    // ---- postings, vocabulary, entries, run records: one record per thread and round, global stores coalesced
#pragma unroll 1
    for (uint32_t u = 0; u < grounds; ++u) {
      const uint32_t c = u * THREADS + tid, blk = u * WARPS + warp;
      const bool in = c < nk;
      const uint2 v = in ? s_cmp[c] : make_uint2(kSentinel, kSentinel);
      const uint2 pv = in && c > 0 ? s_cmp[c - 1u] : make_uint2(kSentinel, kSentinel);
      const uint32_t nx = c + 1u < nk ? s_cmp[c + 1u].x : kSentinel;
      const bool head = in && v.x != pv.x, headn = nx != v.x;
      const bool rep = in && !(head && headn);
      const bool rh = in && (head || (v.y >> kBinRowsLog) != (pv.y >> kBinRowsLog));
      const uint32_t mh = __ballot_sync(kFullMask, head), mrh = __ballot_sync(kFullMask, rh);
      const uint32_t mrep = __ballot_sync(kFullMask, rep), mreph = __ballot_sync(kFullMask, rep && head);
      const uint32_t c0 = c - lane;
      const uint32_t bs = s_bsum[blk];
      const uint32_t P = (bs & 0xFFFFu) + __popc(mrep & lt);         // repeated records before c
      const uint32_t I = (bs >> 16) + __popc(mreph & lt);            // repeated heads before c
      const uint32_t ge = min(nk, (mh & gt) ? c0 + (__ffs(mh & gt) - 1u) : s_bfh[blk]);    // next head after c
      const uint32_t gr = min(nk, (mrh & gt) ? c0 + (__ffs(mrh & gt) - 1u) : s_bfrh[blk]);  // next run head after c
      // the rows' bins: adjacent lanes of one bin share a reservation, issued first (its latency hides behind
      // the rest of the round)
      const uint32_t bin = rep ? v.y >> kBinRowsLog : kSentinel;
      const bool mine = rep && A.owner.mine(v.y);
      const bool leader = mine && rh && gr - c >= 2u;
      const uint32_t pbin = __shfl_up_sync(kFullMask, bin, 1);
      const bool seghead = rep && (lane == 0 || pbin != bin);
      const uint32_t mseg = __ballot_sync(kFullMask, seghead), mlead = __ballot_sync(kFullMask, leader);
      uint32_t s0 = lane, range = 0;
      if (rep) {
        s0 = 31u - __clz(mseg & ~gt);                                            // my segment's first lane
        const uint32_t e0 = (mseg & gt) ? (uint32_t)__ffs(mseg & gt) - 1u : 32u;    // one past its last lane
        range = (e0 >= 32u ? 0xFFFFFFFFu : ((1u << e0) - 1u)) & ~((1u << s0) - 1u);
      }
      uint32_t base = 0;
      if (rep && lane == s0) base = atomicAdd(&A.bin_cursor[bin], (uint32_t)(__popc(mrep & range) + __popc(mlead & range)));
      uint4 ent = make_uint4(0, 0, 0, 0);
      unsigned long long rmask = 0;
      if (rep) {
        const uint32_t row = v.y;
        const uint32_t post = col_base + P;
        const uint32_t lid = I - (head ? 0u : 1u);
        const uint32_t ss = sx_self_score(A.ss3, v.x, A.k);
        A.col[post] = row;
        if (head) {
          const uint32_t f = ge - c;
          A.vocab[id_base + lid] = v.x;
          A.freq[id_base + lid] = f;
          A.selfscore[id_base + lid] = (uint8_t)ss;
          if (mine) {
            ++n_rep;
            nnz_t += f;
            multi += (unsigned long long)f * (f - 1u) / 2u;
          }
        }
        // the suffix starts behind the row's bin-run (its pairs go to the tiles); cross-class mode: and not
        // before the first holder of a later class block
        uint32_t a = gr;
        if (CROSS) {
          const uint32_t target = A.first_after[row];
          uint32_t lo = a, hi = ge;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_cmp[mid].y < target) lo = mid + 1u; else hi = mid;
          }
          a = lo;
        }
        const uint32_t len = ge - a;
        if (mine) work += len;
        const uint2 sf = len == 1u ? make_uint2(s_cmp[a].y, kSentinel) : make_uint2(post + (a - c), post + (ge - c));
        ent = make_uint4(row | (ss << 24), id_base + lid, sf.x, sf.y);
        if (leader)
          for (uint32_t x = c; x < gr; ++x) rmask |= 1ull << (s_cmp[x].y & (kBinRows - 1u));
      }
      base = __shfl_sync(kFullMask, base, s0);
      if (rep) {
        const uint32_t mine_before = range & lt;
        const uint32_t at = base + __popc(mrep & mine_before) + __popc(mlead & mine_before);
        const uint32_t r0 = (v.y) & ~(kBinRows - 1u);
        uint4* d = A.entries + bin_region(A.rowcap_prefix, r0) + at;
        d[0] = ent;
        if (leader)
          d[1] = make_uint4(0x80000000u | (ent.x >> 24), r0 >> kBinRowsLog, (uint32_t)rmask, (uint32_t)(rmask >> 32));
      }
    }
  }
  const unsigned long long n_dist64 = warp_sum64(n_distinct), n_rep64 = warp_sum64(n_rep), nnz64 = warp_sum64(nnz_t),
                           kept64 = warp_sum64(n_kept);
  multi = warp_sum64(multi);
  work = warp_sum64(work);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) max_bucket = max(max_bucket, __shfl_xor_sync(kFullMask, max_bucket, o));
  if (lane == 0) {
    if (n_dist64) atomicAdd(&A.g->n_distinct, n_dist64);
    if (n_rep64) atomicAdd(&A.g->n_repeated, n_rep64);
    if (nnz64) atomicAdd(&A.g->nnz, nnz64);
    if (multi) atomicAdd(&A.g->multi_total, multi);
    if (work) atomicAdd(&A.g->work_total, work);
    if (kept64) atomicAdd(A.n_incid, kept64);
    atomicMax(&A.g->max_bucket, max_bucket);
  }
}

// ---------------------------------------------------------------------------------------
// A bucket that does not fit shared memory: the same steps through global memory, one CTA per
// bucket, tile by tile with carried state.  X / Y = the bucket's slice of the two record arrays.
//   1. stable LSD radix sort on the remaining hash bits (X <-> Y),
//   2. forward: drop adjacent duplicates, compact into Y,
//   3. backward: next group head after every position (stored in X), bucket totals,
//   4. forward: ids, postings, vocabulary, entries, run records.
// ---------------------------------------------------------------------------------------
constexpr int kHugeThreads = 512;
constexpr int kHugeV = 4;
constexpr uint32_t kHugeTile = kHugeThreads * kHugeV;

template <bool CROSS>
__global__ void __launch_bounds__(kHugeThreads) sx_huge_kernel(SxBucketArgs A) {
  __shared__ uint16_t s_cnt[kHugeThreads / 32][256];
  __shared__ uint32_t s_dstart[260], s_goff[256], s_wsum[40];
  __shared__ uint32_t s_carry[4];
  __shared__ unsigned long long s_base[2], s_tot[2];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t c2 = A.plan.c2, r = A.plan.r;
  const uint32_t keymask = r >= 32u ? 0xFFFFFFFFu : ((1u << r) - 1u);
  const uint32_t n_pass = (r + 7u) / 8u;
  const uint32_t n_huge = *A.huge_cnt;
  unsigned long long multi = 0, work = 0;
  uint32_t n_distinct = 0, n_rep = 0, nnz_t = 0;
  unsigned long long n_kept = 0;
  for (uint32_t hb = blockIdx.x; hb < n_huge; hb += gridDim.x) {
    const uint32_t b = A.huge_list[hb];
    const uint32_t beg = A.h2[(size_t)b * c2], end = A.h2[(size_t)(b + 1u) * c2];
    const uint32_t nrec = end - beg;
    uint2* X = const_cast<uint2*>(A.rec) + beg;
    uint2* Y = A.scratch + beg;
    // ---- 1. LSD passes
    for (uint32_t p = 0; p < n_pass; ++p) {
      const uint32_t sh = 8u * p;
      if (tid < 256) s_goff[tid] = 0;
      __syncthreads();
      for (uint32_t i = tid; i < nrec; i += kHugeThreads) atomicAdd(&s_goff[((sx_hash(X[i].x) & keymask) >> sh) & 255u], 1u);
      __syncthreads();
      {
        const uint32_t v = tid < 256 ? s_goff[tid] : 0u;
        uint32_t total;
        const uint32_t ex = sx_block_scan<kHugeThreads>(v, s_wsum, &total);
        if (tid < 256) s_goff[tid] = ex;
      }
      __syncthreads();
      for (uint32_t t0 = 0; t0 < nrec; t0 += kHugeTile) {
        uint32_t km[kHugeV], rw[kHugeV], digit[kHugeV], pos[kHugeV];
#pragma unroll
        for (int j = 0; j < kHugeV; ++j) {
          const uint32_t i = t0 + warp * (kHugeV * 32u) + (uint32_t)j * 32u + lane;
          digit[j] = kNoDigit;
          if (i < nrec) {
            const uint2 v = X[i];
            km[j] = v.x;
            rw[j] = v.y;
            digit[j] = ((sx_hash(v.x) & keymask) >> sh) & 255u;
          }
        }
        sx_tile_rank<kHugeThreads, kHugeV>(digit, pos, 8u, kHugeV, false, &s_cnt[0][0], s_dstart, s_wsum);
#pragma unroll
        for (int j = 0; j < kHugeV; ++j)
          if (digit[j] != kNoDigit) Y[s_goff[digit[j]] + (pos[j] - s_dstart[digit[j]])] = make_uint2(km[j], rw[j]);
        __syncthreads();
        if (tid < 256) s_goff[tid] += s_dstart[tid + 1] - s_dstart[tid];
        __syncthreads();
      }
      uint2* t = X;
      X = Y;
      Y = t;
      __threadfence_block();
      __syncthreads();
    }
    // X = sorted.  ---- 2. drop adjacent duplicates, compact into Y
    if (tid == 0) s_carry[0] = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < nrec; t0 += kHugeThreads) {
      const uint32_t q = t0 + tid;
      uint2 v = make_uint2(kSentinel, kSentinel);
      bool keep = false;
      if (q < nrec) {
        v = X[q];
        keep = true;
        if (q > 0) {
          const uint2 pv = X[q - 1u];
          keep = !(pv.x == v.x && pv.y == v.y);
        }
      }
      uint32_t total;
      const uint32_t ex = sx_block_scan<kHugeThreads>(keep ? 1u : 0u, s_wsum, &total);
      const uint32_t carry = s_carry[0];
      if (keep) Y[carry + ex] = v;
      n_kept += keep && A.owner.mine(v.y);
      __syncthreads();
      if (tid == 0) s_carry[0] = carry + total;
      __syncthreads();
    }
    const uint32_t nk = s_carry[0];
    __threadfence_block();
    __syncthreads();
    // ---- 3. backward: next head after every compacted position -> NX[c]; totals of the repeated groups
    uint32_t* NX = reinterpret_cast<uint32_t*>(X);
    if (tid == 0) {
      s_carry[1] = nk;  // smallest head position in the tiles already visited (to the right)
      s_tot[0] = 0;
      s_tot[1] = 0;
    }
    __syncthreads();
    const uint32_t n_tiles = (nk + kHugeThreads - 1u) / kHugeThreads;
    for (uint32_t ti = n_tiles; ti-- > 0;) {
      const uint32_t c = ti * kHugeThreads + tid;
      bool head = false;
      if (c < nk) head = c == 0 || Y[c - 1u].x != Y[c].x;
      // exclusive suffix-min of the head positions inside the tile, then the carry
      uint32_t v = head ? c : kSentinel;
      uint32_t sfx = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_down_sync(kFullMask, sfx, o);
        if (lane + (uint32_t)o < 32u) sfx = min(sfx, t);
      }
      __syncthreads();
      if (lane == 0) s_wsum[warp] = sfx;
      __syncthreads();
      uint32_t later = s_carry[1];
      for (uint32_t w = warp + 1u; w < (uint32_t)(kHugeThreads / 32); ++w) later = min(later, s_wsum[w]);
      uint32_t nxt = __shfl_down_sync(kFullMask, sfx, 1);
      if (lane == 31u) nxt = kSentinel;
      nxt = min(nxt, later);
      if (c < nk) {
        NX[c] = nxt;
        if (head && nxt - c >= 2u) {
          atomicAdd(&s_tot[0], (unsigned long long)(nxt - c));
          atomicAdd(&s_tot[1], 1ull);
        }
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t m = s_carry[1];
        for (uint32_t w = 0; w < (uint32_t)(kHugeThreads / 32); ++w) m = min(m, s_wsum[w]);
        s_carry[1] = m;
      }
      __syncthreads();
    }
    if (tid == 0) {
      s_base[0] = beg;
      s_base[1] = beg >> 1;
      s_carry[2] = 0;  // postings handed out so far
      s_carry[3] = 0;  // ids handed out so far
    }
    __threadfence_block();
    __syncthreads();
    const uint32_t col_base = (uint32_t)s_base[0], id_base = (uint32_t)s_base[1];
    // ---- 4. forward: emit
    auto rows = [&](uint32_t c) { return Y[c].y; };
    for (uint32_t t0 = 0; t0 < nk; t0 += kHugeThreads) {
      const uint32_t c = t0 + tid;
      bool head = false, rep = false;
      uint2 v = make_uint2(kSentinel, kSentinel);
      uint32_t ge = 0;
      if (c < nk) {
        v = Y[c];
        head = c == 0 || Y[c - 1u].x != v.x;
        const bool headn = c + 1u >= nk || Y[c + 1u].x != v.x;
        rep = !(head && headn);
        ge = NX[c];
      }
      uint32_t total;
      const uint32_t sv = rep ? 1u + (head ? 0x10000u : 0u) : 0u;
      // (a tile holds at most 512 records: the packed halves cannot overflow)
      const uint32_t ex = sx_block_scan<kHugeThreads>(sv, s_wsum, &total);
      const uint32_t post0 = s_carry[2], id0 = s_carry[3];
      const bool mine = c < nk && A.owner.mine(v.y);
      if (c < nk && head && mine) ++n_distinct;
      if (rep) {
        const uint32_t post = col_base + post0 + (ex & 0xFFFFu);
        const uint32_t gid = id_base + id0 + (ex >> 16) - (head ? 0u : 1u);
        const uint32_t ss = (uint32_t)kmer_self_score(v.x, A.k);
        A.col[post] = v.y;
        if (head) {
          const uint32_t f = ge - c;
          A.vocab[gid] = v.x;
          A.freq[gid] = f;
          A.selfscore[gid] = (uint8_t)ss;
          if (mine) {
            ++n_rep;
            nnz_t += f;
            multi += (unsigned long long)f * (f - 1u) / 2u;
          }
        }
        SxEmit em = sx_emit(rows, c, ge, head, v.y, ss, gid, post, CROSS ? A.first_after : nullptr);
        if (!mine) em.rmask = 0;
        if (mine) work += em.len;
        const uint32_t bin = v.y >> kBinRowsLog;
        const uint32_t at = atomicAdd(&A.bin_cursor[bin], 1u + (em.rmask ? 1u : 0u));
        uint4* d = A.entries + bin_region(A.rowcap_prefix, v.y & ~(kBinRows - 1u)) + at;
        d[0] = em.ent;
        if (em.rmask)
          d[1] = make_uint4(0x80000000u | ss, bin, (uint32_t)em.rmask, (uint32_t)(em.rmask >> 32));
      }
      __syncthreads();
      if (tid == 0) {
        s_carry[2] = post0 + (total & 0xFFFFu);
        s_carry[3] = id0 + (total >> 16);
      }
      __syncthreads();
    }
  }
  const unsigned long long n_dist64 = warp_sum64(n_distinct), n_rep64 = warp_sum64(n_rep), nnz64 = warp_sum64(nnz_t);
  n_kept = warp_sum64(n_kept);
  multi = warp_sum64(multi);
  work = warp_sum64(work);
  if (lane == 0) {
    if (n_dist64) atomicAdd(&A.g->n_distinct, n_dist64);
    if (n_rep64) atomicAdd(&A.g->n_repeated, n_rep64);
    if (nnz64) atomicAdd(&A.g->nnz, nnz64);
    if (multi) atomicAdd(&A.g->multi_total, multi);
    if (work) atomicAdd(&A.g->work_total, work);
    if (n_kept) atomicAdd(A.n_incid, n_kept);
  }
}

// row that holds the first residue of every tile of the stream (tile_row[n_tiles] = the last row): the
// level-1 kernels bound their per-thread row search with it instead of searching all rows per tile
__global__ void sx_tile_rows_kernel(const uint32_t* __restrict__ soff, uint32_t n, uint32_t R, uint32_t n_tiles,
                                    uint32_t* __restrict__ tile_row) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  const unsigned long long x64 = (unsigned long long)t * kSxTile;
  const uint32_t x = x64 >= R ? R - 1u : (uint32_t)x64;
  uint32_t lo = 0, hi = n - 1u;  // largest r with soff[r] <= x
  while (lo < hi) {
    const uint32_t mid = (lo + hi + 1u) >> 1;
    if (soff[mid] <= x) lo = mid; else hi = mid - 1u;
  }
  tile_row[t] = lo;
}

// self-score of a k-mer from the table of residue triples: k = 5 -> 3 + 2 digits, k = 7 -> 3 + 3 + 1
__device__ __forceinline__ uint32_t sx_self_score(const uint8_t* __restrict__ ss3, uint32_t kmer, int k) {
  const uint32_t q1 = kmer / 9261u, r1 = kmer - q1 * 9261u;   // low three digits
  // (a partial triple is padded with digit 0, residue C, which scores 9: take the padding out again)
  if (k == 5) return (uint32_t)__ldg(ss3 + r1) + __ldg(ss3 + q1) - 9u;
  const uint32_t q2 = q1 / 9261u, r2 = q1 - q2 * 9261u;       // next three; q2 = the seventh
  return (uint32_t)__ldg(ss3 + r1) + __ldg(ss3 + r2) + __ldg(ss3 + q2) - 18u;
}
__global__ void sx_ss3_kernel(uint8_t* __restrict__ ss3) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 9261u) ss3[i] = (uint8_t)kmer_self_score(i, 3);
}

// ---- dense view of the (capacity-placed) ids for the readback entry point kc_get_pair_index --------
__global__ void sx_mark_ids_kernel(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ rowlen, uint32_t n,
                                   const uint32_t* __restrict__ ids, uint32_t* __restrict__ used) {
  const uint32_t lane = lane_id();
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += (gridDim.x * blockDim.x) >> 5) {
    const uint32_t nl = rowlen[r], ps = rowptr[r];
    for (uint32_t i = lane; i < nl; i += 32) used[ids[ps + i]] = 1u;
  }
}
// dense[map[id]] = sparse[id] for the used ids; map = exclusive scan of `used` (used[id] is still readable
// through the next slot: map[id + 1] - map[id])
__global__ void sx_gather_vocab_kernel(const uint32_t* __restrict__ map, uint64_t n_slots,
                                       const uint32_t* __restrict__ vocab, const uint32_t* __restrict__ freq,
                                       const uint8_t* __restrict__ self, uint32_t* __restrict__ vocab_d,
                                       uint32_t* __restrict__ freq_d, uint8_t* __restrict__ self_d) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t m = map[i];
    if (map[i + 1] != m) {
      vocab_d[m] = vocab[i];
      freq_d[m] = freq[i];
      self_d[m] = self[i];
    }
  }
}
__global__ void sx_remap_ids_kernel(const uint32_t* __restrict__ map, uint32_t* __restrict__ ids, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    ids[i] = map[ids[i]];
}

// ---- small helpers of the build ----------------------------------------------------------
// k-mer positions of every row (the entry capacity of its bin), as the scan's input
struct SxPosIn {
  const uint32_t* soff;
  uint32_t k;
  __device__ unsigned long long operator()(uint64_t r) const {
    const uint32_t len = soff[r + 1] - soff[r];
    return len >= k ? len - k + 1u : 0u;
  }
};

// u32 prefix with the grand total in the slot behind the last element
struct SxExclOutTail {
  uint32_t* p;
  uint64_t n;
  __device__ void operator()(uint64_t i, unsigned long long excl, unsigned long long v) const {
    p[i] = (uint32_t)excl;
    if (i + 1 == n) p[n] = (uint32_t)(excl + v);
  }
};

// the same, shifted by a base the device holds (the end of the previous level-1 pass's region)
struct SxExclOutTailBase {
  uint32_t* p;
  uint64_t n;
  const uint32_t* base;  // null: 0
  __device__ void operator()(uint64_t i, unsigned long long excl, unsigned long long v) const {
    const uint32_t b = base ? *base : 0u;
    p[i] = b + (uint32_t)excl;
    if (i + 1 == n) p[n] = b + (uint32_t)(excl + v);
  }
};

// cross-class mode: the residue stream in the pair (class-major) order.  One warp per row.
__global__ void sx_permute_residues_kernel(const uint8_t* __restrict__ res, const uint32_t* __restrict__ pstart,
                                           const uint32_t* __restrict__ soff, uint32_t n, uint8_t* __restrict__ out) {
  const uint32_t lane = lane_id();
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += (gridDim.x * blockDim.x) >> 5) {
    const uint32_t s = pstart[r], d = soff[r], len = soff[r + 1] - d;
    for (uint32_t i = lane; i < len; i += 32) out[d + i] = res[s + i];
  }
}

}  // namespace kc
