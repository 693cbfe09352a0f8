// index.cuh — K4/K5: the perfect k-mer index, per-protein id lists, kmer_freq, postings.
//
// Replaces (reference root relative):
//   split unique/repeated + Mphf::new x2 + unique bool table   src/main.rs:127-147
//       -> rank dictionary over the census "held by >= 2" bits: dict[w] = {bits, #set bits before word w};
//          id(kmer) = dict[kmer>>5].y + popc(dict[kmer>>5].x & ((1<<(kmer&31))-1)).
//          A minimal perfect, order-preserving hash of the repeated k-mers (one 8-byte load).
//   remove_unique_five_mers + modify_hash_five_mer + kmer_freq  src/protein.rs:151-174,
//       src/main.rs:182-193 -> ids_freq_kernel
//   times_kmer_visited / triangular edge layout                 src/graph/vertex.rs:92-136
//       -> postings (holders of every id, sorted by protein rank) + per-entry suffix ranges
#pragma once
#include "common.cuh"

namespace kc {

// gather the odd (ODD=1) or even (ODD=0) bits of a census word into its low 16 bits
template <int ODD>
__device__ __forceinline__ uint32_t gather_bits(uint32_t s) {
  uint32_t x = (s >> ODD) & 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  x = (x | (x >> 8)) & 0x0000FFFFu;
  return x;
}
// presence bits of dictionary word i (32 consecutive k-mers) from census words 2i, 2i+1
template <int ODD>
__device__ __forceinline__ uint32_t dict_bits(const uint32_t* __restrict__ seen, uint64_t i) {
  const uint2 s = *reinterpret_cast<const uint2*>(seen + 2 * i);
  return gather_bits<ODD>(s.x) | (gather_bits<ODD>(s.y) << 16);
}
template <int ODD>
struct PopcIn {
  const uint32_t* seen;
  __device__ unsigned long long operator()(uint64_t i) const { return __popc(dict_bits<ODD>(seen, i)); }
};
template <int ODD>
struct DictOut {
  const uint32_t* seen;
  uint2* dict;
  __device__ void operator()(uint64_t i, unsigned long long excl, unsigned long long) const {
    dict[i] = make_uint2(dict_bits<ODD>(seen, i), (uint32_t)excl);
  }
};

// number of k-mers held by >= 1 protein (even bits of the census words)
__global__ void popc_reduce_kernel(const uint32_t* __restrict__ bits, uint64_t n_words,
                                   unsigned long long* __restrict__ total) {
  unsigned long long s = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words;
       i += (uint64_t)gridDim.x * blockDim.x)
    s += __popc(bits[i] & 0x55555555u);
  s = warp_sum64(s);
  if (lane_id() == 0 && s) atomicAdd(total, s);
}

// B62[r][r] in the reference's residue order (src/blosum.rs:8-30 diagonal); code 20 -> 0
__constant__ uint8_t c_blosum_diag[21] = {9, 4, 5, 4, 6, 7, 6, 5, 5, 6, 8, 5, 5, 5, 4, 4, 4, 11, 7, 6, 0};
// (the same table as 4-bit fields of two registers: a constant-memory load with a per-lane index is
// replayed once per distinct index)
__device__ __forceinline__ int kmer_self_score(uint32_t kmer, int k) {
  constexpr unsigned long long lo = 0x4455586556764549ull;  // codes 0..15
  constexpr uint32_t hi = 0x67b4u;                        // codes 16..20
  int s = 0;
  for (int i = 0; i < k; ++i) {
    const uint32_t q = kmer / 21u, c = kmer - q * 21u;
    s += c < 16u ? (int)((lo >> (4u * c)) & 15ull) : (int)((hi >> (4u * (c - 16u))) & 15u);
    kmer = q;
  }
  return s;
}

// expand a bitmap into the ascending list of set positions (vocab[id] = kmer); optional
// BLOSUM self-score per id.  `dict` supplies the rank of each word.
__global__ void expand_bitmap_kernel(const uint2* __restrict__ dict, uint64_t n_words,
                                     uint32_t* __restrict__ vocab, uint8_t* __restrict__ selfscore, int k) {
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words;
       w += (uint64_t)gridDim.x * blockDim.x) {
    uint2 d = dict[w];
    uint32_t bits = d.x, id = d.y;
    while (bits) {
      const uint32_t b = __ffs(bits) - 1;
      const uint32_t kmer = (uint32_t)(w << 5) + b;
      vocab[id] = kmer;
      if (selfscore) selfscore[id] = (uint8_t)kmer_self_score(kmer, k);
      ++id;
      bits &= bits - 1;
    }
  }
}

__device__ __forceinline__ uint32_t dict_lookup(const uint2* __restrict__ dict, uint32_t kmer) {
  const uint2 d = dict[kmer >> 5];
  const uint32_t bit = 1u << (kmer & 31u);
  return (d.x & bit) ? d.y + __popc(d.x & (bit - 1u)) : kSentinel;
}

__global__ void lookup_kernel(const uint2* __restrict__ dict, const uint32_t* __restrict__ kmers, uint64_t n,
                              uint32_t universe, uint32_t* __restrict__ ids) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t km = kmers[i];
    ids[i] = km < universe ? dict_lookup(dict, km) : kSentinel;
  }
}

// K5, one k-mer slice per launch, G lanes per protein.  Sorted distinct k-mers of the slice
// (pk[ps+klo[r] .. ps+khi[r])) -> sorted repeated-k-mer ids appended in place behind what the
// earlier slices wrote (rowlen[r]); freq[id] += 1.  islo[r] records where this slice's ids
// start in the row (ishi on the last slice: the final row length), which is what the id-sliced
// passes below read.  The dictionary and freq slices of one launch stay L2-resident.
template <int G>
__global__ void __launch_bounds__(256)
    ids_freq_kernel(const uint2* __restrict__ dict, const uint32_t* __restrict__ pstart,
                    const uint32_t* __restrict__ klo, const uint32_t* __restrict__ khi, uint32_t n,
                    uint32_t* __restrict__ pk, uint32_t* __restrict__ rowlen, uint32_t* __restrict__ islo,
                    uint32_t* __restrict__ ishi_last, uint32_t* __restrict__ freq,
                    unsigned long long* __restrict__ nnz_total) {
  const uint32_t lane = lane_id(), gl = lane % G, gshift = (lane / G) * G;
  const uint32_t gmask = group_mask<G>();
  const uint32_t gg = (blockIdx.x * blockDim.x + threadIdx.x) / G, ng = (gridDim.x * blockDim.x) / G;
  unsigned long long tot = 0;
  for (uint32_t r = gg; r < n; r += ng) {
    const uint32_t i0 = klo[r], i1 = khi[r], ps = pstart[r];
    uint32_t base = rowlen[r];
    const uint32_t base0 = base;
    for (uint32_t c = i0; c < i1; c += G) {
      const uint32_t i = c + gl;
      uint32_t id = kSentinel;
      if (i < i1) id = dict_lookup(dict, pk[ps + i]);
      const uint32_t m = (__ballot_sync(gmask, id != kSentinel) >> gshift) & (G == 32 ? kFullMask : ((1u << (G & 31)) - 1u));
      if (id != kSentinel) {
        pk[ps + base + __popc(m & ((1u << gl) - 1u))] = id;
        atomicAdd(&freq[id], 1u);
      }
      base += __popc(m);
    }
    if (gl == 0) {
      islo[r] = base0;
      if (ishi_last) ishi_last[r] = base;
      if (base != base0) rowlen[r] = base;
      tot += base - base0;
    }
  }
  tot = warp_sum64(tot);
  if (lane == 0 && tot) atomicAdd(nnz_total, tot);
}

// postings fill, one id slice per launch: col[cursor[id]++] = r (cursor starts as a copy of
// colptr).  The cursor and postings slices of one launch stay L2-resident, so the scattered
// 4-byte writes merge into full lines before they go to HBM.
template <int G>
__global__ void __launch_bounds__(256)
    postings_fill_kernel(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ lo,
                         const uint32_t* __restrict__ hi, uint32_t n, const uint32_t* __restrict__ ids,
                         uint32_t* __restrict__ cursor, uint32_t* __restrict__ col) {
  const uint32_t gl = lane_id() % G;
  const uint32_t gg = (blockIdx.x * blockDim.x + threadIdx.x) / G, ng = (gridDim.x * blockDim.x) / G;
  for (uint32_t r = gg; r < n; r += ng) {
    const uint32_t i0 = lo[r], i1 = hi[r], ps = pstart[r];
    for (uint32_t i = i0 + gl; i < i1; i += G) {
      const uint32_t slot = atomicAdd(&cursor[ids[ps + i]], 1u);
      col[slot] = r;
    }
  }
}

// ---- postings sort: the atomic fill leaves every column in arrival order; the pair stage
// needs ascending protein rank.  Columns of <= 8 holders are sorted by one lane in registers,
// 9..32 by one warp (shuffle bitonic), longer ones are appended to work lists.
__device__ __forceinline__ void cswap(uint32_t& a, uint32_t& b) {
  const uint32_t lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}

__device__ __forceinline__ uint32_t warp_bitonic_reg(uint32_t v, uint32_t lane) {
#pragma unroll
  for (uint32_t size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      const uint32_t o = __shfl_xor_sync(kFullMask, v, stride);
      const bool up = (lane & size) == 0;
      const bool lower = (lane & stride) == 0;
      v = (lower == up) ? min(v, o) : max(v, o);
    }
  }
  return v;
}

constexpr uint32_t kColWarpMax = 1024;   // columns up to this length: one warp, shared memory
constexpr uint32_t kColBlockMax = 32768; // up to this: one CTA, shared memory; beyond: global

__global__ void __launch_bounds__(256)
    postings_sort_small_kernel(const uint32_t* __restrict__ colptr, uint32_t n_cols, uint32_t* __restrict__ col,
                               uint32_t* __restrict__ list_mid, uint32_t* __restrict__ list_big,
                               uint32_t* __restrict__ list_huge, uint32_t* __restrict__ list_counts) {
  const uint32_t lane = lane_id();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t c0 = gw * 32; c0 < n_cols; c0 += nw * 32) {
    const uint32_t c = c0 + lane;
    uint32_t lo = 0, f = 0;
    if (c < n_cols) {
      lo = colptr[c];
      f = colptr[c + 1] - lo;
    }
    if (f >= 2 && f <= 8) {
      uint32_t v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (uint32_t)i < f ? col[lo + i] : kSentinel;
      // 19-comparator network for 8 keys
      cswap(v[0], v[1]); cswap(v[2], v[3]); cswap(v[4], v[5]); cswap(v[6], v[7]);
      cswap(v[0], v[2]); cswap(v[1], v[3]); cswap(v[4], v[6]); cswap(v[5], v[7]);
      cswap(v[1], v[2]); cswap(v[5], v[6]); cswap(v[0], v[4]); cswap(v[3], v[7]);
      cswap(v[1], v[5]); cswap(v[2], v[6]);
      cswap(v[1], v[4]); cswap(v[3], v[6]);
      cswap(v[2], v[4]); cswap(v[3], v[5]);
      cswap(v[3], v[4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if ((uint32_t)i < f) col[lo + i] = v[i];
    } else if (f > 32) {
      uint32_t* lst = f <= kColWarpMax ? list_mid : (f <= kColBlockMax ? list_big : list_huge);
      const uint32_t which = f <= kColWarpMax ? 0 : (f <= kColBlockMax ? 1 : 2);
      lst[atomicAdd(&list_counts[which], 1u)] = c;
    }
    uint32_t m = __ballot_sync(kFullMask, f > 8 && f <= 32);
    while (m) {
      const uint32_t src = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t slo = __shfl_sync(kFullMask, lo, src), sf = __shfl_sync(kFullMask, f, src);
      uint32_t v = lane < sf ? col[slo + lane] : kSentinel;
      v = warp_bitonic_reg(v, lane);
      if (lane < sf) col[slo + lane] = v;
    }
  }
}

// mid columns: one warp each, bitonic in shared memory
__global__ void __launch_bounds__(128)
    postings_sort_mid_kernel(const uint32_t* __restrict__ colptr, const uint32_t* __restrict__ list,
                             const uint32_t* __restrict__ list_counts, uint32_t* __restrict__ col) {
  __shared__ uint32_t s_keys[4][kColWarpMax];
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
  const uint32_t n_list = list_counts[0];
  uint32_t* keys = s_keys[w];
  for (uint32_t li = blockIdx.x * 4 + w; li < n_list; li += gridDim.x * 4) {
    const uint32_t c = list[li];
    const uint32_t lo = colptr[c], f = colptr[c + 1] - lo;
    const uint32_t np2 = next_pow2_u32(f);
    for (uint32_t i = lane; i < np2; i += 32) keys[i] = i < f ? col[lo + i] : kSentinel;
    __syncwarp();
    warp_bitonic(keys, np2, lane);
    for (uint32_t i = lane; i < f; i += 32) col[lo + i] = keys[i];
    __syncwarp();
  }
}

// big columns: one CTA each; keys in dynamic shared memory, or in place in global memory
template <bool GLOBAL_KEYS>
__global__ void __launch_bounds__(512)
    postings_sort_big_kernel(const uint32_t* __restrict__ colptr, const uint32_t* __restrict__ list,
                             const uint32_t* __restrict__ list_counts, int which, uint32_t* __restrict__ col,
                             uint32_t* __restrict__ scratch, uint32_t scratch_stride) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  const uint32_t n_list = list_counts[which];
  for (uint32_t li = blockIdx.x; li < n_list; li += gridDim.x) {
    const uint32_t c = list[li];
    const uint32_t lo = colptr[c], f = colptr[c + 1] - lo;
    const uint32_t np2 = next_pow2_u32(f);
    uint32_t* keys = GLOBAL_KEYS ? scratch + (size_t)blockIdx.x * scratch_stride
                                 : reinterpret_cast<uint32_t*>(dyn_smem);
    for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) keys[i] = i < f ? col[lo + i] : kSentinel;
    __syncthreads();
    block_bitonic(keys, np2);
    for (uint32_t i = threadIdx.x; i < f; i += blockDim.x) col[lo + i] = keys[i];
    __syncthreads();
  }
}

// per-entry suffix ranges: for row r and each of its ids, the holders that come after r in the
// pair order (and, when only cross-class pairs are wanted, after r's whole class block) are
// col[suf.x .. suf.y) (or the single partner suf.x when suf.y is the sentinel).
// rowwork64[r] = sum of the range lengths = multi-edges row r accumulates.
// One id slice per launch like the fill (colptr and postings slices stay L2-resident).
template <int G>
__global__ void __launch_bounds__(256)
    suffix_ranges_kernel(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ lo,
                         const uint32_t* __restrict__ hi, uint32_t n, const uint32_t* __restrict__ ids,
                         const uint32_t* __restrict__ colptr, const uint32_t* __restrict__ col,
                         const uint32_t* __restrict__ first_after, const uint8_t* __restrict__ selfscore,
                         uint2* __restrict__ suf, uint8_t* __restrict__ sufss,
                         unsigned long long* __restrict__ rowwork64, uint32_t* __restrict__ rowinl,
                         uint32_t* __restrict__ rowmaxlen, uint32_t* __restrict__ pslo,
                         unsigned long long* __restrict__ work_total) {
  const uint32_t lane = lane_id(), gl = lane % G;
  const uint32_t gg = (blockIdx.x * blockDim.x + threadIdx.x) / G, ng = (gridDim.x * blockDim.x) / G;
  unsigned long long tot = 0;
  for (uint32_t r = gg; r < n; r += ng) {
    const uint32_t i0 = lo[r], i1 = hi[r], ps = pstart[r];
    // where this slice's multi-edges start in the row's materialised list (products_fill_kernel)
    if (pslo && gl == 0) pslo[r] = (uint32_t)min(rowwork64[r], 0xFFFFFFFFull);
    if (i0 >= i1) continue;
    const uint32_t target = first_after ? first_after[r] : r + 1;  // first rank that pairs with r
    unsigned long long work = 0;
    uint32_t n_inl = 0, max_len = 0;
    for (uint32_t i = i0 + gl; i < i1; i += G) {
      const uint32_t id = ids[ps + i];
      uint32_t a = colptr[id];
      const uint32_t end = colptr[id + 1];
      uint32_t partner = kSentinel;  // the holder at position a, when it is already in a register
      if (end - a <= 4u) {
        // short posting list (most k-mers): fetch it whole with independent loads instead of a
        // dependent binary search; padding with the sentinel keeps the comparison branch-free
        const uint32_t f = end - a;
        const uint32_t c0 = col[a];
        const uint32_t c1 = f > 1 ? col[a + 1] : kSentinel;
        const uint32_t c2 = f > 2 ? col[a + 2] : kSentinel;
        const uint32_t c3 = f > 3 ? col[a + 3] : kSentinel;
        const uint32_t below = (c0 < target) + (c1 < target) + (c2 < target) + (c3 < target);
        partner = below == 0 ? c0 : (below == 1 ? c1 : (below == 2 ? c2 : c3));
        a += below;
      } else {
        uint32_t b = end;
        while (a < b) {  // lower_bound(col[a..b), target)
          const uint32_t mid = (a + b) >> 1;
          if (col[mid] < target) a = mid + 1; else b = mid;
        }
        if (end - a == 1u) partner = col[a];
      }
      // a single partner is stored inline ({rank, sentinel}): the pair stage then needs no
      // postings gather for it (a 4-byte read there costs a whole random 32-byte sector)
      suf[ps + i] = end - a == 1u ? make_uint2(partner, kSentinel) : make_uint2(a, end);
      if (sufss) sufss[ps + i] = selfscore[id];  // BLOSUM62 self-score of the entry's k-mer (K9 fused into K7)
      work += end - a;
      n_inl += end - a == 1u;
      max_len = max(max_len, end - a == 1u ? 0u : end - a);
    }
    work = group_sum64<G>(work);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      n_inl += __shfl_xor_sync(group_mask<G>(), n_inl, o);
      max_len = max(max_len, __shfl_xor_sync(group_mask<G>(), max_len, o));
    }
    if (gl == 0) {
      rowwork64[r] += work;
      rowinl[r] += n_inl;
      rowmaxlen[r] = max(rowmaxlen[r], max_len);
      tot += work;
    }
  }
  tot = warp_sum64(tot);
  if (lane == 0 && tot) atomicAdd(work_total, tot);
}

// Materialised multi-edge lists ("plist"): row r's partners, one u32 per multi-edge, in entry order
// at plist[rowbase[r] ..], plus the entry's BLOSUM self-score per multi-edge (pss).  Filled slice
// by slice like the suffix pass (the postings slice stays L2-resident); the pair kernel then
// streams exactly 4 bytes per multi-edge with coalesced loads instead of gathering postings.
template <int G>
__global__ void __launch_bounds__(256)
    products_fill_kernel(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ lo,
                         const uint32_t* __restrict__ hi, uint32_t n, const uint2* __restrict__ suf,
                         const uint8_t* __restrict__ sufss, const uint32_t* __restrict__ col,
                         const unsigned long long* __restrict__ rowbase, const unsigned long long* __restrict__ rowwork64,
                         const uint32_t* __restrict__ pslo, uint32_t* __restrict__ plist, uint8_t* __restrict__ pss) {
  const uint32_t lane = lane_id(), gl = lane % G, gshift = (lane / G) * G;
  const uint32_t gmask = group_mask<G>();
  const uint32_t gg = (blockIdx.x * blockDim.x + threadIdx.x) / G, ng = (gridDim.x * blockDim.x) / G;
  for (uint32_t r = gg; r < n; r += ng) {
    const uint32_t i0 = lo[r], i1 = hi[r], ps = pstart[r];
    if (i0 >= i1 || rowwork64[r] > 0xFFFFFFFFull) continue;  // monstrous rows are never streamed
    unsigned long long dst0 = rowbase[r] + pslo[r];
    for (uint32_t c = i0; c < i1; c += G) {
      const uint32_t i = c + gl;
      uint2 e = make_uint2(0, 0);
      uint32_t ss = 0;
      if (i < i1) {
        e = suf[ps + i];
        if (pss) ss = sufss[ps + i];
      }
      const bool inl = e.y == kSentinel;
      const uint32_t len = inl ? 1u : e.y - e.x;
      // exclusive scan of len over the G lanes of the group
      uint32_t incl = len;
#pragma unroll
      for (int o = 1; o < G; o <<= 1) {
        const uint32_t t = __shfl_up_sync(gmask, incl, o, G);
        if (gl >= (uint32_t)o) incl += t;
      }
      const uint32_t total = __shfl_sync(gmask, incl, gshift + G - 1);
      const unsigned long long dst = dst0 + incl - len;
      if (inl) {
        plist[dst] = e.x;
        if (pss) pss[dst] = (uint8_t)ss;
      } else if (len < 32u) {
        for (uint32_t j = 0; j < len; ++j) {
          plist[dst + j] = col[e.x + j];
          if (pss) pss[dst + j] = (uint8_t)ss;
        }
      }
      // long posting suffixes: the whole group copies them, one after the other
      uint32_t m = (__ballot_sync(gmask, !inl && len >= 32u) >> gshift) & (G == 32 ? kFullMask : ((1u << (G & 31)) - 1u));
      while (m) {
        const uint32_t src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t sx = __shfl_sync(gmask, e.x, gshift + src), sl = __shfl_sync(gmask, len, gshift + src);
        const uint32_t sss = __shfl_sync(gmask, ss, gshift + src);
        const unsigned long long sd =
            (unsigned long long)__shfl_sync(gmask, (uint32_t)(dst >> 32), gshift + src) << 32 |
            __shfl_sync(gmask, (uint32_t)dst, gshift + src);
        for (uint32_t j = gl; j < sl; j += G) {
          plist[sd + j] = col[sx + j];
          if (pss) pss[sd + j] = (uint8_t)sss;
        }
      }
      dst0 += total;
    }
  }
}

struct U64In {
  const unsigned long long* p;
  __device__ unsigned long long operator()(uint64_t i) const { return p[i] > 0xFFFFFFFFull ? 0ull : p[i]; }
};

// clamp the 64-bit per-row work to the 32-bit value the row classifier uses
__global__ void clamp_rowwork_kernel(const unsigned long long* __restrict__ w64, uint32_t n,
                                     uint32_t* __restrict__ w32) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) w32[r] = w64[r] > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)w64[r];
}

// Sum over ids of f(f-1)/2: "Number of total edges", src/graph/mod.rs:44-51
__global__ void multi_edge_total_kernel(const uint32_t* __restrict__ freq, uint32_t n,
                                        unsigned long long* __restrict__ total) {
  unsigned long long s = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long f = freq[i];
    s += f * (f - 1) / 2;
  }
  s = warp_sum64(s);
  if (lane_id() == 0 && s) atomicAdd(total, s);
}

// compact the gapped id rows into a dense CSR for host readback
__global__ void __launch_bounds__(256)
    compact_rows_kernel(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ rowlen,
                        const unsigned long long* __restrict__ row_off, const uint32_t* __restrict__ rank_of,
                        uint32_t n, const uint32_t* __restrict__ ids, uint32_t* __restrict__ out) {
  const uint32_t lane = lane_id();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t p = gw; p < n; p += nw) {  // p = protein in input order
    const uint32_t r = rank_of ? rank_of[p] : p;
    const uint32_t nl = rowlen[r], ps = pstart[r];
    const unsigned long long o = row_off[p];
    for (uint32_t i = lane; i < nl; i += 32) out[o + i] = ids[ps + i];
  }
}

}  // namespace kc
