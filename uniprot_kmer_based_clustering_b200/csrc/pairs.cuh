// pairs.cuh — K7/K8/K9: all-pairs shared-k-mer counts, threshold, edge emission, BLOSUM.
//
// Replaces (reference root relative):
//   Graph::new + update_graph_edges   src/graph/mod.rs:39-193, src/graph/vertex.rs:59-140
//   remove_uninteresting_edges        src/graph/mod.rs:549-697
//   combine_edges                     src/graph/mod.rs:322-546, src/graph/edge.rs:56-85
//   threshold                         src/graph/mod.rs:242
// The reference materialises one heap object per (k-mer, protein pair) and then filters and
// groups them.  Here S = A*A^T (upper triangle in the pair order) is accumulated row by row
// with ON-CHIP counters: row r walks, for each of its ids, the postings suffix
// col[suf.x..suf.y) (holders that pair with r) and bumps one shared-memory counter per
// holder.  Multi-edges are never written anywhere; only pairs over the threshold leave the SM.
//
// Three accumulator shapes, picked per row from its exact multi-edge count (rowwork) and
// the number of candidate partners (span):
//   hash  (one warp per row)  : open addressing, H = 256..4096 slots, for sparse rows
//   hash  (one CTA per row)   : H = 16384 slots
//   dense (one CTA per row)   : one u16/u32 counter per candidate partner, column-blocked
//                               passes when the span exceeds the shared-memory block
#pragma once
#include "common.cuh"

namespace kc {

struct EdgeSink {
  uint4* buf;                   // {rank_a, rank_b, count, blosum}
  unsigned long long* cursor;   // edges wanted so far (may exceed cap: caller grows + retries)
  unsigned long long cap;
  uint32_t threshold;
  uint32_t unscored;            // blosum word written by kernels that do not accumulate scores:
                                // kUnscored when a score is wanted (edge_blosum_kernel fills it), else 0
};
constexpr uint32_t kUnscored = 0x80000000u;
constexpr uint32_t kScoreShift = 12;  // scored tables: value = score << 12 | count (rows of < 4096 ids)

struct PairCounters {            // device-side totals (u64 each)
  unsigned long long n_pairs;    // distinct (row, partner) pairs with count >= 1
  unsigned long long n_edges;    // pairs over the threshold
  unsigned long long sum_count;  // sum of their counts
  unsigned long long n_multi;    // sum of all counts = multi-edges accumulated
};

__device__ __forceinline__ void emit_edge(bool pred, uint32_t ra, uint32_t rb, uint32_t count,
                                          const EdgeSink& sink) {
  const uint32_t m = __ballot_sync(kFullMask, pred);
  if (!m) return;
  unsigned long long base = 0;
  const uint32_t leader = __ffs(m) - 1;
  if (lane_id() == leader) base = atomicAdd(sink.cursor, (unsigned long long)__popc(m));
  base = __shfl_sync(kFullMask, base, leader);
  const unsigned long long idx = base + __popc(m & lanemask_lt());
  if (pred && idx < sink.cap) sink.buf[idx] = make_uint4(ra, rb, count, sink.unscored);
}

// Per-warp staging of emitted edges in shared memory: one global atomic per flush instead of
// one per emitting instruction (ncu: the atomic's return latency was 30 % of all stall samples).
constexpr uint32_t kStageEdges = 170;              // 170 x {row, partner, count} = 2040 B
constexpr uint32_t kStageWords = 512;
struct EdgeStage {
  uint32_t* buf;
  uint32_t cnt;  // warp-uniform
};
__device__ __forceinline__ void stage_flush(EdgeStage& st, const EdgeSink& sink) {
  if (st.cnt == 0) return;
  __syncwarp();
  unsigned long long base = 0;
  if (lane_id() == 0) base = atomicAdd(sink.cursor, (unsigned long long)st.cnt);
  base = __shfl_sync(kFullMask, base, 0);
  for (uint32_t i = lane_id(); i < st.cnt; i += 32) {
    const unsigned long long idx = base + i;
    if (idx < sink.cap)
      sink.buf[idx] = make_uint4(st.buf[3 * i], st.buf[3 * i + 1], st.buf[3 * i + 2], sink.unscored);
  }
  __syncwarp();
  st.cnt = 0;
}
// all 32 lanes call it; the caller guarantees room for 32 more edges
__device__ __forceinline__ void stage_push(EdgeStage& st, bool pred, uint32_t r, uint32_t b, uint32_t c) {
  const uint32_t m = __ballot_sync(kFullMask, pred);
  if (!m) return;
  if (pred) {
    const uint32_t pos = st.cnt + __popc(m & lanemask_lt());
    st.buf[3 * pos] = r;
    st.buf[3 * pos + 1] = b;
    st.buf[3 * pos + 2] = c;
  }
  st.cnt += __popc(m);
}

// scored variant: {row, partner, count, score}
constexpr uint32_t kStageEdges4 = 128;
__device__ __forceinline__ void stage4_flush(EdgeStage& st, const EdgeSink& sink) {
  if (st.cnt == 0) return;
  __syncwarp();
  unsigned long long base = 0;
  if (lane_id() == 0) base = atomicAdd(sink.cursor, (unsigned long long)st.cnt);
  base = __shfl_sync(kFullMask, base, 0);
  for (uint32_t i = lane_id(); i < st.cnt; i += 32) {
    const unsigned long long idx = base + i;
    if (idx < sink.cap) sink.buf[idx] = *reinterpret_cast<const uint4*>(st.buf + 4 * i);
  }
  __syncwarp();
  st.cnt = 0;
}
__device__ __forceinline__ void stage4_push(EdgeStage& st, bool pred, uint32_t r, uint32_t b, uint32_t c,
                                            uint32_t score) {
  const uint32_t m = __ballot_sync(kFullMask, pred);
  if (!m) return;
  if (pred) {
    const uint32_t pos = st.cnt + __popc(m & lanemask_lt());
    *reinterpret_cast<uint4*>(st.buf + 4 * pos) = make_uint4(r, b, c, score);
  }
  st.cnt += __popc(m);
}

// row classes
enum : uint8_t {
  kBinSkip = 0,
  kBinPack8 = 1,    // packed hash, one warp per row, 256 slots   (U <= 128)
  kBinPack9 = 2,    //                                512          (U <= 256)
  kBinPack10 = 3,   //                                1024         (U <= 512)
  kBinPack11 = 4,   //                                2048         (U <= 1024)
  kBinPack12 = 5,   // packed hash, 4 warps per row,  4096         (U <= 2048)
  kBinPack13 = 6,   // packed hash, 8 warps per row,  8192         (U <= 4096)
  kBinPack14 = 7,   // packed hash, 8 warps per row,  16384        (U <= 8192)
  kBinWide = 8,     // key/count in separate words (rows whose counts do not fit a packed slot)
  kBinDense = 9,
  kBinMain = 10,    // packed hash, one warp per row, table sized per row from an optimistic estimate
  kBinMainS = 11,   // scored main kernel with half-size tables: rows whose partner bound fits them for sure
  kNumBins = 12,
  kBinRetry = 16    // added to a safe bin: the optimistic table of the main kernel overflowed
};
constexpr uint32_t kMainLogHMax = 10;  // 1024 slots = 4 KB per warp
constexpr uint32_t kMainCap = 640;     // distinct partners a row may collect in the main kernel (load 0.625)
constexpr uint32_t kMainSLogH = 9;     // the small variant of the scored main kernel: 512 slots,
constexpr uint32_t kMainSCap = 320;    // rows with an EXACT partner bound U <= 320 (they cannot overflow)

// bounds[0..1] = the shard's row range (device memory: no host round trip).
// count_bits = bits left for the counter in a packed slot (32 - bits of a protein rank).
__global__ void __launch_bounds__(256)
    classify_rows_kernel(const uint32_t* __restrict__ rowwork, const uint32_t* __restrict__ rowlen,
                         const uint32_t* __restrict__ rowinl, const uint32_t* __restrict__ rowmaxlen,
                         const uint32_t* __restrict__ first_after, uint32_t n, const uint32_t* __restrict__ bounds,
                         uint32_t dense_single_pass_cols, uint32_t count_bits, uint8_t* __restrict__ rowbin,
                         uint8_t* __restrict__ rowsafe, uint8_t* __restrict__ rowlogh,
                         uint32_t* __restrict__ bin_counts, RowOwner owner, int exact_main, bool small_main) {
  __shared__ uint32_t s_cnt[16];
  if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t row_lo = bounds[0], row_hi = bounds[1];
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) {
    uint8_t bin = kBinSkip;
    const uint32_t P = rowwork[r];
    // the rows this call scores: the shard's row range, or (sharded index) the rows this rank owns
    if (P != 0 && (owner.bin_owner ? owner.mine(r) : (r >= row_lo && r < row_hi))) {
      const uint32_t target = first_after ? first_after[r] : r + 1;
      const uint32_t span = n - target;
      const uint32_t U = min(P, span);  // upper bound on distinct partners
      const bool packable = (rowlen[r] >> count_bits) == 0;  // a count never exceeds the row length
      if (span <= dense_single_pass_cols && span <= 4u * U) bin = kBinDense;
      else if (U > 8192) bin = kBinDense;
      else if (!packable) bin = kBinWide;
      else if (U <= 128) bin = kBinPack8;
      else if (U <= 256) bin = kBinPack9;
      else if (U <= 512) bin = kBinPack10;
      else if (U <= 1024) bin = kBinPack11;
      else if (U <= 2048) bin = kBinPack12;
      else if (U <= 4096) bin = kBinPack13;
      else bin = kBinPack14;
      uint8_t safe = bin;
      if (bin >= kBinPack8 && bin <= kBinPack14) {
        // U counts every multi-edge as a new partner, but related proteins meet the same partners
        // again and again.  The main kernel counts distinct partners exactly as it goes and gives
        // a row up when they pass kMainCap (a bounded loss: at most kMainCap insertions), so a row
        // is tried there unless a lower bound on its partners (inline partners are distinct; the
        // longest suffix holds distinct partners) already comes close to the cap.
        // exact_main (partitioned index): the bin-local partners, the ones that repeat, are scored on
        // the tiles; what is left are mostly distinct partners, so U is a good estimate and a row
        // beyond the cap goes straight to its safely sized kernel instead of failing here first.
        const uint32_t lower = max(rowinl[r], rowmaxlen[r]);
        // exact_main is off when most multi-edges are left to the hash kernels (related proteins are NOT
        // neighbours in the input, the bin-local tiles see little: the family partners come through here with
        // their multiplicity and U overshoots); a row sent to the unscored packed kernels costs a list
        // intersection per emitted edge afterwards (8.5 ms on the shuffled 1 M set, profiles/r2_history.md).
        // exact_main: 0 = the table build (no tiles: always optimistic), 1 = trust U, 2 = little went to the
        // tiles: optimistic for rows whose U is within a few times the cap (rows with thousands of multi-edges,
        // the k = 5 regime, really have that many partners and go straight to their safely sized kernel).
        const bool optimistic = exact_main == 0 || (exact_main == 2 && U <= 4u * kMainCap);
        if ((U <= kMainCap || (optimistic && lower <= kMainCap / 2)) &&
            rowlen[r] < (1u << kScoreShift) &&
            P != 0xFFFFFFFFu) {
          bin = small_main && U <= kMainSCap ? kBinMainS : kBinMain;
          rowlogh[r] = (uint8_t)kMainLogHMax;
        }
      }
      rowsafe[r] = safe;
      atomicAdd(&s_cnt[bin], 1u);
    }
    rowbin[r] = bin;
  }
  __syncthreads();
  if (threadIdx.x < 16 && s_cnt[threadIdx.x]) atomicAdd(&bin_counts[threadIdx.x], s_cnt[threadIdx.x]);
}

template <int LOG_H>
__device__ __forceinline__ void hash_bump(uint32_t* keys, uint32_t* cnt, uint32_t b) {
  constexpr uint32_t H = 1u << LOG_H;
  uint32_t h = (b * 2654435761u) >> (32 - LOG_H);
  for (;;) {
    const uint32_t k = keys[h];
    if (k == b) break;
    if (k == kSentinel) {
      const uint32_t old = atomicCAS(&keys[h], kSentinel, b);
      if (old == kSentinel || old == b) break;
    }
    h = (h + 1u) & (H - 1u);
  }
  atomicAdd(&cnt[h], 1u);
}

// walk one 32-entry chunk of a row: long postings suffixes are read by the whole warp
// (coalesced), short ones by the lane that owns the entry
template <class Bump>
__device__ __forceinline__ void walk_chunk(const uint32_t* __restrict__ col, uint2 e, Bump bump) {
  const uint32_t lane = lane_id();
  if (e.y == kSentinel) {  // single partner stored inline
    bump(e.x);
    e = make_uint2(0, 0);
  }
  const uint32_t len = e.y - e.x;
  uint32_t m = __ballot_sync(kFullMask, len >= 16u);
  while (m) {
    const uint32_t src = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t s = __shfl_sync(kFullMask, e.x, src), t = __shfl_sync(kFullMask, e.y, src);
    for (uint32_t j = s + lane; j < t; j += 32) bump(col[j]);
  }
  if (len < 16u)
    for (uint32_t j = e.x; j < e.y; ++j) bump(col[j]);
}

// ---------------------------------------------------------------------------------------
// hash accumulators.  GROUP_WARPS warps share one table and one row; a CTA holds
// CTA_WARPS / GROUP_WARPS groups.  GROUP_WARPS is 1 (warp per row) or CTA_WARPS (CTA per row).
// ---------------------------------------------------------------------------------------
template <int LOG_H, int GROUP_WARPS, int CTA_WARPS>
__global__ void __launch_bounds__(CTA_WARPS * 32)
    pairs_hash_kernel(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ rowlen,
                      const uint2* __restrict__ suf, const uint32_t* __restrict__ col,
                      const uint8_t* __restrict__ rowbin, uint8_t my_bin, uint32_t n,
                      uint32_t* __restrict__ row_cursor, const uint32_t* __restrict__ bin_counts, EdgeSink sink,
                      PairCounters* __restrict__ counters) {
  static_assert(GROUP_WARPS == 1 || GROUP_WARPS == CTA_WARPS, "group = warp or CTA");
  if (bin_counts[my_bin] == 0) return;  // nothing in this bin
  constexpr uint32_t H = 1u << LOG_H;
  constexpr int GROUPS = CTA_WARPS / GROUP_WARPS;
  constexpr uint32_t GSIZE = GROUP_WARPS * 32;
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  __shared__ uint32_t s_base;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t group = warp / GROUP_WARPS, gwarp = warp % GROUP_WARPS;
  const uint32_t gtid = gwarp * 32 + lane;
  uint32_t* keys = reinterpret_cast<uint32_t*>(dyn_smem) + (size_t)group * 2 * H;
  uint32_t* cnt = keys + H;
  (void)GROUPS;
  unsigned long long n_pairs = 0, n_edges = 0, sum_count = 0, n_multi = 0;

  auto gsync = [&]() {
    if (GROUP_WARPS == 1) __syncwarp(); else __syncthreads();
  };

  for (;;) {
    uint32_t base;
    if (GROUP_WARPS == 1) {
      base = 0;
      if (lane == 0) base = atomicAdd(row_cursor, 32u);
      base = __shfl_sync(kFullMask, base, 0);
    } else {
      __syncthreads();
      if (threadIdx.x == 0) s_base = atomicAdd(row_cursor, 32u);
      __syncthreads();
      base = s_base;
    }
    if (base >= n) break;
    uint32_t todo = __ballot_sync(kFullMask, base + lane < n && rowbin[base + lane] == my_bin);
    while (todo) {
      const uint32_t r = base + __ffs(todo) - 1;
      todo &= todo - 1;
      // clear the table
      for (uint32_t i = gtid * 4; i < H; i += GSIZE * 4) {
        *reinterpret_cast<uint4*>(keys + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
        *reinterpret_cast<uint4*>(cnt + i) = make_uint4(0, 0, 0, 0);
      }
      gsync();
      const uint32_t nl = rowlen[r], ps = pstart[r];
      for (uint32_t c = gwarp * 32; c < nl; c += GSIZE) {
        const uint2 e = c + lane < nl ? ld_stream_u32x2(suf + ps + c + lane) : make_uint2(0, 0);
        walk_chunk(col, e, [&](uint32_t b) { hash_bump<LOG_H>(keys, cnt, b); });
      }
      gsync();
      for (uint32_t i = gtid; i < H; i += GSIZE) {
        const uint32_t b = keys[i];
        const uint32_t c = b != kSentinel ? cnt[i] : 0u;
        const bool out = c > sink.threshold;
        n_pairs += c != 0;
        n_multi += c;
        n_edges += out;
        sum_count += out ? c : 0u;
        emit_edge(out, r, b, c, sink);
      }
      gsync();
    }
  }
  n_pairs = warp_sum64(n_pairs);
  n_edges = warp_sum64(n_edges);
  sum_count = warp_sum64(sum_count);
  n_multi = warp_sum64(n_multi);
  if (lane == 0) {
    if (n_multi) atomicAdd(&counters->n_multi, n_multi);
    if (n_pairs) atomicAdd(&counters->n_pairs, n_pairs);
    if (n_edges) atomicAdd(&counters->n_edges, n_edges);
    if (sum_count) atomicAdd(&counters->sum_count, sum_count);
  }
}

// ---------------------------------------------------------------------------------------
// packed hash accumulators (the main path).  One 32-bit slot = protein rank << count_bits |
// count, so a 2048-slot table is 8 KB and 20+ rows are resident per SM.  The walk over a
// row's postings suffixes is flattened: every lane first writes the postings indices of its
// short suffixes into a per-warp index list, then the warp strides over that list with four
// independent loads in flight per lane (no lane-serial dependent loads, no idle lanes).
// The read-out pass clears the table for the next row.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kIdxPerWarp = 512;  // >= 32 lanes x 15 postings

// scored main-kernel variant: keys and values in separate words; value = score << 12 | count, so
// the count and the BLOSUM self-score sum of a pair grow with ONE shared-memory atomic
__device__ __forceinline__ void scored_bump_dirty(uint32_t* key, uint32_t* val, uint32_t mask, uint32_t log_h,
                                                  uint32_t b, uint32_t inc, uint32_t* dirty_cnt, uint16_t* dirty,
                                                  uint32_t cap, bool& full) {
  if (full) return;
  uint32_t h = (b * 2654435761u) >> (32u - log_h);
  for (;;) {
    const uint32_t k = key[h];
    if (k == b) {
      atomicAdd(&val[h], inc);
      return;
    }
    if (k == kSentinel) {
      const uint32_t old = atomicCAS(&key[h], kSentinel, b);
      if (old == kSentinel) {
        atomicAdd(&val[h], inc);
        const uint32_t pos = atomicAdd(dirty_cnt, 1u);
        if (pos < cap) dirty[pos] = (uint16_t)h;
        else full = true;
        return;
      }
      if (old == b) {
        atomicAdd(&val[h], inc);
        return;
      }
    }
    h = (h + 1u) & mask;
  }
}

// scored main kernel: bump, and return the slot when the partner is NEW (else the sentinel); the
// caller appends new slots to the dirty list with one ballot per round (the list length is a
// warp-uniform register: no shared-memory atomic on a single hot counter)
__device__ __forceinline__ uint32_t scored_bump_new(uint32_t* key, uint32_t* val, uint32_t mask, uint32_t log_h,
                                                    uint32_t b, uint32_t inc) {
  uint32_t h = (b * 2654435761u) >> (32u - log_h);
  for (;;) {
    const uint32_t k = key[h];
    if (k == b) {
      atomicAdd(&val[h], inc);
      return kSentinel;
    }
    if (k == kSentinel) {
      const uint32_t old = atomicCAS(&key[h], kSentinel, b);
      if (old == kSentinel) {
        atomicAdd(&val[h], inc);
        return h;
      }
      if (old == b) {
        atomicAdd(&val[h], inc);
        return kSentinel;
      }
    }
    h = (h + 1u) & mask;
  }
}

// main-kernel variant: every new key's slot is appended to the warp's dirty list (so the read-out
// touches only occupied slots and the distinct-partner count is exact); gives up beyond `cap`
__device__ __forceinline__ void packed_bump_dirty(uint32_t* tab, uint32_t mask, uint32_t log_h, uint32_t cb,
                                                  uint32_t b, uint32_t* dirty_cnt, uint16_t* dirty, uint32_t cap,
                                                  bool& full) {
  if (full) return;
  uint32_t h = (b * 2654435761u) >> (32u - log_h);
  for (;;) {
    const uint32_t s = tab[h];
    if ((s >> cb) == b) {  // an empty slot never matches: its key bits are all ones
      atomicAdd(&tab[h], 1u);
      return;
    }
    if (s == kSentinel) {
      const uint32_t old = atomicCAS(&tab[h], kSentinel, (b << cb) | 1u);
      if (old == kSentinel) {
        const uint32_t pos = atomicAdd(dirty_cnt, 1u);
        if (pos < cap) dirty[pos] = (uint16_t)h;
        else full = true;
        return;
      }
      if ((old >> cb) == b) {
        atomicAdd(&tab[h], 1u);
        return;
      }
    }
    h = (h + 1u) & mask;
  }
}

// bounded variant for optimistically sized tables: counts new keys, gives up when the table is full
__device__ __forceinline__ void packed_bump_checked(uint32_t* tab, uint32_t mask, uint32_t log_h, uint32_t cb,
                                                    uint32_t b, uint32_t& inserted, bool& full) {
  if (full) return;
  uint32_t h = (b * 2654435761u) >> (32u - log_h);
  {  // fast path: the partner is already in its home slot (an empty slot never matches: its key bits are all ones)
    const uint32_t s0 = tab[h];
    if ((s0 >> cb) == b) {
      atomicAdd(&tab[h], 1u);
      return;
    }
  }
  for (uint32_t probe = 0; probe < 128u && probe <= mask; ++probe) {  // long clusters happen at load 0.6
    const uint32_t s = tab[h];
    if ((s >> cb) == b && s != kSentinel) {
      atomicAdd(&tab[h], 1u);
      return;
    }
    if (s == kSentinel) {
      const uint32_t old = atomicCAS(&tab[h], kSentinel, (b << cb) | 1u);
      if (old == kSentinel) {
        ++inserted;
        return;
      }
      if ((old >> cb) == b) {
        atomicAdd(&tab[h], 1u);
        return;
      }
    }
    h = (h + 1u) & mask;
  }
  full = true;
}

__device__ __forceinline__ void packed_bump(uint32_t* tab, uint32_t mask, uint32_t log_h, uint32_t cb,
                                            uint32_t b) {
  uint32_t h = (b * 2654435761u) >> (32u - log_h);
  for (;;) {
    const uint32_t s = tab[h];
    if ((s >> cb) == b && s != kSentinel) {
      atomicAdd(&tab[h], 1u);
      return;
    }
    if (s == kSentinel) {
      const uint32_t old = atomicCAS(&tab[h], kSentinel, (b << cb) | 1u);
      if (old == kSentinel) return;
      if ((old >> cb) == b) {
        atomicAdd(&tab[h], 1u);
        return;
      }
    }
    h = (h + 1u) & mask;
  }
}

// `stop` (optional): lane-local "give up" flag of an optimistically sized table; the loops have
// warp-uniform trip counts so that the whole warp can leave as soon as any lane raises it.
template <class Bump>
__device__ __forceinline__ void walk_chunk_flat(const uint32_t* __restrict__ col, uint2 e, uint32_t* idx,
                                                Bump bump, const bool* stop = nullptr) {
  const uint32_t lane = lane_id();
  if (e.y == kSentinel) {  // single partner stored inline
    bump(e.x);
    e = make_uint2(0, 0);
  }
  const uint32_t len = e.y - e.x;
  // long suffixes: the whole warp reads 32 consecutive postings at a time
  uint32_t m = __ballot_sync(kFullMask, len >= 16u);
  while (m) {
    const uint32_t src = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t s = __shfl_sync(kFullMask, e.x, src), t = __shfl_sync(kFullMask, e.y, src);
    for (uint32_t j0 = s; j0 < t; j0 += 128) {
      const uint32_t j = j0 + lane;
      const uint32_t b0 = j < t ? col[j] : kSentinel;
      const uint32_t b1 = j + 32 < t ? col[j + 32] : kSentinel;
      const uint32_t b2 = j + 64 < t ? col[j + 64] : kSentinel;
      const uint32_t b3 = j + 96 < t ? col[j + 96] : kSentinel;
      if (b0 != kSentinel) bump(b0);
      if (b1 != kSentinel) bump(b1);
      if (b2 != kSentinel) bump(b2);
      if (b3 != kSentinel) bump(b3);
      if (stop && __any_sync(kFullMask, *stop)) return;
    }
  }
  // short suffixes: flatten into the per-warp index list
  const uint32_t slen = len < 16u ? len : 0u;
  const uint32_t incl = warp_scan_incl(slen);
  const uint32_t total = __shfl_sync(kFullMask, incl, 31);
  if (total == 0) return;
  uint32_t w = incl - slen;
  for (uint32_t j = e.x; j < e.x + slen; ++j) idx[w++] = j;
  __syncwarp();
  for (uint32_t t00 = 0; t00 < total; t00 += 128) {
    const uint32_t t0 = t00 + lane;
    const uint32_t b0 = t0 < total ? col[idx[t0]] : kSentinel;
    const uint32_t b1 = t0 + 32 < total ? col[idx[t0 + 32]] : kSentinel;
    const uint32_t b2 = t0 + 64 < total ? col[idx[t0 + 64]] : kSentinel;
    const uint32_t b3 = t0 + 96 < total ? col[idx[t0 + 96]] : kSentinel;
    if (b0 != kSentinel) bump(b0);
    if (b1 != kSentinel) bump(b1);
    if (b2 != kSentinel) bump(b2);
    if (b3 != kSentinel) bump(b3);
    if (stop && __any_sync(kFullMask, *stop)) break;
  }
  __syncwarp();
}

template <int LOG_H, int GROUP_WARPS, int CTA_WARPS>
__global__ void __launch_bounds__(CTA_WARPS * 32)
    pairs_packed_kernel(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ rowlen,
                        const uint2* __restrict__ suf, const uint32_t* __restrict__ col,
                        const uint8_t* __restrict__ rowbin, uint8_t my_bin, uint32_t n, uint32_t count_bits,
                        uint32_t* __restrict__ row_cursor, const uint32_t* __restrict__ bin_counts, EdgeSink sink,
                        PairCounters* __restrict__ counters) {
  static_assert(GROUP_WARPS == 1 || GROUP_WARPS == CTA_WARPS, "group = warp or CTA");
  if (bin_counts[my_bin] == 0) return;  // nothing in this bin
  constexpr uint32_t H = 1u << LOG_H;
  constexpr uint32_t GROUPS = CTA_WARPS / GROUP_WARPS;
  constexpr uint32_t GSIZE = GROUP_WARPS * 32;
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  __shared__ uint32_t s_base;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t group = warp / GROUP_WARPS, gwarp = warp % GROUP_WARPS;
  const uint32_t gtid = gwarp * 32 + lane;
  uint32_t* tab = reinterpret_cast<uint32_t*>(dyn_smem) + (size_t)group * H;
  uint32_t* idx = reinterpret_cast<uint32_t*>(dyn_smem) + (size_t)GROUPS * H + (size_t)warp * kIdxPerWarp;
  EdgeStage stage{reinterpret_cast<uint32_t*>(dyn_smem) + (size_t)GROUPS * H + (size_t)CTA_WARPS * kIdxPerWarp +
                      (size_t)warp * kStageWords,
                  0u};
  const uint32_t cb = count_bits, cmask = (1u << count_bits) - 1u;
  unsigned long long n_pairs = 0, n_edges = 0, sum_count = 0, n_multi = 0;

  auto gsync = [&]() {
    if (GROUP_WARPS == 1) __syncwarp(); else __syncthreads();
  };
  for (uint32_t i = gtid * 4; i < H; i += GSIZE * 4)
    *reinterpret_cast<uint4*>(tab + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
  gsync();

  for (;;) {
    uint32_t base;
    if (GROUP_WARPS == 1) {
      base = 0;
      if (lane == 0) base = atomicAdd(row_cursor, 32u);
      base = __shfl_sync(kFullMask, base, 0);
    } else {
      __syncthreads();
      if (threadIdx.x == 0) s_base = atomicAdd(row_cursor, 32u);
      __syncthreads();
      base = s_base;
    }
    if (base >= n) break;
    uint32_t todo = __ballot_sync(kFullMask, base + lane < n && rowbin[base + lane] == my_bin);
    while (todo) {
      const uint32_t r = base + __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t nl = rowlen[r], ps = pstart[r];
      uint32_t c = gwarp * 32;
      uint2 e_next = c + lane < nl ? ld_stream_u32x2(suf + ps + c + lane) : make_uint2(0, 0);
      for (; c < nl; c += GSIZE) {
        const uint2 e = e_next;
        const uint32_t cn = c + GSIZE;
        e_next = cn + lane < nl ? ld_stream_u32x2(suf + ps + cn + lane) : make_uint2(0, 0);
        walk_chunk_flat(col, e, idx, [&](uint32_t b) { packed_bump(tab, H - 1u, LOG_H, cb, b); });
      }
      gsync();
      // read out and clear
      for (uint32_t i = gtid * 4; i < H; i += GSIZE * 4) {
        const uint4 v = *reinterpret_cast<uint4*>(tab + i);
        *reinterpret_cast<uint4*>(tab + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
        const uint32_t sv[4] = {v.x, v.y, v.z, v.w};
        bool any_out = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t cq = sv[q] != kSentinel ? (sv[q] & cmask) : 0u;
          n_pairs += cq != 0;
          n_multi += cq;
          any_out |= cq > sink.threshold;
        }
        if (__any_sync(kFullMask, any_out)) {
          if (stage.cnt + 128u > kStageEdges) stage_flush(stage, sink);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t cq = sv[q] != kSentinel ? (sv[q] & cmask) : 0u;
            const bool out = cq > sink.threshold;
            n_edges += out;
            sum_count += out ? cq : 0u;
            stage_push(stage, out, r, sv[q] >> cb, cq);
          }
        }
      }
      gsync();
    }
  }
  stage_flush(stage, sink);
  n_pairs = warp_sum64(n_pairs);
  n_edges = warp_sum64(n_edges);
  sum_count = warp_sum64(sum_count);
  n_multi = warp_sum64(n_multi);
  if (lane == 0) {
    if (n_multi) atomicAdd(&counters->n_multi, n_multi);
    if (n_pairs) atomicAdd(&counters->n_pairs, n_pairs);
    if (n_edges) atomicAdd(&counters->n_edges, n_edges);
    if (sum_count) atomicAdd(&counters->sum_count, sum_count);
  }
}

// ---------------------------------------------------------------------------------------
// Main pair kernel: one warp per row, rows handed out two at a time IN ORDER, so that related
// proteins (neighbours in the input, which read the same postings) are scored by different
// warps at the same moment and their postings reads hit L2.  The table size of a row comes from
// an optimistic estimate of its distinct partners (rowlogh); new keys are counted exactly and a
// row whose table passes 3/4 full is abandoned, its table cleared, and the row is flagged for
// the safely sized kernels (rowbin = kBinRetry + safe bin).
// ---------------------------------------------------------------------------------------
constexpr int kMainWarps = 4;

// flatten the short postings suffixes (2..15 holders) of one 32-entry chunk into `idx` and start
// the gathers of the first 128 flattened postings; returns the number of flattened postings
// (`idx` must be spelled as <shared array> + offset at the call site: a pointer picked from an array
// of pointers makes ptxas fall back to generic LD/ST, ~10 % of the kernel's instructions)
__device__ __forceinline__ uint32_t chunk_prepare(const uint32_t* __restrict__ col, uint2 e, uint32_t* idx,
                                                  uint32_t (&v)[4]) {
  const uint32_t lane = lane_id();
  const uint32_t len = e.y == kSentinel ? 0u : e.y - e.x;
  const uint32_t slen = len < 16u ? len : 0u;
  const uint32_t incl = warp_scan_incl(slen);
  const uint32_t total = __shfl_sync(kFullMask, incl, 31);
  uint32_t w = incl - slen;
  for (uint32_t j = e.x; j < e.x + slen; ++j) idx[w++] = j;
  __syncwarp();
#pragma unroll
  for (int u = 0; u < 4; ++u) v[u] = lane + 32 * u < total ? col[idx[lane + 32 * u]] : kSentinel;
  return total;
}

__global__ void __launch_bounds__(kMainWarps * 32)
    pairs_main_kernel(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ rowlen,
                      const uint2* __restrict__ suf, const uint32_t* __restrict__ col, uint8_t* __restrict__ rowbin,
                      const uint8_t* __restrict__ rowsafe, const uint8_t* __restrict__ rowlogh, uint32_t n,
                      uint32_t count_bits, uint32_t* __restrict__ row_cursor, uint32_t* __restrict__ n_overflow,
                      uint32_t* __restrict__ bin_counts, EdgeSink sink, PairCounters* __restrict__ counters) {
  constexpr uint32_t HMAX = 1u << kMainLogHMax;
  if (bin_counts[kBinMain] == 0) return;
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  // per warp: table (HMAX words) | idx buffer 0 | idx buffer 1 (doubles as the edge stage) |
  // dirty list (kMainCap u16 slot numbers) | dirty count
  constexpr uint32_t kWarpWords = HMAX + 2 * kIdxPerWarp + kMainCap / 2 + 4;
  uint32_t* wbase = reinterpret_cast<uint32_t*>(dyn_smem) + (size_t)warp * kWarpWords;
  uint32_t* tab = wbase;
  uint32_t* idx0 = wbase + HMAX;  // idx buffer k lives at idx0 + k * kIdxPerWarp
  uint16_t* dirty = reinterpret_cast<uint16_t*>(wbase + HMAX + 2 * kIdxPerWarp);
  uint32_t* dirty_cnt = wbase + HMAX + 2 * kIdxPerWarp + kMainCap / 2;
  if (lane == 0) *dirty_cnt = 0;
  EdgeStage stage{idx0 + kIdxPerWarp, 0u};
  const uint32_t cb = count_bits, cmask = (1u << count_bits) - 1u;
  unsigned long long n_pairs = 0, n_edges = 0, sum_count = 0, n_multi = 0;
  for (uint32_t i = lane * 4; i < HMAX; i += 128)
    *reinterpret_cast<uint4*>(tab + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
  __syncwarp();
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(row_cursor, 4u);
    base = __shfl_sync(kFullMask, base, 0);
    if (base >= n) break;
    // lane l < 4 fetches the metadata of row base + l (one round trip for the four rows)
    uint32_t m_len = 0, m_ps = 0, m_lh = 0;
    bool mine = false;
    if (lane < 4 && base + lane < n && rowbin[base + lane] == kBinMain) {
      mine = true;
      m_len = rowlen[base + lane];
      m_ps = pstart[base + lane];
    }
    uint32_t todo = __ballot_sync(kFullMask, mine);
    while (todo) {
      const uint32_t l = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t r = base + l;
      const uint32_t nl = __shfl_sync(kFullMask, m_len, l), ps = __shfl_sync(kFullMask, m_ps, l);
      constexpr uint32_t log_h = kMainLogHMax, H = HMAX;
      bool full = false, overflow = false;
      auto bump = [&](uint32_t b) { packed_bump_dirty(tab, H - 1u, log_h, cb, b, dirty_cnt, dirty, kMainCap, full); };
      // software pipeline over 32-entry chunks: while chunk c is bumped into the table, the
      // gathers of chunk c+1 and the entry load of chunk c+2 are in flight
      uint2 e_cur = lane < nl ? ld_stream_u32x2(suf + ps + lane) : make_uint2(0, 0);
      uint2 e_nxt = 32 + lane < nl ? ld_stream_u32x2(suf + ps + 32 + lane) : make_uint2(0, 0);
      uint32_t v[4], w[4];
      uint32_t tot_cur = chunk_prepare(col, e_cur, idx0, v), tot_nxt = 0;
      uint32_t k = 0;
      for (uint32_t c = 0; c < nl; c += 32, k ^= 1u) {
        const bool more = c + 32 < nl;
        uint2 e_nn = make_uint2(0, 0);
        if (more) {
          e_nn = c + 64 + lane < nl ? ld_stream_u32x2(suf + ps + c + 64 + lane) : make_uint2(0, 0);
          tot_nxt = chunk_prepare(col, e_nxt, idx0 + (k ^ 1u) * kIdxPerWarp, w);
        }
        // consume chunk c
        if (e_cur.y == kSentinel) bump(e_cur.x);  // inline single partner
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (v[u] != kSentinel) bump(v[u]);
        for (uint32_t t00 = 128; t00 < tot_cur; t00 += 128) {  // rare: more than 128 short postings
          const uint32_t t0 = t00 + lane;
          uint32_t x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = t0 + 32 * u < tot_cur ? col[idx0[k * kIdxPerWarp + t0 + 32 * u]] : kSentinel;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (x[u] != kSentinel) bump(x[u]);
        }
        {  // long suffixes: the whole warp reads 32 consecutive postings at a time
          const uint32_t len = e_cur.y == kSentinel ? 0u : e_cur.y - e_cur.x;
          uint32_t m = __ballot_sync(kFullMask, len >= 16u);
          while (m) {
            const uint32_t src = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t s0 = __shfl_sync(kFullMask, e_cur.x, src), t1 = __shfl_sync(kFullMask, e_cur.y, src);
            for (uint32_t j0 = s0; j0 < t1; j0 += 128) {
              const uint32_t j = j0 + lane;
              uint32_t x[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) x[u] = j + 32 * u < t1 ? col[j + 32 * u] : kSentinel;
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (x[u] != kSentinel) bump(x[u]);
              if (__any_sync(kFullMask, full)) break;
            }
          }
        }
        if (__any_sync(kFullMask, full)) {
          overflow = true;
          break;
        }
        e_cur = e_nxt;
        e_nxt = e_nn;
        tot_cur = tot_nxt;
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = w[u];
      }
      __syncwarp();
      if (overflow) {
        for (uint32_t i = lane * 4; i < H; i += 128)
          *reinterpret_cast<uint4*>(tab + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
        if (lane == 0) {
          rowbin[r] = (uint8_t)(kBinRetry + rowsafe[r]);
          atomicAdd(n_overflow, 1u);
          atomicAdd(&bin_counts[kBinRetry + rowsafe[r]], 1u);
          *dirty_cnt = 0;
        }
        __syncwarp();
        continue;
      }
      // read out and clear exactly the occupied slots
      const uint32_t n_dirty = *dirty_cnt;
      __syncwarp();
      for (uint32_t i0 = 0; i0 < n_dirty; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t sv = kSentinel;
        if (i < n_dirty) {
          const uint32_t h = dirty[i];
          sv = tab[h];
          tab[h] = kSentinel;
        }
        const uint32_t cq = sv != kSentinel ? (sv & cmask) : 0u;
        const bool out = cq > sink.threshold;
        n_pairs += cq != 0;
        n_multi += cq;
        n_edges += out;
        sum_count += out ? cq : 0u;
        if (stage.cnt + 32u > kStageEdges) stage_flush(stage, sink);
        stage_push(stage, out, r, sv >> cb, cq);
      }
      if (lane == 0) *dirty_cnt = 0;
      stage_flush(stage, sink);  // the stage shares idx buffer 1 with the next row's walk
      __syncwarp();
    }
  }
  n_pairs = warp_sum64(n_pairs);
  n_edges = warp_sum64(n_edges);
  sum_count = warp_sum64(sum_count);
  n_multi = warp_sum64(n_multi);
  if (lane == 0) {
    if (n_multi) atomicAdd(&counters->n_multi, n_multi);
    if (n_pairs) atomicAdd(&counters->n_pairs, n_pairs);
    if (n_edges) atomicAdd(&counters->n_edges, n_edges);
    if (sum_count) atomicAdd(&counters->sum_count, sum_count);
  }
}

// ---------------------------------------------------------------------------------------
// Scored main kernel (want_blosum): same walk as pairs_main_kernel, but every multi-edge also
// carries the BLOSUM62 self-score of its k-mer (sufss, one byte per row entry), accumulated in
// the same atomic as the count.  K9 fused into K7: no per-edge list intersection afterwards.
// ---------------------------------------------------------------------------------------
// two instances: <kMainLogHMax, kMainCap, kBinMain, 5 warps> and the half-size <kMainSLogH, kMainSCap, kBinMainS, 7 warps>
// (9.9 KB instead of 14.6 KB of shared memory per warp: 21 instead of 15 rows in flight per SM)
template <uint32_t LOG_H, uint32_t CAP>
__host__ __device__ constexpr uint32_t scored_warp_words() {
  return 2 * (1u << LOG_H) + 2 * kIdxPerWarp + 2 * (kIdxPerWarp / 4) + CAP / 2 + 4;
}

__device__ __forceinline__ uint32_t chunk_prepare_scored(const uint32_t* __restrict__ col, uint2 e, uint32_t ss,
                                                         uint32_t* idx, uint8_t* idxs, uint32_t (&v)[4],
                                                         uint32_t (&sv)[4]) {
  const uint32_t lane = lane_id();
  const uint32_t len = e.y == kSentinel ? 0u : e.y - e.x;
  const uint32_t slen = len < 16u ? len : 0u;
  const uint32_t incl = warp_scan_incl(slen);
  const uint32_t total = __shfl_sync(kFullMask, incl, 31);
  uint32_t w = incl - slen;
  for (uint32_t j = e.x; j < e.x + slen; ++j) {
    idx[w] = j;
    idxs[w++] = (uint8_t)ss;
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint32_t t = lane + 32 * u;
    v[u] = t < total ? col[idx[t]] : kSentinel;
    sv[u] = t < total ? idxs[t] : 0u;
  }
  return total;
}

template <uint32_t LOG_H, uint32_t CAP, uint8_t BIN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
    pairs_main_scored_kernel(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ rowlen,
                             const uint2* __restrict__ suf, const uint8_t* __restrict__ sufss,
                             const uint32_t* __restrict__ col, uint8_t* __restrict__ rowbin,
                             const uint8_t* __restrict__ rowsafe, uint32_t n, uint32_t* __restrict__ row_cursor,
                             uint32_t* __restrict__ n_overflow, uint32_t* __restrict__ bin_counts, EdgeSink sink,
                             PairCounters* __restrict__ counters) {
  constexpr uint32_t HMAX = 1u << LOG_H;
  constexpr uint32_t log_h = LOG_H;
  constexpr uint32_t kWords = scored_warp_words<LOG_H, CAP>();
  if (bin_counts[BIN] == 0) return;
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  // per warp: keys | values | idx buffers 0,1 (buffer 1 doubles as the edge stage) | idx score
  // buffers 0,1 | dirty list | dirty count
  uint32_t* wbase = reinterpret_cast<uint32_t*>(dyn_smem) + (size_t)warp * kWords;
  uint32_t* key = wbase;
  uint32_t* val = wbase + HMAX;
  uint32_t* idx0 = wbase + 2 * HMAX;
  uint8_t* idxs0 = reinterpret_cast<uint8_t*>(wbase + 2 * HMAX + 2 * kIdxPerWarp);
  uint16_t* dirty = reinterpret_cast<uint16_t*>(wbase + 2 * HMAX + 2 * kIdxPerWarp + 2 * (kIdxPerWarp / 4));
  uint32_t* dirty_cnt = wbase + 2 * HMAX + 2 * kIdxPerWarp + 2 * (kIdxPerWarp / 4) + CAP / 2;
  if (lane == 0) *dirty_cnt = 0;
  EdgeStage stage{idx0 + kIdxPerWarp, 0u};
  unsigned long long n_pairs = 0, n_edges = 0, sum_count = 0, n_multi = 0;
  for (uint32_t i = lane * 4; i < HMAX; i += 128) {
    *reinterpret_cast<uint4*>(key + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
    *reinterpret_cast<uint4*>(val + i) = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(row_cursor, 4u);
    base = __shfl_sync(kFullMask, base, 0);
    if (base >= n) break;
    uint32_t m_len = 0, m_ps = 0;
    bool mine = false;
    if (lane < 4 && base + lane < n && rowbin[base + lane] == BIN) {
      mine = true;
      m_len = rowlen[base + lane];
      m_ps = pstart[base + lane];
    }
    uint32_t todo = __ballot_sync(kFullMask, mine);
    while (todo) {
      const uint32_t l = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t r = base + l;
      const uint32_t nl = __shfl_sync(kFullMask, m_len, l), ps = __shfl_sync(kFullMask, m_ps, l);
      bool full = false, overflow = false;
      uint32_t n_new = 0;  // distinct partners so far = length of the dirty list (warp-uniform)
      // all 32 lanes call it: bump the lane's partner (if any) and append the new slots to the dirty list
      auto bump = [&](bool valid, uint32_t b, uint32_t ss) {
        const uint32_t slot =
            valid && !full ? scored_bump_new(key, val, HMAX - 1u, log_h, b, (ss << kScoreShift) | 1u) : kSentinel;
        const uint32_t m = __ballot_sync(kFullMask, slot != kSentinel);
        if (m) {
          if (slot != kSentinel) {
            const uint32_t pos = n_new + __popc(m & lanemask_lt());
            if (pos < CAP) dirty[pos] = (uint16_t)slot;
          }
          n_new += __popc(m);
          full = n_new > CAP;  // at most 32 slots beyond the cap: the table (1024 slots) never fills up
        }
      };
      uint2 e_cur = lane < nl ? ld_stream_u32x2(suf + ps + lane) : make_uint2(0, 0);
      uint32_t s_cur = lane < nl ? sufss[ps + lane] : 0u;
      uint2 e_nxt = 32 + lane < nl ? ld_stream_u32x2(suf + ps + 32 + lane) : make_uint2(0, 0);
      uint32_t s_nxt = 32 + lane < nl ? sufss[ps + 32 + lane] : 0u;
      uint32_t v[4], sv[4], w[4], sw[4];
      uint32_t tot_cur = chunk_prepare_scored(col, e_cur, s_cur, idx0, idxs0, v, sv), tot_nxt = 0;
      uint32_t k = 0;
      for (uint32_t c = 0; c < nl; c += 32, k ^= 1u) {
        const bool more = c + 32 < nl;
        uint2 e_nn = make_uint2(0, 0);
        uint32_t s_nn = 0;
        if (more) {
          if (c + 64 + lane < nl) {
            e_nn = ld_stream_u32x2(suf + ps + c + 64 + lane);
            s_nn = sufss[ps + c + 64 + lane];
          }
          tot_nxt = chunk_prepare_scored(col, e_nxt, s_nxt, idx0 + (k ^ 1u) * kIdxPerWarp,
                                         idxs0 + (k ^ 1u) * kIdxPerWarp, w, sw);
        }
        bump(e_cur.y == kSentinel, e_cur.x, s_cur);  // inline single partner
#pragma unroll
        for (int u = 0; u < 4; ++u) bump(v[u] != kSentinel, v[u], sv[u]);
        for (uint32_t t00 = 128; t00 < tot_cur; t00 += 128) {  // rare: more than 128 short postings
          const uint32_t t0 = t00 + lane;
          uint32_t x[4], sx[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t t = t0 + 32 * u;
            x[u] = t < tot_cur ? col[idx0[k * kIdxPerWarp + t]] : kSentinel;
            sx[u] = t < tot_cur ? idxs0[k * kIdxPerWarp + t] : 0u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) bump(x[u] != kSentinel, x[u], sx[u]);
        }
        {  // long suffixes: the whole warp reads 32 consecutive postings at a time
          const uint32_t len = e_cur.y == kSentinel ? 0u : e_cur.y - e_cur.x;
          uint32_t m = __ballot_sync(kFullMask, len >= 16u);
          while (m) {
            const uint32_t src = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t s0 = __shfl_sync(kFullMask, e_cur.x, src), t1 = __shfl_sync(kFullMask, e_cur.y, src);
            const uint32_t ssl = __shfl_sync(kFullMask, s_cur, src);
            for (uint32_t j0 = s0; j0 < t1; j0 += 128) {
              const uint32_t j = j0 + lane;
              uint32_t x[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) x[u] = j + 32 * u < t1 ? col[j + 32 * u] : kSentinel;
#pragma unroll
              for (int u = 0; u < 4; ++u) bump(x[u] != kSentinel, x[u], ssl);
              if (full) break;
            }
          }
        }
        if (full) {
          overflow = true;
          break;
        }
        e_cur = e_nxt;
        s_cur = s_nxt;
        e_nxt = e_nn;
        s_nxt = s_nn;
        tot_cur = tot_nxt;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u] = w[u];
          sv[u] = sw[u];
        }
      }
      __syncwarp();
      if (overflow) {
        for (uint32_t i = lane * 4; i < HMAX; i += 128) {
          *reinterpret_cast<uint4*>(key + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
          *reinterpret_cast<uint4*>(val + i) = make_uint4(0, 0, 0, 0);
        }
        if (lane == 0) {
          rowbin[r] = (uint8_t)(kBinRetry + rowsafe[r]);
          atomicAdd(n_overflow, 1u);
          atomicAdd(&bin_counts[kBinRetry + rowsafe[r]], 1u);
          *dirty_cnt = 0;
        }
        __syncwarp();
        continue;
      }
      const uint32_t n_dirty = n_new;
      __syncwarp();
      for (uint32_t i0 = 0; i0 < n_dirty; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t b = kSentinel, vv = 0;
        if (i < n_dirty) {
          const uint32_t h = dirty[i];
          b = key[h];
          vv = val[h];
          key[h] = kSentinel;
          val[h] = 0;
        }
        const uint32_t cq = vv & ((1u << kScoreShift) - 1u);
        const bool out = cq > sink.threshold;
        n_pairs += cq != 0;
        n_multi += cq;
        n_edges += out;
        sum_count += out ? cq : 0u;
        if (stage.cnt + 32u > kStageEdges4) stage4_flush(stage, sink);
        stage4_push(stage, out, r, b, cq, vv >> kScoreShift);
      }
      if (lane == 0) *dirty_cnt = 0;
      stage4_flush(stage, sink);  // the stage shares idx buffer 1 with the next row's walk
      __syncwarp();
    }
  }
  n_pairs = warp_sum64(n_pairs);
  n_edges = warp_sum64(n_edges);
  sum_count = warp_sum64(sum_count);
  n_multi = warp_sum64(n_multi);
  if (lane == 0) {
    if (n_multi) atomicAdd(&counters->n_multi, n_multi);
    if (n_pairs) atomicAdd(&counters->n_pairs, n_pairs);
    if (n_edges) atomicAdd(&counters->n_edges, n_edges);
    if (sum_count) atomicAdd(&counters->sum_count, sum_count);
  }
}

// ---------------------------------------------------------------------------------------
// Stream kernel (the fast path when the materialised multi-edge lists fit in HBM): row r's
// partners are plist[rowbase[r] .. +rowwork[r]), one u32 per multi-edge, written by the index
// stage.  A warp streams them with coalesced 128-byte loads (next 128 in flight while the
// current 128 are bumped) into its shared-memory table: exactly the 4 algorithmic bytes per
// multi-edge, no postings gather, no flattening.  Same table / dirty-list / overflow protocol
// as pairs_main_kernel; SCORED adds the BLOSUM self-score stream (one byte per multi-edge).
// ---------------------------------------------------------------------------------------
constexpr int kStreamWarps = 4;
constexpr uint32_t kStreamStageWords = 256;  // 64 scored or 85 unscored edges per flush
template <bool SCORED>
__host__ __device__ constexpr uint32_t stream_warp_words() {
  return (SCORED ? 2u : 1u) * (1u << kMainLogHMax) + kMainCap / 2 + 4 + kStreamStageWords;
}

template <bool SCORED>
__global__ void __launch_bounds__(kStreamWarps * 32)
    pairs_stream_kernel(const unsigned long long* __restrict__ rowbase, const uint32_t* __restrict__ rowwork,
                        const uint32_t* __restrict__ plist, const uint8_t* __restrict__ pss,
                        uint8_t* __restrict__ rowbin, const uint8_t* __restrict__ rowsafe, uint32_t n,
                        uint32_t count_bits, uint32_t* __restrict__ row_cursor, uint32_t* __restrict__ n_overflow,
                        uint32_t* __restrict__ bin_counts, EdgeSink sink, PairCounters* __restrict__ counters) {
  constexpr uint32_t HMAX = 1u << kMainLogHMax;
  constexpr uint32_t log_h = kMainLogHMax;
  if (bin_counts[kBinMain] == 0) return;
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t* wbase = reinterpret_cast<uint32_t*>(dyn_smem) + (size_t)warp * stream_warp_words<SCORED>();
  uint32_t* key = wbase;                              // SCORED: keys; else packed key|count slots
  uint32_t* val = wbase + HMAX;                       // SCORED only: score << 12 | count
  uint32_t* after = wbase + (SCORED ? 2u : 1u) * HMAX;
  uint16_t* dirty = reinterpret_cast<uint16_t*>(after);
  uint32_t* dirty_cnt = after + kMainCap / 2;
  EdgeStage stage{after + kMainCap / 2 + 4, 0u};
  const uint32_t cb = count_bits, cmask = (1u << count_bits) - 1u;
  unsigned long long n_pairs = 0, n_edges = 0, sum_count = 0, n_multi = 0;
  if (lane == 0) *dirty_cnt = 0;
  for (uint32_t i = lane * 4; i < HMAX; i += 128) {
    *reinterpret_cast<uint4*>(key + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
    if (SCORED) *reinterpret_cast<uint4*>(val + i) = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(row_cursor, 4u);
    base = __shfl_sync(kFullMask, base, 0);
    if (base >= n) break;
    uint32_t m_work = 0, m_lo = 0, m_hi = 0;
    bool mine = false;
    if (lane < 4 && base + lane < n && rowbin[base + lane] == kBinMain) {
      mine = true;
      m_work = rowwork[base + lane];
      const unsigned long long rb = rowbase[base + lane];
      m_lo = (uint32_t)rb;
      m_hi = (uint32_t)(rb >> 32);
    }
    uint32_t todo = __ballot_sync(kFullMask, mine);
    while (todo) {
      const uint32_t l = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t r = base + l;
      const uint32_t P = __shfl_sync(kFullMask, m_work, l);
      const unsigned long long rb =
          ((unsigned long long)__shfl_sync(kFullMask, m_hi, l) << 32) | __shfl_sync(kFullMask, m_lo, l);
      const uint32_t* pl = plist + rb;
      const uint8_t* sl = SCORED ? pss + rb : nullptr;
      bool full = false;
      uint32_t v[4], sv[4], w[4], sw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t t = lane + 32 * u;
        v[u] = t < P ? ld_stream_u32(pl + t) : kSentinel;
        sv[u] = SCORED && t < P ? sl[t] : 0u;
      }
      for (uint32_t t0 = 0; t0 < P; t0 += 128) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // next 128 multi-edges in flight while these are bumped
          const uint32_t t = t0 + 128 + lane + 32 * u;
          w[u] = t < P ? ld_stream_u32(pl + t) : kSentinel;
          sw[u] = SCORED && t < P ? sl[t] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (v[u] == kSentinel) continue;
          if (SCORED)
            scored_bump_dirty(key, val, HMAX - 1u, log_h, v[u], (sv[u] << kScoreShift) | 1u, dirty_cnt, dirty,
                              kMainCap, full);
          else
            packed_bump_dirty(key, HMAX - 1u, log_h, cb, v[u], dirty_cnt, dirty, kMainCap, full);
        }
        if (__any_sync(kFullMask, full)) break;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u] = w[u];
          sv[u] = sw[u];
        }
      }
      __syncwarp();
      if (__any_sync(kFullMask, full)) {  // more distinct partners than the table takes: safe kernels
        for (uint32_t i = lane * 4; i < HMAX; i += 128) {
          *reinterpret_cast<uint4*>(key + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
          if (SCORED) *reinterpret_cast<uint4*>(val + i) = make_uint4(0, 0, 0, 0);
        }
        if (lane == 0) {
          rowbin[r] = (uint8_t)(kBinRetry + rowsafe[r]);
          atomicAdd(n_overflow, 1u);
          atomicAdd(&bin_counts[kBinRetry + rowsafe[r]], 1u);
          *dirty_cnt = 0;
        }
        __syncwarp();
        continue;
      }
      const uint32_t n_dirty = *dirty_cnt;
      __syncwarp();
      for (uint32_t i0 = 0; i0 < n_dirty; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t b = kSentinel, cq = 0, score = 0;
        if (i < n_dirty) {
          const uint32_t h = dirty[i];
          if (SCORED) {
            b = key[h];
            const uint32_t vv = val[h];
            key[h] = kSentinel;
            val[h] = 0;
            cq = vv & ((1u << kScoreShift) - 1u);
            score = vv >> kScoreShift;
          } else {
            const uint32_t sv0 = key[h];
            key[h] = kSentinel;
            b = sv0 >> cb;
            cq = sv0 & cmask;
          }
        }
        const bool out = cq > sink.threshold;
        n_pairs += cq != 0;
        n_multi += cq;
        n_edges += out;
        sum_count += out ? cq : 0u;
        if (SCORED) {
          if (stage.cnt + 32u > kStreamStageWords / 4) stage4_flush(stage, sink);
          stage4_push(stage, out, r, b, cq, score);
        } else {
          if (stage.cnt + 32u > kStreamStageWords / 3) stage_flush(stage, sink);
          stage_push(stage, out, r, b, cq);
        }
      }
      if (lane == 0) *dirty_cnt = 0;
      __syncwarp();
    }
  }
  if (SCORED) stage4_flush(stage, sink); else stage_flush(stage, sink);
  n_pairs = warp_sum64(n_pairs);
  n_edges = warp_sum64(n_edges);
  sum_count = warp_sum64(sum_count);
  n_multi = warp_sum64(n_multi);
  if (lane == 0) {
    if (n_multi) atomicAdd(&counters->n_multi, n_multi);
    if (n_pairs) atomicAdd(&counters->n_pairs, n_pairs);
    if (n_edges) atomicAdd(&counters->n_edges, n_edges);
    if (sum_count) atomicAdd(&counters->sum_count, sum_count);
  }
}

// ---------------------------------------------------------------------------------------
// Bin-local pairs (partitioned index, bucket.cuh): both rows in the same 64-row bin.  The index
// stage leaves, per bin, the "run records" of its k-mers: the holders of a k-mer inside the bin
// as a 64-bit mask plus the k-mer's BLOSUM self-score.  One CTA per bin accumulates every pair
// of set bits on a dense 64 x 64 tile of shared-memory counters (count and score), then reads the
// tile out: no hashing, no postings.  The hash kernels never see these partners (the suffix of an
// entry starts behind the row's bin).
// ---------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;
constexpr size_t kTileSmemBytes = (size_t)(2 * kBinRows * kBinRows + (kTileThreads / 32) * kStageWords) * 4;

template <bool SCORED, bool CROSS>
__global__ void __launch_bounds__(kTileThreads)
    pairs_tile_kernel(const uint4* __restrict__ runs, const uint32_t* __restrict__ rowcap_prefix,
                      const uint32_t* __restrict__ run_cnt, uint32_t n, uint32_t n_bins,
                      const uint32_t* __restrict__ first_after, const uint32_t* __restrict__ bounds, RowOwner owner,
                      EdgeSink sink, PairCounters* __restrict__ counters) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(dyn_smem);  // [64][64] shared k-mers of (i, j), i < j
  uint32_t* s_score = s_cnt + kBinRows * kBinRows;          // [64][64] sum of their self-scores
  __shared__ uint32_t s_fa[kBinRows];
  const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  EdgeStage stage{s_score + kBinRows * kBinRows + warp * kStageWords, 0u};
  const uint32_t row_lo = bounds ? bounds[0] : 0u, row_hi = bounds ? bounds[1] : n;
  unsigned long long n_pairs = 0, n_edges = 0, sum_count = 0, n_multi = 0;
  for (uint32_t i = tid; i < kBinRows * kBinRows; i += kTileThreads) {
    s_cnt[i] = 0;
    s_score[i] = 0;
  }
  __syncthreads();
  for (uint32_t bin = blockIdx.x; bin < n_bins; bin += gridDim.x) {
    const uint32_t r0 = bin << kBinRowsLog;
    const uint32_t nr = run_cnt[bin];
    if (nr == 0 || !owner.mine(r0) || r0 >= row_hi || r0 + kBinRows <= row_lo) continue;  // uniform per CTA
    if (CROSS && tid < kBinRows) s_fa[tid] = r0 + tid < n ? first_after[r0 + tid] : n;
    if (CROSS) __syncthreads();
    const uint4* src = runs + (rowcap_prefix[r0] >> 1);
    for (uint32_t t = tid; t < nr; t += kTileThreads) {
      const uint4 rec = ld_stream_u32x4(src + t);
      const uint32_t ss = rec.x & 0xFFu;
      unsigned long long m = ((unsigned long long)rec.w << 32) | rec.z;
      while (m) {
        const uint32_t i = (uint32_t)__ffsll((long long)m) - 1u;
        m &= m - 1ull;
        if (r0 + i < row_lo || r0 + i >= row_hi) continue;  // row i is scored by another shard
        unsigned long long rest = m;  // the holders after row i
        if (CROSS) {  // only the rows of later class blocks pair with row i
          const uint32_t fa = s_fa[i];
          rest = fa >= r0 + kBinRows ? 0ull : rest & ~((1ull << (fa - r0)) - 1ull);
        }
        while (rest) {
          const uint32_t j = (uint32_t)__ffsll((long long)rest) - 1u;
          rest &= rest - 1ull;
          atomicAdd(&s_cnt[i * kBinRows + j], 1u);
          if (SCORED) atomicAdd(&s_score[i * kBinRows + j], ss);
        }
      }
    }
    __syncthreads();
    for (uint32_t c0 = 0; c0 < kBinRows * kBinRows; c0 += kTileThreads) {  // read out and clear
      const uint32_t cell = c0 + tid;
      const uint32_t cq = s_cnt[cell];
      if (__any_sync(kFullMask, cq != 0)) {
        const uint32_t score = SCORED ? s_score[cell] : 0u;
        if (cq) {
          s_cnt[cell] = 0;
          if (SCORED) s_score[cell] = 0;
        }
        const bool out = cq > sink.threshold;
        n_pairs += cq != 0;
        n_multi += cq;
        n_edges += out;
        sum_count += out ? cq : 0u;
        if (SCORED) {
          if (stage.cnt + 32u > kStageEdges4) stage4_flush(stage, sink);
          stage4_push(stage, out, r0 + cell / kBinRows, r0 + cell % kBinRows, cq, score);
        } else {
          if (stage.cnt + 32u > kStageEdges) stage_flush(stage, sink);
          stage_push(stage, out, r0 + cell / kBinRows, r0 + cell % kBinRows, cq);
        }
      }
    }
    __syncthreads();
  }
  if (SCORED) stage4_flush(stage, sink); else stage_flush(stage, sink);
  n_pairs = warp_sum64(n_pairs);
  n_edges = warp_sum64(n_edges);
  sum_count = warp_sum64(sum_count);
  n_multi = warp_sum64(n_multi);
  if (lane == 0) {
    if (n_multi) atomicAdd(&counters->n_multi, n_multi);
    if (n_pairs) atomicAdd(&counters->n_pairs, n_pairs);
    if (n_edges) atomicAdd(&counters->n_edges, n_edges);
    if (sum_count) atomicAdd(&counters->sum_count, sum_count);
  }
}

// ---------------------------------------------------------------------------------------
// dense accumulators: one CTA per row, one counter per candidate partner, in column blocks of
// `block_cols` partners.  WIDE = u32 counters (rows with >= 65535 ids), else two u16 per word.
// The rows that land here hold thousands to hundreds of thousands of partners (the k = 5 sets; long
// proteins of large k = 7 sets), far fewer than the columns they span, so what costs is the sweep over the
// counters and the clipping of the postings suffixes to the block, not the increments:
//   * 1 024 threads per CTA (one CTA per SM at 200 KB of counters: 32 warps instead of 8 to issue the sweep);
//   * ONE sweep per block, 16-byte loads: read, count, emit, and store zeros back only where something was
//     counted (the counters are zeroed once per CTA; the multi-edge total comes from the walk, not the sweep);
//     a warp whose 128 words hold no counter over the threshold never enters the emission code;
//   * blocks ascend, so every thread carries, for the entries it owns, where the previous block's clip
//     ended (registers): one galloping search per entry and block instead of two binary searches.
// ---------------------------------------------------------------------------------------
constexpr int kDenseThreads = 1024;
constexpr int kDenseCarry = 4;  // entries per thread with a carried cursor (rows of <= 4 096 entries; beyond: searched)

// first posting >= x in col[lo, hi) (ascending), galloping from lo
__device__ __forceinline__ uint32_t dense_clip_end(const uint32_t* __restrict__ col, uint32_t lo, uint32_t hi, uint32_t x) {
  uint32_t step = 8;
  while (lo + step < hi && col[lo + step - 1u] < x) {
    lo += step;
    step <<= 1;
  }
  hi = min(hi, lo + step);
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (col[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <bool WIDE>
__global__ void __launch_bounds__(kDenseThreads)
    pairs_dense_kernel(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ rowlen,
                       const uint2* __restrict__ suf, const uint32_t* __restrict__ col,
                       const uint32_t* __restrict__ first_after, const uint8_t* __restrict__ rowbin, uint32_t n,
                       uint32_t block_cols, uint32_t* __restrict__ row_cursor, const uint32_t* __restrict__ bin_counts,
                       EdgeSink sink, PairCounters* __restrict__ counters) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  if (bin_counts[kBinDense] == 0) return;
  __shared__ uint32_t s_base;
  uint32_t* acc = reinterpret_cast<uint32_t*>(dyn_smem);
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned long long n_pairs = 0, n_edges = 0, sum_count = 0, n_multi = 0;
  {  // the counters start (and every sweep leaves them) all zero
    const uint32_t cap_words4 = ((WIDE ? block_cols : (block_cols + 1) / 2) + 3u) & ~3u;
    for (uint32_t i = threadIdx.x * 4; i < cap_words4; i += kDenseThreads * 4)
      *reinterpret_cast<uint4*>(acc + i) = make_uint4(0, 0, 0, 0);
  }
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_base = atomicAdd(row_cursor, 32u);
    __syncthreads();
    const uint32_t base = s_base;
    if (base >= n) break;
    uint32_t todo = __ballot_sync(kFullMask, base + lane < n && rowbin[base + lane] == kBinDense);
    while (todo) {
      const uint32_t r = base + __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t nl = rowlen[r], ps = pstart[r];
      const uint32_t first = first_after ? first_after[r] : r + 1;
      const bool multipass = n - first > block_cols;
      uint32_t cur[kDenseCarry];  // entry threadIdx.x + k * kDenseThreads: where the last block's clip ended
#pragma unroll
      for (int k = 0; k < kDenseCarry; ++k) cur[k] = kSentinel;
      uint32_t bumps = 0;
      for (uint32_t blk_lo = first; blk_lo < n; blk_lo += block_cols) {
        const uint32_t blk_hi = min(n, blk_lo + block_cols);
        const uint32_t ncols = blk_hi - blk_lo;
        const uint32_t words = WIDE ? ncols : (ncols + 1) / 2;  // only what this block needs
        const uint32_t words4 = (words + 3u) & ~3u;
        auto bump = [&](uint32_t b) {
          const uint32_t idx = b - blk_lo;
          ++bumps;
          if (WIDE) atomicAdd(&acc[idx], 1u);
          else atomicAdd(&acc[idx >> 1], 1u << ((idx & 1u) * 16u));
        };
#pragma unroll
        for (int k = 0; k < kDenseCarry; ++k) {
          const uint32_t c = warp * 32u + (uint32_t)k * kDenseThreads;
          if (c >= nl) break;  // warp-uniform
          uint2 e = c + lane < nl ? ld_stream_u32x2(suf + ps + c + lane) : make_uint2(0, 0);
          if (multipass) {
            if (e.y == kSentinel) {
              if (e.x < blk_lo || e.x >= blk_hi) e = make_uint2(0, 0);
            } else if (e.y > e.x) {  // the holders in [blk_lo, blk_hi): from where the previous block stopped
              const uint32_t lo = cur[k] != kSentinel ? cur[k] : e.x;
              const uint32_t hi = blk_hi >= n ? e.y : dense_clip_end(col, lo, e.y, blk_hi);
              cur[k] = hi;
              e = make_uint2(lo, hi);
            }
          }
          walk_chunk(col, e, bump);
        }
        for (uint32_t c = warp * 32u + (uint32_t)kDenseCarry * kDenseThreads; c < nl; c += kDenseThreads) {
          uint2 e = c + lane < nl ? ld_stream_u32x2(suf + ps + c + lane) : make_uint2(0, 0);
          if (multipass) {  // (rows of more than 4 096 entries: both ends searched)
            if (e.y == kSentinel) {
              if (e.x < blk_lo || e.x >= blk_hi) e = make_uint2(0, 0);
            } else if (e.y > e.x) {
              uint32_t lo = e.x, hi = e.y;
              while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (col[mid] < blk_lo) lo = mid + 1; else hi = mid;
              }
              e = make_uint2(lo, dense_clip_end(col, lo, e.y, blk_hi));
            }
          }
          walk_chunk(col, e, bump);
        }
        __syncthreads();
        // the sweep: read, count, emit, clear (warp-uniform trip count)
        for (uint32_t i0 = warp * 128u; i0 < words4; i0 += kDenseThreads * 4u) {
          const uint32_t i = i0 + lane * 4u;
          uint4 x = make_uint4(0, 0, 0, 0);
          if (i < words4) x = *reinterpret_cast<const uint4*>(acc + i);
          const bool any = (x.x | x.y | x.z | x.w) != 0u;
          if (!__any_sync(kFullMask, any)) continue;
          if (any) *reinterpret_cast<uint4*>(acc + i) = make_uint4(0, 0, 0, 0);
          const uint32_t w[4] = {x.x, x.y, x.z, x.w};
          const uint32_t thr = sink.threshold;
          bool hot = false;
          uint32_t nz = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (WIDE) {
              nz += w[q] != 0u;
              hot |= w[q] > thr;
            } else {
              const uint32_t c0 = w[q] & 0xFFFFu, c1 = w[q] >> 16;
              nz += (c0 != 0u) + (c1 != 0u);
              hot |= c0 > thr || c1 > thr;
            }
          }
          n_pairs += nz;
          if (!__any_sync(kFullMask, hot)) continue;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (WIDE) {
              const bool out = w[q] > thr;
              n_edges += out;
              sum_count += out ? w[q] : 0u;
              emit_edge(out, r, blk_lo + i + q, w[q], sink);
            } else {
              const uint32_t c0 = w[q] & 0xFFFFu, c1 = w[q] >> 16;
              const bool o0 = c0 > thr, o1 = c1 > thr;
              n_edges += (uint32_t)o0 + (uint32_t)o1;
              sum_count += (o0 ? c0 : 0u) + (o1 ? c1 : 0u);
              emit_edge(o0, r, blk_lo + 2u * (i + q), c0, sink);
              emit_edge(o1, r, blk_lo + 2u * (i + q) + 1u, c1, sink);
            }
          }
        }
        __syncthreads();
      }
      n_multi += bumps;
    }
  }
  n_pairs = warp_sum64(n_pairs);
  n_edges = warp_sum64(n_edges);
  sum_count = warp_sum64(sum_count);
  n_multi = warp_sum64(n_multi);
  if (lane == 0) {
    if (n_multi) atomicAdd(&counters->n_multi, n_multi);
    if (n_pairs) atomicAdd(&counters->n_pairs, n_pairs);
    if (n_edges) atomicAdd(&counters->n_edges, n_edges);
    if (sum_count) atomicAdd(&counters->sum_count, sum_count);
  }
}

// ---------------------------------------------------------------------------------------
// K9: per-edge sorted-list intersection of the two id rows.  One warp per edge: lanes take the
// shorter row's ids and binary-search the longer row.  MODE 0: sum of BLOSUM62 self-scores of
// the shared k-mers -> edge.w.  (The list itself is produced by shared_kmers_kernel.)
// ---------------------------------------------------------------------------------------
// per-warp id hash of SLOTS entries: rows of up to 0.7 * SLOTS ids (longer rows: binary search)
// Runs over the SORTED edge list (keys = a << 32 | b in input order).  A warp takes a window of
// 32 consecutive edges; while `a` stays the same it keeps row a's ids in a shared-memory hash
// set (id -> BLOSUM62 self-score), streams row b's ids with coalesced loads and probes.
// vals = count | blosum << 32 (blosum filled in here).
// <1024, 22, 4>: rows of up to 716 ids; <4096, 20, 2>: longer rows, load factor <= 0.5 (rows of up to 2 048 ids are hashed
// once per run of edges with the same first row; at 2 048 slots and load 0.7 the unsuccessful probes of row b, most of
// them, took ~6 steps each and rows beyond 1 433 ids were re-hashed in chunks for EVERY edge: 12 000 warp instructions
// per edge at 5.5 active lanes on the k = 5 sets, profiles/r2_history.md)
template <uint32_t SLOTS, uint32_t SHIFT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
    edge_blosum_kernel(const unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals,
                       unsigned long long n_edges, const uint32_t* __restrict__ rank_of,
                       const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ rowlen,
                       const uint32_t* __restrict__ ids, const uint8_t* __restrict__ selfscore) {
  __shared__ uint32_t s_key[WARPS][SLOTS];
  __shared__ uint8_t s_val[WARPS][SLOTS];
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
  uint32_t* hk = s_key[w];
  uint8_t* hv = s_val[w];
  const unsigned long long n_win = (n_edges + 31) / 32;
  const unsigned long long gw = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long win = gw; win < n_win; win += nw) {
    // lane l fetches the metadata of edge win*32+l: all dependent loads of the window overlap
    const unsigned long long my_e = win * 32 + lane;
    uint32_t m_a = kSentinel, m_pa = 0, m_na = 0, m_pb = 0, m_nb = 0, m_score = 0;
    unsigned long long m_val = 0;
    if (my_e < n_edges) {
      const unsigned long long key = keys[my_e];
      m_val = vals[my_e];
      m_a = (uint32_t)(key >> 32);
      const uint32_t b = (uint32_t)key;
      const uint32_t ra = rank_of ? rank_of[m_a] : m_a, rb = rank_of ? rank_of[b] : b;
      m_pa = pstart[ra];
      m_na = rowlen[ra];
      m_pb = pstart[rb];
      m_nb = rowlen[rb];
    }
    const uint32_t n_in_win = (uint32_t)min(32ull, n_edges - win * 32);
    uint32_t cur_a = kSentinel;
    bool hashed = false;
    const uint32_t* A = nullptr;
    uint32_t na = 0;
    // software pipeline: the first 256 ids of the next edge's row b are loaded while the
    // current edge is probed
    uint32_t nxt[8];
    {
      const uint32_t* B0 = ids + __shfl_sync(kFullMask, m_pb, 0);
      const uint32_t nb0 = __shfl_sync(kFullMask, m_nb, 0);
#pragma unroll
      for (int u = 0; u < 8; ++u) nxt[u] = lane + 32 * u < nb0 ? B0[lane + 32 * u] : kSentinel;
    }
    // edges whose score came out of the scored pair kernel keep it; only the rest is intersected
    const uint32_t need = __ballot_sync(kFullMask, my_e < n_edges && (m_val >> 63) != 0);
    if (!need) continue;
    m_score = (uint32_t)(m_val >> 32);
    for (uint32_t l = 0; l < n_in_win; ++l) {
      uint32_t xs[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) xs[u] = nxt[u];
      if (l + 1 < n_in_win) {
        const uint32_t* B1 = ids + __shfl_sync(kFullMask, m_pb, l + 1);
        const uint32_t nb1 = __shfl_sync(kFullMask, m_nb, l + 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) nxt[u] = lane + 32 * u < nb1 ? B1[lane + 32 * u] : kSentinel;
      }
      const uint32_t a = __shfl_sync(kFullMask, m_a, l);
      const uint32_t pa = __shfl_sync(kFullMask, m_pa, l), na_l = __shfl_sync(kFullMask, m_na, l);
      if (!((need >> l) & 1u)) continue;
      // hash (a chunk of) row a's ids into the warp's table: id -> self-score
      auto build = [&](uint32_t c0, uint32_t c1) {
        __syncwarp();
        for (uint32_t i = lane * 4; i < SLOTS; i += 128)
          *reinterpret_cast<uint4*>(hk + i) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);
        __syncwarp();
        for (uint32_t i0 = c0 + lane; i0 < c1; i0 += 256) {
          uint32_t as[8];
          uint8_t sc[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) as[u] = i0 + 32 * u < c1 ? A[i0 + 32 * u] : kSentinel;
#pragma unroll
          for (int u = 0; u < 8; ++u) sc[u] = as[u] != kSentinel ? selfscore[as[u]] : (uint8_t)0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t x = as[u];
            if (x == kSentinel) continue;
            uint32_t h = (x * 2654435761u) >> SHIFT;
            while (atomicCAS(&hk[h], kSentinel, x) != kSentinel) h = (h + 1u) & (SLOTS - 1u);
            hv[h] = sc[u];
          }
        }
        __syncwarp();
      };
      auto probe = [&](uint32_t x) -> int {
        if (x == kSentinel) return 0;
        uint32_t h = (x * 2654435761u) >> SHIFT;
        for (;;) {
          const uint32_t k = hk[h];
          if (k == x) return hv[h];
          if (k == kSentinel) return 0;
          h = (h + 1u) & (SLOTS - 1u);
        }
      };
      constexpr uint32_t kChunk = SLOTS >= 4096u ? SLOTS / 2u : 7u * SLOTS / 10u;
      if (a != cur_a) {
        cur_a = a;
        A = ids + pa;
        na = na_l;
        hashed = na <= kChunk;
        if (hashed) build(0, na);
      }
      const uint32_t* B = ids + __shfl_sync(kFullMask, m_pb, l);
      const uint32_t nb = __shfl_sync(kFullMask, m_nb, l);
      int s = 0;
      if (hashed) {
        for (uint32_t i0 = lane; i0 < nb; i0 += 256) {
          if (i0 >= 256) {  // rows longer than the prefetched 256 ids
#pragma unroll
            for (int u = 0; u < 8; ++u) xs[u] = i0 + 32 * u < nb ? B[i0 + 32 * u] : kSentinel;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) s += probe(xs[u]);
        }
      } else {
        // row a does not fit the table: hash it chunk by chunk and stream row b against every
        // chunk (the id rows need not be sorted: the partitioned index leaves them in arrival order)
        for (uint32_t c0 = 0; c0 < na; c0 += kChunk) {
          build(c0, min(na, c0 + kChunk));
          for (uint32_t i = lane; i < nb; i += 32) s += probe(B[i]);
        }
        cur_a = kSentinel;  // the table holds the last chunk only
      }
      s = warp_sum_i(s);
      if (lane == l) m_score = (uint32_t)s;
    }
    if (my_e < n_edges) vals[my_e] = (m_val & 0xFFFFFFFFull) | ((unsigned long long)m_score << 32);
  }
}

// single-warp listing of the shared ids of two rows as k-mer values (ascending)
__global__ void shared_kmers_kernel(uint32_t ra, uint32_t rb, const uint32_t* __restrict__ pstart,
                                    const uint32_t* __restrict__ rowlen, const uint32_t* __restrict__ ids,
                                    const uint32_t* __restrict__ vocab, uint32_t* __restrict__ out,
                                    uint32_t cap, uint32_t* __restrict__ n_out) {
  const uint32_t lane = lane_id();
  const uint32_t* A = ids + pstart[ra];
  const uint32_t na = rowlen[ra];
  const uint32_t* B = ids + pstart[rb];
  const uint32_t nb = rowlen[rb];
  uint32_t base = 0;
  for (uint32_t c = 0; c < na; c += 32) {
    const uint32_t i = c + lane;
    bool hit = false;
    uint32_t x = 0;
    if (i < na) {
      x = A[i];
      uint32_t lo = 0, hi = nb;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (B[mid] < x) lo = mid + 1; else hi = mid;
      }
      hit = lo < nb && B[lo] == x;
    }
    const uint32_t m = __ballot_sync(kFullMask, hit);
    const uint32_t pos = base + __popc(m & lanemask_lt());
    if (hit && pos < cap) out[pos] = vocab[x];
    base += __popc(m);
  }
  if (lane == 0) *n_out = base;
}

// rank space -> input order, a < b, packed for the final sort:
//   key = a << 32 | b, val = count | blosum << 32
__global__ void finalize_edges_kernel(const uint4* __restrict__ edges, unsigned long long n_edges,
                                      const uint32_t* __restrict__ orig_of, unsigned long long* __restrict__ keys,
                                      unsigned long long* __restrict__ vals) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint4 e = edges[i];
    uint32_t a = orig_of ? orig_of[e.x] : e.x, b = orig_of ? orig_of[e.y] : e.y;
    if (a > b) {
      const uint32_t t = a;
      a = b;
      b = t;
    }
    keys[i] = ((unsigned long long)a << 32) | b;
    vals[i] = (unsigned long long)e.z | ((unsigned long long)e.w << 32);
  }
}

__global__ void assemble_edges_kernel(const unsigned long long* __restrict__ keys,
                                      const unsigned long long* __restrict__ vals, unsigned long long n_edges,
                                      uint4* __restrict__ out) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i], v = vals[i];
    out[i] = make_uint4((uint32_t)(k >> 32), (uint32_t)k, (uint32_t)v, (uint32_t)(v >> 32));
  }
}

// ---------------------------------------------------------------------------------------
// Dense presence-bitset path: per-protein bitsets over the repeated-k-mer vocabulary and
// AND+popcount pair counts (one warp per pair, shuffle reduction).
// ---------------------------------------------------------------------------------------
__global__ void bitset_fill_kernel(const uint32_t* __restrict__ rows, uint32_t n_rows,
                                   const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ rowlen,
                                   const uint32_t* __restrict__ ids, uint32_t words, uint32_t* __restrict__ bits) {
  const uint32_t lane = lane_id();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = gw; i < n_rows; i += nw) {
    const uint32_t r = rows[i];
    const uint32_t* A = ids + pstart[r];
    for (uint32_t j = lane; j < rowlen[r]; j += 32) {
      const uint32_t id = A[j];
      atomicOr(&bits[(size_t)i * words + (id >> 5)], 1u << (id & 31u));
    }
  }
}

// Dense tile of the pair matrix from presence bitsets: counts[i][j] = popc(bits_i & bits_j) summed over the
// vocabulary words.  INT / POPC-issue bound by construction: every warp keeps a 4 x 8 register tile of
// accumulators (its 4 rows against the CTA's 8 columns), so one 32-word slab costs 4 global loads and
// 8 shared-memory loads per lane against 32 x (LOP3.AND + POPC + IADD).  CTA = 8 warps = 32 rows x 8 columns.
constexpr int kBitCols = 8;          // columns per CTA (staged in shared memory)
constexpr int kBitRowsPerWarp = 4;   // rows per warp (registers)
constexpr int kBitWarps = 8;
constexpr int kBitRows = kBitWarps * kBitRowsPerWarp;  // rows per CTA
__global__ void __launch_bounds__(kBitWarps * 32)
    bitset_pairs_kernel(const uint32_t* __restrict__ bits, uint32_t n_rows, uint32_t words,
                        uint32_t* __restrict__ counts) {
  __shared__ uint32_t s_col[kBitCols][256];
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
  const uint32_t i0 = blockIdx.y * kBitRows + w * kBitRowsPerWarp, j0 = blockIdx.x * kBitCols;
  uint32_t acc[kBitRowsPerWarp][kBitCols];
#pragma unroll
  for (int r = 0; r < kBitRowsPerWarp; ++r)
#pragma unroll
    for (int q = 0; q < kBitCols; ++q) acc[r][q] = 0;
  for (uint32_t w0 = 0; w0 < words; w0 += 256) {
    __syncthreads();
    for (uint32_t x = threadIdx.x; x < kBitCols * 256; x += kBitWarps * 32) {
      const uint32_t q = x >> 8, ww = w0 + (x & 255u), j = j0 + q;
      s_col[q][x & 255u] = (j < n_rows && ww < words) ? bits[(size_t)j * words + ww] : 0u;
    }
    __syncthreads();
#pragma unroll 2
    for (int s = 0; s < 8; ++s) {
      const uint32_t ww = w0 + s * 32 + lane;
      uint32_t a[kBitRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kBitRowsPerWarp; ++r)
        a[r] = (i0 + r < n_rows && ww < words) ? __ldg(bits + (size_t)(i0 + r) * words + ww) : 0u;
#pragma unroll
      for (int q = 0; q < kBitCols; ++q) {
        const uint32_t bq = s_col[q][s * 32 + lane];
#pragma unroll
        for (int r = 0; r < kBitRowsPerWarp; ++r) acc[r][q] += __popc(a[r] & bq);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kBitRowsPerWarp; ++r)
#pragma unroll
    for (int q = 0; q < kBitCols; ++q) {
      const uint32_t v = warp_sum(acc[r][q]);
      if (lane == 0 && i0 + r < n_rows && j0 + q < n_rows) counts[(size_t)(i0 + r) * n_rows + j0 + q] = v;
    }
}

// POPC issue-rate microbenchmark (the roofline denominator of the bitset path): every thread runs
// `iters` x 8 independent AND + POPC + ADD chains on registers; nothing touches memory until the end.
__global__ void __launch_bounds__(256) popc_microbench_kernel(uint32_t iters, uint32_t seed, uint32_t* __restrict__ out) {
  uint32_t x[8], acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    x[q] = seed * 2654435761u + threadIdx.x * 40503u + q * 0x9E3779B9u + blockIdx.x;
    acc[q] = 0;
  }
  const uint32_t m = seed ^ 0x5A5A5A5Au;
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t p;
      asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(x[q] & (m + it)));
      acc[q] += p;
      x[q] += p;  // (keeps the chain data dependent: the compiler cannot hoist or fold it)
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += acc[q];
  if (s == 0xFFFFFFFFu) out[0] = s;  // never true in practice: keeps the loop alive
}

}  // namespace kc
