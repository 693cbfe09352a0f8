// bucket.cuh — K3/K4/K5 for sparse k-mer universes: the index built by PARTITIONING instead of
// by random access into universe-sized tables (index.cuh).
//
// Replaces (reference root relative), same stages as index.cuh:
//   merge_sort census                        src/main.rs:23-48,103-116
//   split unique/repeated + Mphf::new x2     src/main.rs:127-147
//   remove_unique_five_mers + modify_hash_five_mer + kmer_freq   src/protein.rs:151-174, src/main.rs:182-193
//   times_kmer_visited / triangular layout   src/graph/vertex.rs:92-136
//
// Flow (every random access is either a shared-memory access or an append through an L2-resident
// cursor, whose 8/16-byte writes merge into full sectors in L2 before they reach HBM):
//   1. The extract kernels (extract.cuh) append every (distinct k-mer, row) incidence to the
//      bucket its k-mer hashes to: n_buckets fixed-size slots of kBkCap records, one cursor each.
//   2. One CTA per bucket: hash-group the bucket in shared memory (the census: holders per
//      k-mer), give every repeated k-mer an id and a postings range, rank-sort the holders of
//      every k-mer, and emit postings (col), vocabulary (k-mer, freq, BLOSUM self-score per id)
//      and one entry {row, id, postings suffix of the holders after row} per (row, repeated
//      k-mer), appended to the entry bin of the row's 64-row block.
//   3. One CTA per entry bin: split the bin by row into the CSR the pair stage reads (ids /
//      suffix ranges / self-scores per row) and sum the per-row totals the row classifier needs.
// The ids are a minimal perfect hash of the repeated k-mers in bucket order (boomphf's ids are
// arbitrary too); the canonical (ascending k-mer) view is derived on demand.
#pragma once
#include "common.cuh"
#include "extract.cuh"
#include "index.cuh"

namespace kc {

// Bucket geometry: CAP records per slot = CAP hash slots (distinct <= incidences), CAP / 8 threads,
// mean fill 62 % of a slot.  4096 (two CTAs per SM) by default; 8192 (one CTA per SM) when a bucket
// of 4096 overflowed (a k-mer with thousands of holders).
constexpr uint32_t bk_target_fill(uint32_t cap) { return cap / 8u * 5u; }
constexpr size_t bk_smem_bytes(uint32_t cap) {
  return (size_t)cap * (4 + 4 + 4 + 2) + (size_t)cap * (4 + 2 + 4 + 2) + (size_t)(cap / 2) * 2;
}

// Entry bin of the rows [r0, r0 + 64): room for one entry per distinct k-mer of its rows plus one run
// record per two entries (1.5 x the capacity prefix)
__device__ __forceinline__ size_t bin_region(const uint32_t* __restrict__ rowcap_prefix, uint32_t r0) {
  const size_t p = rowcap_prefix[r0];
  return p + (p >> 1);
}

struct BucketGlobals {
  unsigned long long col_cursor;  // postings / entries written so far (= nnz at the end)
  unsigned long long id_cursor;   // ids handed out (= n_repeated at the end)
  // totals over the k-mers this build OWNS (the first holder is one of its rows; every k-mer
  // when the build is not sharded): summed over the ranks they are the whole-set numbers
  unsigned long long n_distinct, n_repeated, nnz, multi_total, work_total;
  unsigned long long n_records;   // records the buckets received (sizes the next build's bucket count)
  uint32_t overflow;              // some bucket was sent more than kBkCap records
  uint32_t max_bucket;
};

// ---------------------------------------------------------------------------------------
// The bucket kernel.  Shared memory per CTA (216 KB): the hash table of the bucket's k-mers
// (key, counter/cursor, id|self-score, holders), the records (row, slot) and the grouped +
// sorted holders.  The next bucket's records are prefetched into registers while the current
// one is processed.
// ---------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t bucket_slot_hash(uint32_t kmer) {
  uint32_t h = kmer ^ (kmer >> 15);
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

template <bool CROSS, uint32_t kBkCap>
__global__ void __launch_bounds__(kBkCap / 8, kBkCap == 4096 ? 2 : 1)
    bucket_build_kernel(const uint2* __restrict__ rec, const uint32_t* __restrict__ bucket_cnt, uint32_t n_buckets,
                        const uint32_t* __restrict__ first_after, int k, uint32_t* __restrict__ col,
                        uint4* __restrict__ entries, const uint32_t* __restrict__ rowcap_prefix,
                        uint32_t* __restrict__ bin_cursor, uint32_t* __restrict__ vocab,
                        uint32_t* __restrict__ freq, uint8_t* __restrict__ selfscore, RowOwner owner,
                        BucketGlobals* __restrict__ g) {
  // Pairs of two rows of the SAME 64-row bin are scored from "run records" (the holders of a
  // k-mer inside one bin as a 64-bit mask) on dense shared-memory tiles (pairs_tile_kernel); the
  // suffix of an entry then starts behind the row's bin, so the hash kernels only see the partners
  // of other bins.  Related proteins are usually neighbours in the input: most multi-edges are
  // bin-local and never touch a hash table or the postings.
  constexpr uint32_t kBkSlots = kBkCap;
  constexpr int kBkThreads = kBkCap / 8;
  constexpr int kBkPerThread = 8;
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  uint32_t* s_key = reinterpret_cast<uint32_t*>(dyn_smem);      // [slots] k-mer
  uint32_t* s_val = s_key + kBkSlots;                           // [slots] holders -> cursor -> group end
  uint32_t* s_meta = s_val + kBkSlots;                          // [slots] local id | self-score << 16
  uint32_t* s_row = s_meta + kBkSlots;                          // [cap] record rows -> sorted holders
  uint32_t* s_col = s_row + kBkCap;                             // [cap] grouped holders (arrival order)
  uint16_t* s_cnt = reinterpret_cast<uint16_t*>(s_col + kBkCap);  // [slots] holders
  uint16_t* s_slot = s_cnt + kBkSlots;                          // [cap] record slots -> suffix starts (CROSS)
  uint16_t* s_grp = s_slot + kBkCap;                            // [cap] slot of every grouped holder
  uint16_t* s_rep = s_grp + kBkCap;                             // [cap / 2] slot of every repeated k-mer, by local id
  __shared__ uint32_t s_wsum[32];
  __shared__ unsigned long long s_base[2];
  __shared__ uint32_t s_nnz;
  const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  unsigned long long multi = 0, work = 0;
  uint32_t n_distinct = 0, n_rep_owned = 0, nnz_owned = 0;  // per thread: far below 2^32
  uint32_t max_bucket = 0;
  unsigned long long n_records = 0;

  uint2 nxt[kBkPerThread];
  uint32_t b = blockIdx.x;
  uint32_t n_cur = 0;
  if (b < n_buckets) {
    n_cur = bucket_cnt[b];
    const uint2* src = rec + (size_t)b * kBkCap;
#pragma unroll
    for (int j = 0; j < kBkPerThread; ++j) {
      const uint32_t i = tid + j * kBkThreads;
      if (i < n_cur && n_cur <= kBkCap) nxt[j] = ld_stream_u32x2(src + i);
    }
  }
  for (; b < n_buckets; b += gridDim.x) {
    const uint32_t nrec = n_cur;
    max_bucket = max(max_bucket, nrec);
    if (tid == 0) n_records += nrec;
    // ---- P0: clear the table
#pragma unroll
    for (int j = 0; j < (int)(kBkSlots / kBkThreads); ++j) {
      const uint32_t s = tid + j * kBkThreads;
      s_key[s] = kSentinel;
      s_val[s] = 0;
    }
    __syncthreads();
    const bool ok = nrec <= kBkCap;
    // ---- P1: insert this bucket's records (from registers), prefetch the next bucket's
    if (ok) {
#pragma unroll
      for (int j = 0; j < kBkPerThread; ++j) {
        const uint32_t i = tid + j * kBkThreads;
        if (i < nrec) {
          const uint32_t km = nxt[j].x;
          uint32_t h = bucket_slot_hash(km) & (kBkSlots - 1u);
          for (;;) {
            const uint32_t cur = s_key[h];
            if (cur == km) break;
            if (cur == kSentinel) {
              const uint32_t old = atomicCAS(&s_key[h], kSentinel, km);
              if (old == kSentinel || old == km) break;
            }
            h = (h + 1u) & (kBkSlots - 1u);
          }
          atomicAdd(&s_val[h], 1u);
          s_slot[i] = (uint16_t)h;
          s_row[i] = nxt[j].y;
        }
      }
    } else if (tid == 0) {
      atomicOr(&g->overflow, 1u);
    }
    {
      const uint32_t bn = b + gridDim.x;
      if (bn < n_buckets) {
        n_cur = bucket_cnt[bn];
        const uint2* src = rec + (size_t)bn * kBkCap;
#pragma unroll
        for (int j = 0; j < kBkPerThread; ++j) {
          const uint32_t i = tid + j * kBkThreads;
          if (i < n_cur && n_cur <= kBkCap) nxt[j] = ld_stream_u32x2(src + i);
        }
      }
    }
    __syncthreads();
    if (!ok) continue;  // uniform: the host falls back to the universe-table index
    // ---- P2: every thread owns kBkSlots / kBkThreads slots (strided: no bank conflicts); block
    // scan of (repeated k-mers, their holders) packed as holders << 16 | repeated
    constexpr int SPT = kBkSlots / kBkThreads;
    uint32_t cnts[SPT];
    uint32_t local = 0;
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const uint32_t s = tid + j * kBkThreads;
      const uint32_t c = s_key[s] != kSentinel ? s_val[s] : 0u;
      cnts[j] = c;
      s_cnt[s] = (uint16_t)c;
      if (c >= 2u) local += (c << 16) | 1u;
    }
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const uint32_t x = lane < (uint32_t)(kBkThreads / 32) ? s_wsum[lane] : 0u;
      uint32_t xs = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, xs, o);
        if (lane >= (uint32_t)o) xs += t;
      }
      s_wsum[lane] = xs - x;
      if (lane == 31) {  // block totals: reserve the bucket's ids and postings
        s_nnz = xs;
        s_base[0] = atomicAdd(&g->col_cursor, (unsigned long long)(xs >> 16));
        s_base[1] = atomicAdd(&g->id_cursor, (unsigned long long)(xs & 0xFFFFu));
      }
    }
    __syncthreads();
    const uint32_t excl = incl - local + s_wsum[warp];
    const unsigned long long col_base = s_base[0], id_base = s_base[1];
    const uint32_t nnz = s_nnz >> 16, nrep = s_nnz & 0xFFFFu;
    {
      uint32_t lid = excl & 0xFFFFu, coff = excl >> 16;
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        const uint32_t c = cnts[j];
        if (c >= 2u) {
          const uint32_t s = tid + j * kBkThreads;
          s_val[s] = coff;
          s_rep[lid] = (uint16_t)s;
          ++lid;
          coff += c;
        }
      }
    }
    __syncthreads();
    // vocabulary of the bucket, one thread per repeated k-mer (dense: the self-score costs k
    // divisions, and the global writes are coalesced)
    for (uint32_t l = tid; l < nrep; l += kBkThreads) {
      const uint32_t s = s_rep[l];
      const uint32_t km = s_key[s];
      const uint32_t ss = (uint32_t)kmer_self_score(km, k);
      s_meta[s] = l | (ss << 16);
      vocab[id_base + l] = km;
      freq[id_base + l] = s_cnt[s];
      selfscore[id_base + l] = (uint8_t)ss;
    }
    // ---- P3: group the holders of every repeated k-mer
#pragma unroll
    for (int j = 0; j < kBkPerThread; ++j) {
      const uint32_t i = tid + j * kBkThreads;
      if (i < nrec) {
        const uint32_t s = s_slot[i];
        const uint32_t c = s_cnt[s];
        if (c >= 2u) {
          const uint32_t pos = atomicAdd(&s_val[s], 1u);
          s_col[pos] = s_row[i];
          s_grp[pos] = (uint16_t)s;
        } else {  // a k-mer with one holder: owned by that holder's rank
          const uint32_t row = s_row[i];
          n_distinct += owner.mine(row);
        }
      }
    }
    __syncthreads();
    // ---- P4: rank-sort every group (holders are distinct rows): s_row[start + rank] = holder.
    // The same pass over the group finds the run of holders inside the row's 64-row bin; the
    // suffix of a holder starts behind that run (CROSS: and not before the first holder of a later
    // class block).  s_slot[sorted position] = suffix start | (first holder of a run of >= 2) << 15.
#pragma unroll
    for (int j = 0; j < kBkPerThread; ++j) {
      const uint32_t p = tid + j * kBkThreads;
      if (p < nnz) {
        const uint32_t s = s_grp[p];
        const uint32_t end = s_val[s], start = end - s_cnt[s];
        const uint32_t v = s_col[p];
        const uint32_t vbin = v >> kBinRowsLog;
        uint32_t target = 0;
        if (CROSS) target = first_after[v];
        uint32_t rank = 0, below = 0, bins_lt = 0, bins_le = 0;
        for (uint32_t q = start; q < end; ++q) {
          const uint32_t x = s_col[q];
          rank += x < v;
          bins_lt += (x >> kBinRowsLog) < vbin;
          bins_le += (x >> kBinRowsLog) <= vbin;
          if (CROSS) below += x < target;
        }
        s_row[start + rank] = v;
        const uint32_t a = CROSS ? max(bins_le, below) : bins_le;  // relative to the group start
        const bool run_leader = rank == bins_lt && bins_le - bins_lt >= 2u;
        s_slot[start + rank] = (uint16_t)((start + a) | (run_leader ? 0x8000u : 0u));
      }
    }
    __syncthreads();
    // ---- P5: emit postings (contiguous per bucket) and entries (appended to the row block's bin;
    // lanes that hit the same bin share one reservation, and four reservations per thread are in
    // flight before the first entry is stored)
#pragma unroll
    for (int j0 = 0; j0 < kBkPerThread; j0 += 4) {
      uint4 ent[4];
      unsigned long long rmask[4];  // holders of the k-mer in the row's bin (set by the run's first holder)
      uint32_t who[4], pend[4];     // who = leader lane << 16 | slot among the records the bin's lanes append
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t q = tid + (j0 + u) * kBkThreads;
        const bool act = q < nnz;
        uint32_t bin = kSentinel;
        bool has_run = false;
        ent[u].x = kSentinel;
        rmask[u] = 0;
        if (act) {
          const uint32_t s = s_grp[q];
          const uint32_t end = s_val[s];
          const uint32_t row = s_row[q];
          const bool mine = owner.mine(row);
          col[col_base + q] = row;
          const uint32_t c = s_cnt[s];
          const uint32_t start = end - c;
          if (mine && q == start) {  // first holder: this build owns the k-mer's totals
            ++n_distinct;
            ++n_rep_owned;
            nnz_owned += c;
            multi += (unsigned long long)c * (c - 1u) / 2u;
          }
          bin = row >> kBinRowsLog;
          const uint32_t sl = s_slot[q];
          const uint32_t a = sl & 0x7FFFu;  // the partners behind the row's bin (P4)
          const uint32_t meta = s_meta[s];
          if (mine && (sl & 0x8000u)) {  // first holder of a run of >= 2 holders inside the bin: its record
            for (uint32_t p2 = q; p2 < end && (s_row[p2] >> kBinRowsLog) == bin; ++p2)
              rmask[u] |= 1ull << (s_row[p2] & (kBinRows - 1u));
            has_run = true;
          }
          // Every holder gets its entry, also the rows of other shards that passed the filter: the
          // pair stage scores this build's rows only, but the BLOSUM pass over unscored edges
          // (edge_blosum_kernel) intersects the id lists of BOTH endpoints, and the k-mers a foreign
          // row shares with this build's rows are exactly the ones that passed.
          const uint32_t len = end - a;
          // a single partner is stored inline ({rank, sentinel}): no postings gather in the pair stage
          const uint2 sf = len == 1u ? make_uint2(s_row[a], kSentinel)
                                     : make_uint2((uint32_t)col_base + a, (uint32_t)col_base + end);
          ent[u] = make_uint4(row | ((meta >> 16) << 24), (uint32_t)id_base + (meta & 0xFFFFu), sf.x, sf.y);
          if (mine) work += len;
        }
        const uint32_t peers = __match_any_sync(kFullMask, bin);
        const uint32_t runs = __ballot_sync(kFullMask, has_run) & peers;
        const uint32_t leader = __ffs(peers) - 1;
        who[u] = (leader << 16) | (__popc(peers & lanemask_lt()) + __popc(runs & lanemask_lt()));
        pend[u] = 0;
        if (bin != kSentinel && lane == leader)
          pend[u] = atomicAdd(&bin_cursor[bin], (uint32_t)(__popc(peers) + __popc(runs)));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t base = __shfl_sync(kFullMask, pend[u], who[u] >> 16);
        if (ent[u].x != kSentinel) {
          const uint32_t r0 = (ent[u].x & 0xFFFFFFu) & ~(kBinRows - 1u);
          uint4* dst = entries + bin_region(rowcap_prefix, r0) + base + (who[u] & 0xFFFFu);
          dst[0] = ent[u];
          if (rmask[u])  // run record: {tag | self-score, bin, mask}
            dst[1] = make_uint4(0x80000000u | (ent[u].x >> 24), r0 >> kBinRowsLog, (uint32_t)rmask[u],
                                (uint32_t)(rmask[u] >> 32));
        }
      }
    }
    __syncthreads();
  }
  const unsigned long long n_dist64 = warp_sum64(n_distinct), n_rep64 = warp_sum64(n_rep_owned),
                           nnz64 = warp_sum64(nnz_owned);
  multi = warp_sum64(multi);
  work = warp_sum64(work);
  if (lane == 0) {
    if (n_dist64) atomicAdd(&g->n_distinct, n_dist64);
    if (n_rep64) atomicAdd(&g->n_repeated, n_rep64);
    if (nnz64) atomicAdd(&g->nnz, nnz64);
    if (multi) atomicAdd(&g->multi_total, multi);
    if (work) atomicAdd(&g->work_total, work);
    atomicMax(&g->max_bucket, max_bucket);
    if (n_records) atomicAdd(&g->n_records, n_records);
  }
}

// sharded build: distinct k-mers of this rank's rows (its share of n_incidences)
__global__ void own_incidences_kernel(const uint32_t* __restrict__ ndist, uint32_t n, RowOwner owner,
                                      unsigned long long* __restrict__ total) {
  unsigned long long s = 0;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    if (owner.mine(r)) s += ndist[r];
  s = warp_sum64(s);
  if (lane_id() == 0 && s) atomicAdd(total, s);
}

// distinct k-mers of the rows before row r (entry capacity of the rows' bins): prefix[r], prefix[n]
struct RowCapOut {
  uint32_t* p;
  uint64_t n;
  __device__ void operator()(uint64_t i, unsigned long long excl, unsigned long long v) const {
    p[i] = (uint32_t)excl;
    if (i + 1 == n) p[n] = (uint32_t)(excl + v);
  }
};

// ---------------------------------------------------------------------------------------
// One CTA per entry bin (64 rows): ONE pass over the bin splits it by row into the arrays the
// pair stage reads (ids / suffix ranges / self-scores).  The bin is taken in chunks of 4096
// entries that are counting-sorted by row in shared memory first, so that the global stores
// are runs of consecutive entries of one row (full sectors) instead of one scattered 4/8/1-byte
// store per entry.  The rows are laid out by capacity (row r starts at rowcap_prefix[r], room
// for all its distinct k-mers; the pair stage reads rowlen[r] entries from there), so no
// counting pass over the whole bin and no scan are needed.  Also sums what the pair stage's row
// classifier needs (rowwork = multi-edges of the row, rowinl = inline partners, rowmaxlen =
// longest suffix; suffix_ranges_kernel in index.cuh produces the same numbers for the table build).
// ---------------------------------------------------------------------------------------
constexpr int kFinThreads = 512;
constexpr int kFinPer = 8;
constexpr uint32_t kFinChunk = kFinThreads * kFinPer;
constexpr size_t kFinSmemBytes = (size_t)kFinChunk * 16;
__global__ void __launch_bounds__(kFinThreads)
    rows_finalize_kernel(const uint4* __restrict__ entries, const uint32_t* __restrict__ rowcap_prefix,
                         const uint32_t* __restrict__ bin_cnt, uint32_t n, uint32_t bin_lo, uint32_t n_bins,
                         uint4* __restrict__ runs_out, uint32_t* __restrict__ run_cnt,
                         uint32_t* __restrict__ rowlen, uint32_t* __restrict__ rowlen_pair,
                         uint32_t* __restrict__ ids, uint2* __restrict__ suf,
                         uint8_t* __restrict__ sufss, unsigned long long* __restrict__ rowwork64,
                         uint32_t* __restrict__ rowwork, uint32_t* __restrict__ rowinl,
                         uint32_t* __restrict__ rowmaxlen) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  uint4* s_ent = reinterpret_cast<uint4*>(dyn_smem);
  // per row of the bin: id-list cursor / chunk count / chunk start, the same for the pair list
  // (entries with partners only), and the row totals
  __shared__ uint32_t s_off[kBinRows], s_cnt[kBinRows], s_start[kBinRows];
  __shared__ uint32_t s_offp[kBinRows], s_cntp[kBinRows], s_startp[kBinRows];
  __shared__ uint32_t s_inl[kBinRows], s_max[kBinRows];
  __shared__ uint32_t s_wlo[kBinRows], s_whi[kBinRows];  // multi-edges of the row: low word + carries
  __shared__ uint32_t s_nruns;
  const uint32_t tid = threadIdx.x;
  for (uint32_t bin = bin_lo + blockIdx.x; bin < n_bins; bin += gridDim.x) {  // [bin_lo, n_bins): this build's rows
    const uint32_t r0 = bin << kBinRowsLog;
    const uint32_t nrows = min(kBinRows, n - r0);
    const uint4* src = entries + bin_region(rowcap_prefix, r0);
    uint4* rdst = runs_out + (rowcap_prefix[r0] >> 1);  // run records of the bin, compacted for pairs_tile_kernel
    const uint32_t cnt = bin_cnt[bin];
    if (tid == 0) s_nruns = 0;
    if (tid < kBinRows) {
      s_off[tid] = s_offp[tid] = tid < nrows ? rowcap_prefix[r0 + tid] : 0u;
      s_cnt[tid] = s_cntp[tid] = 0;
      s_inl[tid] = 0;
      s_max[tid] = 0;
      s_wlo[tid] = 0;
      s_whi[tid] = 0;
    }
    __syncthreads();
    for (uint32_t c0 = 0; c0 < cnt; c0 += kFinChunk) {
      const uint32_t m = min(kFinChunk, cnt - c0);
      uint4 e[kFinPer];
      uint32_t rk[kFinPer];  // row << 26 | rank in the row's pair list (0x1FFF: none) << 13 | rank in its id list
#pragma unroll
      for (int u = 0; u < kFinPer; ++u)
        if (u * kFinThreads + tid < m) e[u] = ld_stream_u32x4(src + c0 + u * kFinThreads + tid);
#pragma unroll
      for (int u = 0; u < kFinPer; ++u) {
        rk[u] = kSentinel;
        if (u * kFinThreads + tid >= m) continue;
        if (e[u].x >> 31) {  // run record
          rdst[atomicAdd(&s_nruns, 1u)] = e[u];
          continue;
        }
        const uint32_t lr = (e[u].x & 0xFFFFFFu) - r0;
        const bool inl = e[u].w == kSentinel;
        const uint32_t len = inl ? 1u : e[u].w - e[u].z;
        uint32_t rp = 0x1FFFu;
        if (len) {  // the pair stage only reads the entries that have partners
          rp = atomicAdd(&s_cntp[lr], 1u);
          const uint32_t old = atomicAdd(&s_wlo[lr], len);
          if (old + len < old) atomicAdd(&s_whi[lr], 1u);
          if (inl) atomicAdd(&s_inl[lr], 1u);
          else if (len > 1u) atomicMax(&s_max[lr], len);
        }
        rk[u] = (lr << 26) | (rp << 13) | atomicAdd(&s_cnt[lr], 1u);
      }
      __syncthreads();
      if (tid < 64) {  // exclusive scans of the 64 per-row counts of the chunk: warp 0 ids, warp 1 pair lists
        const uint32_t l = tid & 31u;
        const uint32_t* cntv = tid < 32 ? s_cnt : s_cntp;
        uint32_t* startv = tid < 32 ? s_start : s_startp;
        const uint32_t a = cntv[2 * l], b = cntv[2 * l + 1];
        uint32_t incl = a + b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
          if (l >= (uint32_t)o) incl += t;
        }
        startv[2 * l] = incl - a - b;
        startv[2 * l + 1] = incl - b;
      }
      __syncthreads();
      // id lists: every entry, sorted by row in shared memory, written as runs
#pragma unroll
      for (int u = 0; u < kFinPer; ++u)
        if (rk[u] != kSentinel) s_ent[s_start[rk[u] >> 26] + (rk[u] & 0x1FFFu)] = e[u];
      __syncthreads();
      const uint32_t m_rows = s_start[kBinRows - 1] + s_cnt[kBinRows - 1];  // the chunk's entries without the runs
      for (uint32_t i = tid; i < m_rows; i += kFinThreads) {
        const uint4 v = s_ent[i];
        const uint32_t lr = (v.x & 0xFFFFFFu) - r0;
        ids[s_off[lr] + (i - s_start[lr])] = v.y;
      }
      __syncthreads();
      // pair lists: the entries with partners
#pragma unroll
      for (int u = 0; u < kFinPer; ++u)
        if (rk[u] != kSentinel && ((rk[u] >> 13) & 0x1FFFu) != 0x1FFFu)
          s_ent[s_startp[rk[u] >> 26] + ((rk[u] >> 13) & 0x1FFFu)] = e[u];
      __syncthreads();
      const uint32_t m_pairs = s_startp[kBinRows - 1] + s_cntp[kBinRows - 1];
      for (uint32_t i = tid; i < m_pairs; i += kFinThreads) {
        const uint4 v = s_ent[i];
        const uint32_t lr = (v.x & 0xFFFFFFu) - r0;
        const uint32_t pos = s_offp[lr] + (i - s_startp[lr]);
        suf[pos] = make_uint2(v.z, v.w);
        if (sufss) sufss[pos] = (uint8_t)(v.x >> 24);
      }
      __syncthreads();
      if (tid < kBinRows) {
        s_off[tid] += s_cnt[tid];
        s_offp[tid] += s_cntp[tid];
        s_cnt[tid] = s_cntp[tid] = 0;
      }
      __syncthreads();
    }
    if (tid < nrows) {
      const uint32_t r = r0 + tid;
      const unsigned long long w = ((unsigned long long)s_whi[tid] << 32) | s_wlo[tid];
      rowlen[r] = s_off[tid] - rowcap_prefix[r];
      rowlen_pair[r] = s_offp[tid] - rowcap_prefix[r];
      rowwork64[r] = w;
      rowwork[r] = w > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)w;
      rowinl[r] = s_inl[tid];
      rowmaxlen[r] = s_max[tid];
    }
    if (tid == 0) run_cnt[bin] = s_nruns;
    __syncthreads();
  }
}

}  // namespace kc
