// engine.cu — C ABI (include/kc_b200.h) over the sm_100a kernels.  One engine per GPU.
// No CPU fallback: every compute entry point launches CUDA kernels or fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/kc_b200.h"
#include "common.cuh"
#include "extract.cuh"
#include "index.cuh"
#include "bucket.cuh"
#include "stream_index.cuh"
#include "pairs.cuh"
#include "dist.cuh"
#include "primitives.cuh"

using namespace kc;

namespace {

struct DBuf {  // grow-only device buffer
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t want) {
    if (want <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    want = (want + 255) & ~size_t(255);
    cudaError_t rc = cudaMalloc(&p, want);
    if (rc == cudaSuccess) bytes = want;
    return rc;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

struct DeviceScalars {  // one small block of u64 counters, zeroed per stage
  unsigned long long n_incid, n_distinct, n_repeated, nnz, work_total, multi_total, scan_total;
  unsigned long long edge_cursor;
  PairCounters pc;
  uint32_t list_counts[4];
  uint32_t row_cursor[32];
  uint32_t bin_counts[32];
  uint32_t n_overflow;
  uint32_t pad2;
  uint32_t shard_rows[2];
  uint32_t n_shared;
  uint32_t pad;
  BucketGlobals bg;
  uint32_t ent_seg[2];  // {0, nnz}: the one segment of the first entry-partition pass
  uint32_t sx_mid_cnt, sx_huge_cnt;  // streaming build: buckets beyond a warp's / a CTA's shared memory
};

enum Ev { EV_H2D0, EV_H2D1, EV_X0, EV_X1, EV_I0, EV_IC0, EV_IC1, EV_I1, EV_P0, EV_PK0, EV_PK1, EV_P1, EV_E1, EV_D0, EV_D1, EV_COUNT };

}  // namespace

struct kc_engine {
  kc_config cfg{};
  int dev = 0;
  int num_sm = kNumSM;
  size_t smem_optin = 0;
  size_t total_mem = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // chunked upload of the residue stream on a second stream: the extract kernels of the next
  // kc_build_index start on the first proteins while the rest is still crossing PCIe
  static constexpr int kUploadChunks = 8;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t chunk_ev[kUploadChunks]{}, main_ev = nullptr;
  uint32_t chunk_row[kUploadChunks + 1]{};
  uint64_t chunk_byte[kUploadChunks + 1]{};  // chunk c = bytes [chunk_byte[c], chunk_byte[c+1]) (tile-aligned cuts)
  int n_chunks = 0;  // > 0: an upload is (possibly) in flight; chunk c holds the rows [chunk_row[c], chunk_row[c+1])
  std::string err;
  uint32_t launches = 0;
  cudaEvent_t ev[EV_COUNT]{};
  bool ev_set[EV_COUNT]{};

  // proteins
  uint64_t n = 0, R = 0;
  bool have_proteins = false, have_index = false, have_pairs = false;
  std::vector<uint64_t> h_off;
  std::vector<uint32_t> h_cls, h_orig, h_rank, h_first_after, h_pstart, h_plen;
  std::vector<uint32_t> h_long, h_huge;
  std::vector<uint32_t> h_cta, h_vlong;  // partitioned build: CTA-hashed rows (717..2800 positions), sorted rows beyond
  std::vector<unsigned long long> h_huge_off;
  uint32_t max_block_np2 = 0, max_block_len = 0;
  uint32_t n_mid_rows = 0;  // proteins of kHashMaxPos+1 .. kWarpMaxPos positions
  std::vector<unsigned long long> h_pospref;  // k-mer positions (after subsampling) of the rows before row r
  unsigned long long n_pos_unsampled = 0;
  uint32_t max_plen = 0;
  unsigned long long huge_total = 0;  // scratch words of the rows beyond a CTA's shared memory
  // Deferred host copies (kc_set_proteins from host buffers, pair order = input order): the caller's arrays
  // stay valid until the next kc_build_index* / kc_extract_kmers returns, so the engine's own copies of the
  // offsets and classes (h_off, h_cls: readbacks, kc_extract_kmers, shard filters) are made by ensure_staged()
  // - inside the build, after its kernels are queued - instead of on the critical path in front of them.
  // The row tables (h_pospref, the row lists) are derived from the caller's array directly.
  const uint64_t* pend_off = nullptr;
  const uint32_t* pend_cls = nullptr;
  bool retry_full_buckets = false;
  uint32_t try_cap = 0;  // bucket slot size of the running attempt (0: choose)
  // the slot size that worked for the protein set with this signature (a k-mer with thousands of
  // holders overflows 4096-record buckets: start with 8192 next time)
  uint32_t cap_hint = 4096;
  uint64_t cap_hint_n = 0, cap_hint_R = 0;
  DBuf d_res, d_off, d_kpos, d_pstart, d_plen, d_orig, d_rank, d_first_after, d_long, d_huge, d_huge_off,
      d_huge_scratch, d_cta, d_vlong;
  // index
  uint32_t universe = 0;
  uint64_t n_words = 0;
  kc_index_stats istats{};
  uint64_t multi_total = 0, work_total = 0;
  DBuf d_pk, d_ndist, d_rowlen, d_seen, d_dict, d_vocab, d_freq, d_self, d_colptr, d_cursor, d_col,
      d_suf, d_sufss, d_rowwork, d_lists, d_colscratch, d_workprefix, d_ksplit, d_isplit, d_rowwork64, d_rowinl, d_rowmaxlen, d_psplit, d_rowbase, d_plist, d_pss;
  bool have_plist = false;
  uint32_t slice_shift = 31, n_slices = 1;
  // partitioned index (bucket.cuh): the pair stage reads d_rowcap / d_ids / d_self_h instead of
  // d_pstart / d_pk / d_self; the canonical view (legacy arrays) is derived on demand
  bool bucketed = false, canonical_ready = false;
  // sharded build (kc_build_index_shard with n_shards > 1): this engine holds the index of the rows
  // of rank `ishard` only (two row blocks of the pair order, see block_bounds)
  uint32_t ishard = 0, ishards = 1;
  std::vector<uint32_t> block_bounds;  // n_blocks + 1 rows of the pair order (2 blocks per rank, zig-zag)
  std::vector<uint8_t> h_binowner;
  uint32_t n_own_rows = 0;
  uint64_t v_local = 0;    // ids handed out by this build (= n_repeated when not sharded)
  uint64_t kept_hint = 0;  // records the last sharded build kept (sizes the bucket count of the next one)
  uint32_t hint_shards = 0;
  DBuf d_filter, d_binowner, d_runs, d_run_cnt, d_rowlen_p;
  // rows of the pair lists (entries with partners): shorter than the id rows once bin-local pairs go to the tiles
  const uint32_t* pair_rowlen() const { return (bucketed ? d_rowlen_p : d_rowlen).as<uint32_t>(); }
  RowOwner owner() const {
    return RowOwner{ishards > 1 ? d_binowner.as<uint8_t>() : nullptr, ishard};
  }
  DBuf d_rec, d_entries, d_bin_cnt, d_rowcap, d_bucket_cnt, d_ids, d_vocab_h, d_freq_h, d_self_h,
      d_zero, d_rowlen_c, d_islo_c;
  // streaming partitioned build (stream_index.cuh): the residue stream in the pair order (an alias of d_res
  // unless the pair order is class-major), its row starts, the two record arrays and the scanned histograms
  bool streamed = false;
  DBuf d_sres_own, d_soff, d_rec_a, d_rec_b, d_h1, d_h2, d_huge_list, d_mid_list, d_tile_row, d_ss3, d_keepmask;
  std::vector<uint32_t> h_soff;
  uint32_t sx_huge_last = 0, sx_mid_last = 0, sx_max_bucket = 0;
  double bitset_word_ops = 0;  // AND+POPC+ADD word operations of the last kc_bitset_pair_counts
  uint64_t index_records = 0;  // records the last streaming build partitioned (a sharded build: what it kept)
  uint32_t rank_of(uint64_t p) const { return cfg.cross_class_only ? h_rank[p] : (uint32_t)p; }
  const uint8_t* sres() const { return cfg.cross_class_only ? d_sres_own.as<uint8_t>() : d_res.as<uint8_t>(); }
  const uint32_t* pair_rowptr() const { return (bucketed ? d_rowcap : d_pstart).as<uint32_t>(); }  // d_rowcap: capacity prefix
  const uint32_t* pair_ids() const { return (bucketed ? d_ids : d_pk).as<uint32_t>(); }
  const uint8_t* pair_self() const { return (bucketed ? d_self_h : d_self).as<uint8_t>(); }
  const uint32_t* canon_rowlen() const { return (bucketed ? d_rowlen_c : d_rowlen).as<uint32_t>(); }
  // pairs
  DBuf d_rowbin, d_rowsafe, d_rowlogh, d_edges, d_keys_a, d_keys_b, d_vals_a, d_vals_b, d_hist, d_edges_sorted;
  uint64_t edge_cap = 0, n_edges = 0;
  kc_pair_stats pstats{};
  // multi-GPU (dist.cuh): this engine's NCCL rank
  ncclComm_t comm = nullptr;
  int crank = 0, cworld = 1;
  DBuf d_comm, d_gather;
  // misc
  DBuf d_scalars, d_scan_tiles, d_tmp;
  ScanScratch scan;
  DeviceScalars* ds = nullptr;
};

namespace {

#define KC_CUDA(e, call)                                                                      \
  do {                                                                                        \
    cudaError_t rc__ = (call);                                                                \
    if (rc__ != cudaSuccess) {                                                                \
      (e)->err = std::string(#call) + ": " + cudaGetErrorString(rc__);                        \
      return rc__ == cudaErrorMemoryAllocation ? KC_ENOMEM : KC_ECUDA;                        \
    }                                                                                         \
  } while (0)

#define KC_LAUNCH(e, kernel, grid, block, smem, ...)                \
  do {                                                              \
    ++(e)->launches;                                                \
    kernel<<<(grid), (block), (smem), (e)->stream>>>(__VA_ARGS__);  \
  } while (0)

int fail(kc_engine* e, int code, const std::string& msg) {
  e->err = msg;
  return code;
}

void mark(kc_engine* e, Ev which) {
  cudaEventRecord(e->ev[which], e->stream);
  e->ev_set[which] = true;
}

float elapsed(kc_engine* e, Ev a, Ev b) {
  if (!e->ev_set[a] || !e->ev_set[b]) return 0.f;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, e->ev[a], e->ev[b]) != cudaSuccess) {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

// the main stream waits for the whole chunked upload (paths that do not consume it chunk by chunk)
int wait_upload(kc_engine* e) {
  if (e->n_chunks) {
    KC_CUDA(e, cudaStreamWaitEvent(e->stream, e->chunk_ev[e->n_chunks - 1], 0));
    e->n_chunks = 0;
  }
  return KC_OK;
}

int ensure_scan(kc_engine* e, uint64_t n_items) {
  const uint64_t tiles = (n_items + kScanTile - 1) / kScanTile + 1;
  if (tiles > e->scan.cap_tiles) {
    KC_CUDA(e, e->d_scan_tiles.ensure(tiles * 8));
    e->scan.tile_sums = e->d_scan_tiles.as<unsigned long long>();
    e->scan.cap_tiles = tiles;
  }
  e->scan.total = &e->ds->scan_total;
  return KC_OK;
}

uint32_t blocks_for(uint64_t items, uint32_t per_block, uint32_t cap) {
  uint64_t b = (items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return (uint32_t)std::min<uint64_t>(b, cap);
}

// pair order = input order (all-classes mode): the row layout straight from the offsets on the device
__global__ void layout_identity_kernel(const unsigned long long* __restrict__ off, uint32_t n,
                                       uint32_t* __restrict__ pstart, uint32_t* __restrict__ plen,
                                       uint32_t* __restrict__ orig, uint32_t* __restrict__ rank,
                                       uint32_t* __restrict__ soff) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const unsigned long long a = off[r], b = off[r + 1];
  soff[r] = (uint32_t)a;
  if (r + 1 == n) soff[n] = (uint32_t)b;
  pstart[r] = (uint32_t)a;
  plen[r] = (uint32_t)(b - a);
  orig[r] = r;
  rank[r] = r;
}

// residue buffers are padded with zeros to whole tiles of the streaming kernels plus one tile (their 16-byte
// loads and the k-mer halo never need a bounds check)
size_t padded_res_bytes(uint64_t R) { return (size_t)((R + kSxTile - 1) / kSxTile + 1) * kSxTile; }

// ---- host-side staging shared by both kc_set_proteins flavours ---------------------------
// (this runs while the residue stream is crossing PCIe: one pass over the offsets)
// host part: pair order, position prefix, row lists (needs h_off / h_cls)
// `off`: the engine's own copy, or the caller's array while that copy is still deferred (kc_engine::pend_off);
// cross-class mode reads the classes from h_cls.  Checks that the offsets do not decrease.
int stage_layout_host(kc_engine* e, const uint64_t* off) {
  const uint64_t n = e->n;
  const int k = e->cfg.k;
  if (off[0] != 0) return fail(e, KC_EINVAL, "offsets[0] must be 0");
  e->R = off[n];
  if (e->R >= 0xFFFF0000ull) return fail(e, KC_ETOOLARGE, "more than 2^32-65536 residues");
  if (n >= 0xFFFFFFF0ull) return fail(e, KC_ETOOLARGE, "too many proteins");
  // pair order: input order, or class-major (stable) when only cross-class pairs are wanted,
  // so that the same-class holders a row must skip are one contiguous run of every posting
  const bool cross = e->cfg.cross_class_only != 0;
  e->h_first_after.clear();
  if (cross) {
    e->h_orig.resize(n);
    e->h_rank.resize(n);
    std::iota(e->h_orig.begin(), e->h_orig.end(), 0u);
    std::stable_sort(e->h_orig.begin(), e->h_orig.end(),
                     [&](uint32_t a, uint32_t b) { return e->h_cls[a] < e->h_cls[b]; });
    e->h_first_after.resize(n);
    uint64_t i = 0;
    while (i < n) {
      uint64_t j = i;
      while (j < n && e->h_cls[e->h_orig[j]] == e->h_cls[e->h_orig[i]]) ++j;
      for (uint64_t r = i; r < j; ++r) e->h_first_after[r] = (uint32_t)j;
      i = j;
    }
    for (uint64_t r = 0; r < n; ++r) e->h_rank[e->h_orig[r]] = (uint32_t)r;
    e->h_pstart.resize(n);
    e->h_plen.resize(n);
    e->h_soff.assign(n + 1, 0u);
  }
  e->h_long.clear();
  e->h_huge.clear();
  e->h_huge_off.clear();
  e->h_cta.clear();
  e->h_vlong.clear();
  e->max_block_np2 = 0;
  e->max_block_len = 0;
  e->n_mid_rows = 0;
  e->h_pospref.resize(n + 1);
  e->h_pospref[0] = 0;
  e->n_pos_unsampled = 0;
  e->max_plen = 0;
  const uint64_t every = e->cfg.sample_every > 1 ? e->cfg.sample_every : 1;
  // One pass over the rows (pair order), in parallel slabs: k-mer positions per row (prefix in a second
  // step), the longest row, and the rows the sorting kernels of the table / bucket builds take one by one.
  // (This runs while the residue stream crosses PCIe; at 8 ranks per host it was the largest part of a step.)
  struct Slab {
    unsigned long long pos = 0, pos_all = 0;
    uint32_t max_plen = 0, max_np2 = 0, max_len = 0, n_mid = 0, bad = 0;
    std::vector<uint32_t> lng, cta, vlong, huge;
  };
  const unsigned n_slabs = n >= (1u << 16) ? 4u : 1u;
  std::vector<Slab> slabs(n_slabs);
  auto slab_rows = [&](unsigned t) { return std::pair<uint64_t, uint64_t>{n * t / n_slabs, n * (t + 1) / n_slabs}; };
  auto pass1 = [&](unsigned t) {
    Slab sl;  // (a local: the slabs' counters would share cache lines between the threads)
    const auto [lo, hi] = slab_rows(t);
    for (uint64_t r = lo; r < hi; ++r) {
      const uint32_t p = cross ? e->h_orig[r] : (uint32_t)r;
      if (off[p + 1] < off[p]) {
        sl.bad = 1;
        break;
      }
      const uint64_t len = off[p + 1] - off[p];
      if (cross) {
        e->h_pstart[r] = (uint32_t)off[p];
        e->h_plen[r] = (uint32_t)len;
      }
      sl.max_plen = std::max(sl.max_plen, (uint32_t)len);
      uint32_t kept = 0;
      if (len >= (uint64_t)k) {
        const uint32_t npos = (uint32_t)(len - k + 1);
        kept = (uint32_t)(npos / every);
        sl.pos_all += npos;
        if (npos > kHashMaxPos) {
          if (npos > kBlockMaxPos) {
            sl.huge.push_back((uint32_t)r);
          } else if (npos > kWarpMaxPos) {
            sl.lng.push_back((uint32_t)r);
            sl.max_np2 = std::max(sl.max_np2, next_pow2_u32(npos));
            sl.max_len = std::max(sl.max_len, (uint32_t)len);
            (npos > kCtaHashMaxPos ? sl.vlong : sl.cta).push_back((uint32_t)r);
          } else {
            ++sl.n_mid;
            sl.cta.push_back((uint32_t)r);
          }
        }
      }
      e->h_pospref[r + 1] = kept;  // (per row for now)
      sl.pos += kept;
    }
    slabs[t] = std::move(sl);
  };
  {
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_slabs; ++t) pool.emplace_back(pass1, t);
    pass1(0);
    for (auto& th : pool) th.join();
  }
  for (const Slab& sl : slabs)
    if (sl.bad) return fail(e, KC_EINVAL, "offsets must be non-decreasing");
  unsigned long long huge_total = 0;
  {
    std::vector<unsigned long long> base(n_slabs + 1, 0);
    for (unsigned t = 0; t < n_slabs; ++t) base[t + 1] = base[t] + slabs[t].pos;
    auto pass2 = [&](unsigned t) {
      const auto [lo, hi] = slab_rows(t);
      unsigned long long acc = base[t];
      for (uint64_t r = lo; r < hi; ++r) {
        acc += e->h_pospref[r + 1];
        e->h_pospref[r + 1] = acc;
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_slabs; ++t) pool.emplace_back(pass2, t);
    pass2(0);
    for (auto& th : pool) th.join();
    for (unsigned t = 0; t < n_slabs; ++t) {
      const Slab& sl = slabs[t];
      e->n_pos_unsampled += sl.pos_all;
      e->max_plen = std::max(e->max_plen, sl.max_plen);
      e->max_block_np2 = std::max(e->max_block_np2, sl.max_np2);
      e->max_block_len = std::max(e->max_block_len, sl.max_len);
      e->n_mid_rows += sl.n_mid;
      e->h_long.insert(e->h_long.end(), sl.lng.begin(), sl.lng.end());
      e->h_cta.insert(e->h_cta.end(), sl.cta.begin(), sl.cta.end());
      e->h_vlong.insert(e->h_vlong.end(), sl.vlong.begin(), sl.vlong.end());
      for (uint32_t r : sl.huge) {
        const uint32_t p = cross ? e->h_orig[r] : r;
        const uint32_t npos = (uint32_t)(off[p + 1] - off[p] - k + 1);
        e->h_huge.push_back(r);
        e->h_huge_off.push_back(huge_total);
        huge_total += next_pow2_u32(npos);
      }
    }
    if (cross)
      for (uint64_t r = 0; r < n; ++r) e->h_soff[r + 1] = e->h_soff[r] + e->h_plen[r];
  }
  e->huge_total = huge_total;
  return KC_OK;
}

static cudaError_t stage_up(kc_engine* e, DBuf& b, const void* src, size_t bytes) {
  cudaError_t rc = b.ensure(std::max<size_t>(bytes, 16));
  if (rc != cudaSuccess) return rc;
  if (bytes == 0) return cudaSuccess;
  return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, e->stream);
}

// device part 1: the row layout in the pair order (input order: derived on the device from d_off alone)
int stage_layout_device(kc_engine* e) {
  const uint64_t n = e->n;
  const bool cross = e->cfg.cross_class_only != 0;
  auto up = [&](DBuf& b, const void* src, size_t bytes) { return stage_up(e, b, src, bytes); };
  if (!cross) {  // d_off is already on its way on the same stream
    KC_CUDA(e, e->d_pstart.ensure(std::max<size_t>(n * 4, 16)));
    KC_CUDA(e, e->d_plen.ensure(std::max<size_t>(n * 4, 16)));
    KC_CUDA(e, e->d_orig.ensure(std::max<size_t>(n * 4, 16)));
    KC_CUDA(e, e->d_rank.ensure(std::max<size_t>(n * 4, 16)));
    KC_CUDA(e, e->d_soff.ensure((n + 1) * 4 + 16));
    if (n)
      KC_LAUNCH(e, layout_identity_kernel, (uint32_t)((n + 255) / 256), 256, 0, e->d_off.as<unsigned long long>(),
                (uint32_t)n, e->d_pstart.as<uint32_t>(), e->d_plen.as<uint32_t>(), e->d_orig.as<uint32_t>(),
                e->d_rank.as<uint32_t>(), e->d_soff.as<uint32_t>());
  } else {
    // the residue stream in the pair (class-major) order for the streaming index build
    KC_CUDA(e, up(e->d_soff, e->h_soff.data(), (n + 1) * 4));
    const size_t padded = padded_res_bytes(e->R);
    KC_CUDA(e, e->d_sres_own.ensure(padded));
    KC_CUDA(e, cudaMemsetAsync(e->d_sres_own.as<uint8_t>() + e->R, 0, padded - e->R, e->stream));
    KC_CUDA(e, up(e->d_pstart, e->h_pstart.data(), n * 4));
    KC_CUDA(e, up(e->d_plen, e->h_plen.data(), n * 4));
    KC_CUDA(e, up(e->d_orig, e->h_orig.data(), n * 4));
    KC_CUDA(e, up(e->d_rank, e->h_rank.data(), n * 4));
    KC_CUDA(e, up(e->d_first_after, e->h_first_after.data(), n * 4));
    if (n && e->R)
      KC_LAUNCH(e, sx_permute_residues_kernel, (uint32_t)std::min<uint64_t>((n + 7) / 8, (uint64_t)e->num_sm * 16), 256, 0,
                e->d_res.as<uint8_t>(), e->d_pstart.as<uint32_t>(), e->d_soff.as<uint32_t>(), (uint32_t)n,
                e->d_sres_own.as<uint8_t>());
  }
  return KC_OK;
}

// device part 2: the row lists of the sorting extract kernels (table / bucket builds, kc_extract_kmers)
int stage_lists_device(kc_engine* e) {
  auto up = [&](DBuf& b, const void* src, size_t bytes) { return stage_up(e, b, src, bytes); };
  KC_CUDA(e, up(e->d_long, e->h_long.data(), e->h_long.size() * 4));
  KC_CUDA(e, up(e->d_cta, e->h_cta.data(), e->h_cta.size() * 4));
  KC_CUDA(e, up(e->d_vlong, e->h_vlong.data(), e->h_vlong.size() * 4));
  KC_CUDA(e, up(e->d_huge, e->h_huge.data(), e->h_huge.size() * 4));
  KC_CUDA(e, up(e->d_huge_off, e->h_huge_off.data(), e->h_huge_off.size() * 8));
  if (e->huge_total) KC_CUDA(e, e->d_huge_scratch.ensure(e->huge_total * 4));
  return KC_OK;
}

// everything at once (the callers that have h_off / h_cls in place)
int stage_layout(kc_engine* e) {
  int rc = stage_layout_host(e, e->h_off.data());
  if (rc == KC_OK) rc = stage_layout_device(e);
  if (rc == KC_OK) rc = stage_lists_device(e);
  if (rc != KC_OK) return rc;
  e->have_proteins = true;
  e->have_index = e->have_pairs = false;
  return KC_OK;
}

// the host's own copy of the caller's offsets / classes, in parallel slabs
static void copy_host_arrays(kc_engine* e, const uint64_t* offsets, const uint32_t* class_id) {
  const uint64_t n = e->n;
  e->h_off.resize(n + 1);
  e->h_cls.resize(n);
  const unsigned n_slabs = n >= (1u << 16) ? 4u : 1u;
  auto slab = [&](unsigned t) {
    const uint64_t lo = n * t / n_slabs, hi = n * (t + 1) / n_slabs;
    std::memcpy(e->h_off.data() + lo, offsets + lo, (hi - lo + (t + 1 == n_slabs ? 1 : 0)) * 8);
    if (hi > lo) std::memcpy(e->h_cls.data() + lo, class_id + lo, (hi - lo) * 4);
  };
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < n_slabs; ++t) pool.emplace_back(slab, t);
  slab(0);
  for (auto& th : pool) th.join();
}

// completes a deferred host staging (see kc_engine::pend_off); a no-op otherwise
int ensure_staged(kc_engine* e) {
  if (!e->pend_off) return KC_OK;
  copy_host_arrays(e, e->pend_off, e->pend_cls);
  e->pend_off = nullptr;
  e->pend_cls = nullptr;
  return KC_OK;
}


template <int K>
int run_extract_census(kc_engine* e, BucketScatter scatter = BucketScatter{nullptr, nullptr, 0u, 0u, nullptr, 0u, RowOwner{nullptr, 0u}}) {
  const uint32_t n = (uint32_t)e->n;
  DeviceScalars* ds = e->ds;
  const uint8_t* res = e->d_res.as<uint8_t>();
  uint32_t* pk = e->d_pk.as<uint32_t>();
  uint32_t* ndist = e->d_ndist.as<uint32_t>();
  uint32_t* ksplit = e->d_ksplit.as<uint32_t>();
  // partitioned build: short proteins are deduplicated by hashing (no pk), the sorting warp kernel
  // only takes the proteins of more than kHashMaxPos positions
  uint32_t min_pos = 0;
  if (n && scatter.rec) {
    min_pos = kHashMaxPos;
    // chunk by chunk while the upload of the residue stream is still in flight, else one launch
    const int chunks = e->n_chunks ? e->n_chunks : 1;
    for (int c = 0; c < chunks; ++c) {
      const uint32_t r0 = e->n_chunks ? e->chunk_row[c] : 0u, r1 = e->n_chunks ? e->chunk_row[c + 1] : n;
      if (e->n_chunks) KC_CUDA(e, cudaStreamWaitEvent(e->stream, e->chunk_ev[c], 0));
      if (r1 <= r0) continue;
      const uint32_t grid = blocks_for(r1 - r0, kXsWarps, e->num_sm * 5);
      KC_LAUNCH(e, extract_scatter_warp_kernel<K>, grid, kXsWarps * 32, 0, res, e->d_pstart.as<uint32_t>(),
                e->d_plen.as<uint32_t>(), r0, r1, ndist, e->cfg.sample_every, e->cfg.sample_seed,
                e->d_orig.as<uint32_t>(), &ds->n_incid, scatter);
    }
    e->n_chunks = 0;
  } else {
    int rcw = wait_upload(e);
    if (rcw) return rcw;
  }
  if (scatter.rec && !e->h_cta.empty()) {  // 717 .. 2800 positions: one CTA per protein, shared-memory hash set
    KC_LAUNCH(e, extract_scatter_cta_kernel<K>, (uint32_t)std::min<size_t>(e->h_cta.size(), (size_t)e->num_sm * 16),
              kXcThreads, 0, res, e->d_pstart.as<uint32_t>(), e->d_plen.as<uint32_t>(), e->d_cta.as<uint32_t>(),
              (uint32_t)e->h_cta.size(), ndist, e->cfg.sample_every, e->cfg.sample_seed, e->d_orig.as<uint32_t>(),
              &ds->n_incid, scatter);
  }
  if (n && !scatter.rec) {
    const uint32_t grid = blocks_for(n, kExtractWarps, e->num_sm * 5);
    KC_LAUNCH(e, extract_dedup_warp_kernel<K>, grid, kExtractWarps * 32, 0, res, e->d_pstart.as<uint32_t>(),
              e->d_plen.as<uint32_t>(), n, pk, ndist, e->slice_shift, e->n_slices, ksplit, e->cfg.sample_every,
              e->cfg.sample_seed, e->d_orig.as<uint32_t>(), &ds->n_incid, scatter, min_pos);
  }
  // sorted in shared memory: every protein of more than 1 024 positions, or (partitioned build) more than 2 800
  const std::vector<uint32_t>& blk = scatter.rec ? e->h_vlong : e->h_long;
  if (!blk.empty()) {
    const size_t smem = (size_t)e->max_block_np2 * 4 + e->max_block_len + 16;
    KC_CUDA(e, cudaFuncSetAttribute(extract_dedup_block_kernel<K, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KC_LAUNCH(e, (extract_dedup_block_kernel<K, false>), (uint32_t)blk.size(), 512, smem, res,
              e->d_pstart.as<uint32_t>(), e->d_plen.as<uint32_t>(), (scatter.rec ? e->d_vlong : e->d_long).as<uint32_t>(),
              nullptr, nullptr,
              pk, ndist, n, e->slice_shift, e->n_slices, ksplit, e->cfg.sample_every, e->cfg.sample_seed,
              e->d_orig.as<uint32_t>(), &ds->n_incid, scatter);
  }
  if (!e->h_huge.empty()) {
    KC_LAUNCH(e, (extract_dedup_block_kernel<K, true>), (uint32_t)e->h_huge.size(), 512, 0, res,
              e->d_pstart.as<uint32_t>(), e->d_plen.as<uint32_t>(), e->d_huge.as<uint32_t>(),
              e->d_huge_off.as<unsigned long long>(), e->d_huge_scratch.as<uint32_t>(), pk, ndist, n,
              e->slice_shift, e->n_slices, ksplit, e->cfg.sample_every, e->cfg.sample_seed,
              e->d_orig.as<uint32_t>(), &ds->n_incid, scatter);
  }
  return KC_OK;
}

template <int K>
int run_positions(kc_engine* e, uint32_t* out) {
  if (e->cfg.sample_every > 1) {
    if (e->n)
      KC_LAUNCH(e, kmers_sampled_kernel<K>, blocks_for(e->n, 8, e->num_sm * 8), 256, 0, e->d_res.as<uint8_t>(),
                e->d_off.as<unsigned long long>(), e->d_kpos.as<unsigned long long>(), (uint32_t)e->n,
                e->cfg.sample_every, e->cfg.sample_seed, out);
    return KC_OK;
  }
  const uint32_t grid = (uint32_t)((e->R + kTileRes - 1) / kTileRes);
  if (grid)
    KC_LAUNCH(e, kmers_per_position_kernel<K>, grid, 256, 0, e->d_res.as<uint8_t>(), e->R,
              e->d_off.as<unsigned long long>(), e->d_kpos.as<unsigned long long>(), (uint32_t)e->n, out);
  return KC_OK;
}

template <int LOG_H, int GROUP_WARPS, int CTA_WARPS>
int launch_hash(kc_engine* e, uint8_t bin, const EdgeSink& sink) {
  constexpr size_t smem = (size_t)(CTA_WARPS / GROUP_WARPS) * 2 * (1u << LOG_H) * 4;
  auto kern = pairs_hash_kernel<LOG_H, GROUP_WARPS, CTA_WARPS>;
  KC_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CTA_WARPS * 32, smem));
  if (per_sm < 1) per_sm = 1;
  const uint32_t grid = (uint32_t)(e->num_sm * per_sm);
  KC_LAUNCH(e, kern, grid, CTA_WARPS * 32, smem, e->pair_rowptr(), e->pair_rowlen(),
            e->d_suf.as<uint2>(), e->d_col.as<uint32_t>(), e->d_rowbin.as<uint8_t>(), bin, (uint32_t)e->n,
            &e->ds->row_cursor[bin], e->ds->bin_counts, sink, &e->ds->pc);
  return KC_OK;
}

template <int LOG_H, int GROUP_WARPS, int CTA_WARPS>
int launch_packed(kc_engine* e, uint8_t bin, uint32_t count_bits, const EdgeSink& sink) {
  constexpr size_t smem =
      ((size_t)(CTA_WARPS / GROUP_WARPS) * (1u << LOG_H) + (size_t)CTA_WARPS * (kIdxPerWarp + kStageWords)) * 4;
  auto kern = pairs_packed_kernel<LOG_H, GROUP_WARPS, CTA_WARPS>;
  KC_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CTA_WARPS * 32, smem));
  if (per_sm < 1) per_sm = 1;
  const uint32_t grid = (uint32_t)(e->num_sm * per_sm);
  KC_LAUNCH(e, kern, grid, CTA_WARPS * 32, smem, e->pair_rowptr(), e->pair_rowlen(),
            e->d_suf.as<uint2>(), e->d_col.as<uint32_t>(), e->d_rowbin.as<uint8_t>(), bin, (uint32_t)e->n,
            count_bits, &e->ds->row_cursor[bin], e->ds->bin_counts, sink, &e->ds->pc);
  return KC_OK;
}

__global__ void shard_bounds_kernel(const unsigned long long* __restrict__ prefix, uint32_t n, uint32_t shard,
                                    uint32_t n_shards, uint32_t* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned long long total = prefix[n];
  for (int side = 0; side < 2; ++side) {
    const uint32_t s = shard + side;
    uint32_t row;
    if (s == 0) row = 0;
    else if (s >= n_shards) row = n;
    else {
      // first row whose prefix reaches total * s / n_shards
      const unsigned long long target = total / n_shards * s + (total % n_shards) * s / n_shards;
      uint32_t lo = 0, hi = n;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (prefix[mid] < target) lo = mid + 1; else hi = mid;
      }
      row = lo;
    }
    out[side] = row;
  }
}

struct WorkIn {
  const uint32_t* rowwork;
  const uint32_t* rowlen;
  __device__ unsigned long long operator()(uint64_t i) const {
    return (unsigned long long)rowwork[i] + 2ull * rowlen[i] + (rowwork[i] ? 64ull : 0ull);
  }
};
struct U64ExclOut {
  unsigned long long* p;
  __device__ void operator()(uint64_t i, unsigned long long excl, unsigned long long v) const {
    p[i] = excl;
    (void)v;
  }
};
struct U64ExclOutWithTail {
  unsigned long long* p;
  uint64_t n;
  __device__ void operator()(uint64_t i, unsigned long long excl, unsigned long long v) const {
    p[i] = excl;
    if (i + 1 == n) p[n] = excl + v;
  }
};
struct ColptrOut {
  uint32_t* p;
  __device__ void operator()(uint64_t i, unsigned long long excl, unsigned long long) const {
    p[i] = (uint32_t)excl;
  }
};

}  // namespace


// ---- the partitioned index build (bucket.cuh) -----------------------------------------------
static int build_index_bucketed(kc_engine* e, uint32_t shard, uint32_t n_shards, kc_index_stats* stats,
                                bool* overflow) {
  const uint32_t n = (uint32_t)e->n;
  const uint64_t R = e->R;
  DeviceScalars* ds = e->ds;
  *overflow = false;
  const unsigned long long n_positions = e->h_pospref[n];
  const uint64_t E = std::max<unsigned long long>(n_positions, 1);  // upper bound on the incidences
  const uint32_t n_bins = (n + kBinRows - 1) >> kBinRowsLog;                // entry bins
  // Sharded build: the pair order is cut into 2 * n_shards row blocks of equal k-mer positions (at
  // bin borders); rank g owns blocks g and 2 * n_shards - 1 - g.  The work of a row falls with its
  // position in the pair order (it is scored against the rows after it), so the zig-zag gives every
  // rank one early and one late block.
  unsigned long long own_positions = n_positions;
  uint32_t n_own_rows = n;
  std::vector<uint32_t> bounds;
  if (n_shards > 1) {
    const uint32_t n_blocks = 2 * n_shards;
    bounds.assign(n_blocks + 1, n);
    bounds[0] = 0;
    for (uint32_t b = 1; b < n_blocks; ++b) {
      const unsigned long long target = n_positions / n_blocks * b + (n_positions % n_blocks) * b / n_blocks;
      const uint32_t r = (uint32_t)(std::lower_bound(e->h_pospref.begin(), e->h_pospref.end(), target) -
                                    e->h_pospref.begin());  // first r with positions(rows < r) >= target
      bounds[b] = std::max(bounds[b - 1], std::min<uint32_t>(n, (r + kBinRows - 1) & ~(kBinRows - 1u)));
    }
    e->h_binowner.assign(n_bins, 0);
    own_positions = 0;
    n_own_rows = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
      const uint32_t who = b < n_shards ? b : n_blocks - 1 - b;
      for (uint32_t bin = bounds[b] >> kBinRowsLog; bin < (bounds[b + 1] + kBinRows - 1) >> kBinRowsLog; ++bin)
        e->h_binowner[bin] = (uint8_t)who;
      if (who != shard) continue;
      n_own_rows += bounds[b + 1] - bounds[b];
      own_positions += e->h_pospref[bounds[b + 1]] - e->h_pospref[bounds[b]];
    }
    KC_CUDA(e, e->d_binowner.ensure((size_t)n_bins + 64));
    KC_CUDA(e, cudaMemcpyAsync(e->d_binowner.p, e->h_binowner.data(), n_bins, cudaMemcpyHostToDevice, e->stream));
  }
  const RowOwner owner{n_shards > 1 ? e->d_binowner.as<uint8_t>() : nullptr, shard};
  // The records this build keeps: all of its own rows' plus what passes the filter.  Bucket count
  // from what the last build of this shape kept, else from a guess (a bucket overflow retries with
  // the whole-set count).
  const uint32_t cap = e->try_cap ? e->try_cap
                                  : (e->cap_hint_n == e->n && e->cap_hint_R == e->R && e->cap_hint ? e->cap_hint : 4096u);
  const uint32_t kBkTargetFill = bk_target_fill(cap);
  const size_t kBkSmemBytes = bk_smem_bytes(cap);
  const uint32_t NB_full = (uint32_t)((E + kBkTargetFill - 1) / kBkTargetFill);
  uint32_t NB = NB_full;
  if (n_shards > 1) {
    const unsigned long long guess = e->kept_hint && e->hint_shards == n_shards
                                         ? e->kept_hint + e->kept_hint / 4
                                         : std::min<unsigned long long>(E, 2 * own_positions + E / 16);
    NB = (uint32_t)std::min<unsigned long long>(NB_full, (guess + kBkTargetFill - 1) / kBkTargetFill + 1);
  }
  if (e->retry_full_buckets) NB = NB_full;
  uint32_t filter_bits = 0;
  if (n_shards > 1) {
    filter_bits = 1u << 20;
    while (filter_bits < (1u << 30) && (unsigned long long)filter_bits < 8ull * std::max<unsigned long long>(own_positions, 1))
      filter_bits <<= 1;
    KC_CUDA(e, e->d_filter.ensure((size_t)filter_bits / 8 + 64));
  }
  KC_CUDA(e, e->d_pk.ensure((R + 64) * 4));
  KC_CUDA(e, e->d_ndist.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rowlen.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rec.ensure((uint64_t)NB * cap * 8));
  KC_CUDA(e, e->d_bucket_cnt.ensure(((uint64_t)NB_full + 2) * 4));
  KC_CUDA(e, e->d_entries.ensure((E + E / 2 + 64) * 16));  // entries + run records (bin_region)
  KC_CUDA(e, e->d_runs.ensure((E / 2 + 64) * 16));
  KC_CUDA(e, e->d_run_cnt.ensure(((uint64_t)n_bins + 2) * 4));
  KC_CUDA(e, e->d_rowlen_p.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_bucket_cnt.ensure(((uint64_t)NB + 2) * 4));                // bucket cursors
  KC_CUDA(e, e->d_rowcap.ensure(((uint64_t)n + 2) * 4));                    // rowcap prefix
  KC_CUDA(e, e->d_bin_cnt.ensure(((uint64_t)n_bins + 2) * 4));                // bin cursors
  KC_CUDA(e, e->d_ids.ensure((E + 64) * 4));
  KC_CUDA(e, e->d_col.ensure((E + 64) * 4));
  KC_CUDA(e, e->d_suf.ensure((E + 64) * 8));
  if (e->cfg.want_blosum) KC_CUDA(e, e->d_sufss.ensure(E + 64));
  KC_CUDA(e, e->d_vocab_h.ensure((E / 2 + 2) * 4));
  KC_CUDA(e, e->d_freq_h.ensure((E / 2 + 2) * 4));
  KC_CUDA(e, e->d_self_h.ensure(E / 2 + 16));
  KC_CUDA(e, e->d_rowwork.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rowwork64.ensure(((uint64_t)n + 1) * 8));
  KC_CUDA(e, e->d_rowinl.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rowmaxlen.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_workprefix.ensure(((uint64_t)n + 2) * 8));
  int rc = ensure_scan(e, (uint64_t)n + 1);
  if (rc) return rc;
  e->have_plist = false;
  e->n_slices = 0;
  uint32_t* bucket_cnt = e->d_bucket_cnt.as<uint32_t>();
  uint32_t* rowcap = e->d_rowcap.as<uint32_t>();
  uint32_t* bin_cnt = e->d_bin_cnt.as<uint32_t>();
  uint2* rec = e->d_rec.as<uint2>();
  uint4* ent = e->d_entries.as<uint4>();

  mark(e, EV_I0);
  KC_CUDA(e, cudaMemsetAsync(ds, 0, sizeof(DeviceScalars), e->stream));
  KC_CUDA(e, cudaMemsetAsync(bucket_cnt, 0, ((uint64_t)NB + 1) * 4, e->stream));
  KC_CUDA(e, cudaMemsetAsync(bin_cnt, 0, ((uint64_t)n_bins + 1) * 4, e->stream));
  // K1/K2: extract + per-protein dedup; every (distinct k-mer, row) is appended to its bucket
  mark(e, EV_IC0);
  if (n_shards > 1) {  // filter of the k-mers this rank's rows hold
    KC_CUDA(e, cudaMemsetAsync(e->d_filter.p, 0, (size_t)filter_bits / 8, e->stream));
    if (n_own_rows) {  // chunk by chunk while the upload is in flight (the extract pass needs the whole filter)
      const int chunks = e->n_chunks ? e->n_chunks : 1;
      for (int c = 0; c < chunks; ++c) {
        const uint32_t r0 = e->n_chunks ? e->chunk_row[c] : 0u, r1 = e->n_chunks ? e->chunk_row[c + 1] : n;
        if (e->n_chunks) KC_CUDA(e, cudaStreamWaitEvent(e->stream, e->chunk_ev[c], 0));
        if (r1 <= r0) continue;
        const uint32_t fgrid = blocks_for(r1 - r0, 8, e->num_sm * 8);
        if (e->cfg.k == 5)
          KC_LAUNCH(e, kmer_filter_build_kernel<5>, fgrid, 256, 0, e->d_res.as<uint8_t>(), e->d_pstart.as<uint32_t>(),
                    e->d_plen.as<uint32_t>(), r0, r1, owner, e->cfg.sample_every, e->cfg.sample_seed,
                    e->d_orig.as<uint32_t>(), e->d_filter.as<uint32_t>(), filter_bits - 1u);
        else
          KC_LAUNCH(e, kmer_filter_build_kernel<7>, fgrid, 256, 0, e->d_res.as<uint8_t>(), e->d_pstart.as<uint32_t>(),
                    e->d_plen.as<uint32_t>(), r0, r1, owner, e->cfg.sample_every, e->cfg.sample_seed,
                    e->d_orig.as<uint32_t>(), e->d_filter.as<uint32_t>(), filter_bits - 1u);
      }
    }
    rc = wait_upload(e);
    if (rc) return rc;
  }
  {
    DBuf none;
    std::swap(none, e->d_ksplit);  // the rows are not sliced: run_extract_census passes ksplit = null
    const BucketScatter scatter{rec, bucket_cnt, NB, cap, n_shards > 1 ? e->d_filter.as<uint32_t>() : nullptr,
                                filter_bits - 1u, owner};
    rc = e->cfg.k == 5 ? run_extract_census<5>(e, scatter) : run_extract_census<7>(e, scatter);
    std::swap(none, e->d_ksplit);
    if (rc) return rc;
  }
  mark(e, EV_IC1);
  // entry capacity of every row block: the rows' distinct k-mers
  e->launches += exclusive_scan(U32In{e->d_ndist.as<uint32_t>()}, RowCapOut{rowcap, n}, n, e->scan, e->stream);
  // buckets: census, ids, postings, entries
  {
    const uint32_t* fa = e->cfg.cross_class_only ? e->d_first_after.as<uint32_t>() : nullptr;
#define KC_BUCKETS(CROSS, CAP)                                                                                  \
  do {                                                                                                          \
    KC_CUDA(e, cudaFuncSetAttribute((bucket_build_kernel<CROSS, CAP>), cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)kBkSmemBytes));                                                        \
    int per_sm = 1;                                                                                             \
    KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (bucket_build_kernel<CROSS, CAP>), CAP / 8, \
                                                             kBkSmemBytes));                                    \
    const uint32_t grid = std::min<uint32_t>(NB, (uint32_t)(e->num_sm * std::max(per_sm, 1)));                   \
    KC_LAUNCH(e, (bucket_build_kernel<CROSS, CAP>), grid, CAP / 8, kBkSmemBytes, rec, bucket_cnt, NB, fa, e->cfg.k, \
              e->d_col.as<uint32_t>(), ent, rowcap, bin_cnt, e->d_vocab_h.as<uint32_t>(),                       \
              e->d_freq_h.as<uint32_t>(), e->d_self_h.as<uint8_t>(), owner, &ds->bg);                           \
  } while (0)
    if (cap == 4096u) {
      if (fa) KC_BUCKETS(true, 4096u); else KC_BUCKETS(false, 4096u);
    } else {
      if (fa) KC_BUCKETS(true, 8192u); else KC_BUCKETS(false, 8192u);
    }
#undef KC_BUCKETS
  }
  // entry bins -> rows (laid out by capacity: row r starts at rowcap[r])
  KC_CUDA(e, cudaFuncSetAttribute(rows_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinSmemBytes));
  const uint32_t bin_lo = 0, bin_hi = n_bins;  // foreign rows keep (short) id lists too: see bucket_build_kernel
  if (bin_hi > bin_lo)
    KC_LAUNCH(e, rows_finalize_kernel, std::min<uint32_t>(bin_hi - bin_lo, (uint32_t)e->num_sm * 3), kFinThreads,
              kFinSmemBytes, ent, rowcap, bin_cnt, n, bin_lo, bin_hi, e->d_runs.as<uint4>(),
              e->d_run_cnt.as<uint32_t>(), e->d_rowlen.as<uint32_t>(), e->d_rowlen_p.as<uint32_t>(), e->d_ids.as<uint32_t>(), e->d_suf.as<uint2>(),
            e->cfg.want_blosum ? e->d_sufss.as<uint8_t>() : nullptr, e->d_rowwork64.as<unsigned long long>(),
            e->d_rowwork.as<uint32_t>(), e->d_rowinl.as<uint32_t>(), e->d_rowmaxlen.as<uint32_t>());
  if (n_shards > 1)  // distinct k-mers of the own rows (ds->multi_total is free in this build: bg holds the totals)
    KC_LAUNCH(e, own_incidences_kernel, blocks_for(n, 256, e->num_sm * 4), 256, 0, e->d_ndist.as<uint32_t>(), n, owner,
              &ds->multi_total);
  else
    e->launches += exclusive_scan(WorkIn{e->d_rowwork.as<uint32_t>(), e->d_rowlen.as<uint32_t>()},
                                  U64ExclOutWithTail{e->d_workprefix.as<unsigned long long>(), n}, n, e->scan,
                                  e->stream);
  mark(e, EV_I1);
  DeviceScalars hs{};
  KC_CUDA(e, cudaMemcpyAsync(&hs, ds, sizeof(hs), cudaMemcpyDeviceToHost, e->stream));
  ensure_staged(e);  // the engine's own copies of offsets / classes, while the kernels above run
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  KC_CUDA(e, cudaGetLastError());
  if (hs.bg.overflow) {
    if (NB < NB_full && !e->retry_full_buckets) {  // the guess of what a sharded build keeps was too small
      e->retry_full_buckets = true;
      int rc2 = build_index_bucketed(e, shard, n_shards, stats, overflow);
      e->retry_full_buckets = false;
      return rc2;
    }
    if (cap == 4096u) {  // a k-mer (or a clump) with thousands of holders: twice the slot size
      e->try_cap = 8192u;
      int rc2 = build_index_bucketed(e, shard, n_shards, stats, overflow);
      e->try_cap = 0;
      return rc2;
    }
    *overflow = true;
    return KC_OK;
  }
  e->cap_hint = cap;
  e->cap_hint_n = e->n;
  e->cap_hint_R = e->R;
  e->kept_hint = hs.bg.n_records;
  e->hint_shards = n_shards;
  // totals over the k-mers this build owns: the whole-set numbers when summed over the shards
  const unsigned long long own_incid = n_shards > 1 ? hs.multi_total : hs.n_incid;  // own_incidences_kernel
  e->istats.n_positions = own_positions;
  e->istats.n_incidences = own_incid;
  e->istats.n_distinct = hs.bg.n_distinct;
  e->istats.n_repeated = hs.bg.n_repeated;
  e->istats.n_singleton = hs.bg.n_distinct - hs.bg.n_repeated;
  e->istats.nnz = hs.bg.nnz;
  e->v_local = hs.bg.id_cursor;
  e->multi_total = hs.bg.multi_total;
  e->work_total = hs.bg.work_total;
  e->ishard = shard;
  e->ishards = n_shards;
  e->block_bounds = bounds;
  e->n_own_rows = n_own_rows;
  if (stats) *stats = e->istats;
  e->bucketed = true;
  e->canonical_ready = false;
  e->have_index = true;
  return KC_OK;
}

// ---- the streaming partitioned index build (stream_index.cuh): the default -------------------
// pass_end_tile: null / n_pass <= 1 = one level-1 pass over the whole stream; else pass u ends before tile
// pass_end_tile[u] (the upload chunks of an asynchronous kc_set_proteins)
static SxPlan sx_make_plan(uint64_t E, uint64_t R, int num_sm, int n_pass = 1, const uint32_t* pass_end_tile = nullptr) {
  SxPlan p{};
  // ~330 records per bucket: one warp sorts a bucket in its own slice of shared memory (512 records).
  // b >= 12: the CTA kernel's sort key (32 - b bits) must fit a u32 beside a 12-bit record index;
  // b <= 20: 1 024 digits per level.
  uint32_t b = 12;
  while (b < 20 && (E >> b) > 330) ++b;
  p.b1 = (b + 1) / 2;
  p.b2 = b - p.b1;
  p.r = 32 - b;
  const uint32_t tiles = (uint32_t)std::max<uint64_t>(1, (R + kSxTile - 1) / kSxTile);
  p.n_pass = 1;
  p.pass_tile[0] = 0;
  p.pass_tile[1] = tiles;
  if (n_pass > 1 && pass_end_tile) {
    p.n_pass = (uint32_t)std::min(n_pass, kSxMaxPass);
    for (uint32_t u = 0; u < p.n_pass; ++u)
      p.pass_tile[u + 1] = u + 1 == p.n_pass ? tiles : std::max(p.pass_tile[u], std::min(tiles, pass_end_tile[u]));
  }
  uint32_t widest = 1;
  for (uint32_t u = 0; u < p.n_pass; ++u) widest = std::max(widest, p.pass_tile[u + 1] - p.pass_tile[u]);
  uint32_t g1 = std::min<uint32_t>(widest, (uint32_t)num_sm * 2u);  // two persistent CTAs per SM
  const uint32_t tpc = (widest + g1 - 1) / g1;
  p.g1 = (widest + tpc - 1) / tpc;
  p.c2 = std::min<uint32_t>(16u, std::max<uint32_t>(1u, ((uint32_t)num_sm * 16u) >> p.b1));
  p.ballots = 1;  // level 1 / 2 (9-10 digit bits): one ballot per bit measured faster than match.any
  return p;
}

// Row blocks of a sharded build: the pair order is cut into 2 * n_shards blocks of equal k-mer positions (at bin
// borders); rank g owns blocks g and 2 * n_shards - 1 - g (a row is scored against the rows after it, so its
// work falls with its position: the zig-zag gives every rank one early and one late block).
static void shard_blocks(const kc_engine* e, uint32_t shard, uint32_t n_shards, std::vector<uint32_t>& bounds,
                         std::vector<uint8_t>& binowner, unsigned long long* own_positions, uint32_t* n_own_rows) {
  const uint32_t n = (uint32_t)e->n;
  const unsigned long long n_positions = e->h_pospref[n];
  const uint32_t n_bins = (n + kBinRows - 1) >> kBinRowsLog;
  const uint32_t n_blocks = 2 * n_shards;
  bounds.assign(n_blocks + 1, n);
  bounds[0] = 0;
  for (uint32_t b = 1; b < n_blocks; ++b) {
    const unsigned long long target = n_positions / n_blocks * b + (n_positions % n_blocks) * b / n_blocks;
    const uint32_t r = (uint32_t)(std::lower_bound(e->h_pospref.begin(), e->h_pospref.end(), target) -
                                  e->h_pospref.begin());  // first r with positions(rows < r) >= target
    bounds[b] = std::max(bounds[b - 1], std::min<uint32_t>(n, (r + kBinRows - 1) & ~(kBinRows - 1u)));
  }
  binowner.assign(n_bins, 0);
  *own_positions = 0;
  *n_own_rows = 0;
  for (uint32_t b = 0; b < n_blocks; ++b) {
    const uint32_t who = b < n_shards ? b : n_blocks - 1 - b;
    for (uint32_t bin = bounds[b] >> kBinRowsLog; bin < (bounds[b + 1] + kBinRows - 1) >> kBinRowsLog; ++bin)
      binowner[bin] = (uint8_t)who;
    if (who != shard) continue;
    *n_own_rows += bounds[b + 1] - bounds[b];
    *own_positions += e->h_pospref[bounds[b + 1]] - e->h_pospref[bounds[b]];
  }
}

static int build_index_stream(kc_engine* e, uint32_t shard, uint32_t n_shards, kc_index_stats* stats) {
  const uint32_t n = (uint32_t)e->n;
  const uint64_t R = e->R;
  DeviceScalars* ds = e->ds;
  if (n_shards > 1)
    if (int rcs = ensure_staged(e)) return rcs;
  const unsigned long long n_positions = e->h_pospref[n];
  const uint64_t E = std::max<unsigned long long>(n_positions, 1);
  const uint32_t n_bins = (n + kBinRows - 1) >> kBinRowsLog;
  // sharded build (owner computes): this rank's row blocks, the filter of its rows' k-mers
  unsigned long long own_positions = n_positions;
  uint32_t n_own_rows = n, filter_bits = 0;
  std::vector<uint32_t> bounds;
  if (n_shards > 1) {
    shard_blocks(e, shard, n_shards, bounds, e->h_binowner, &own_positions, &n_own_rows);
    KC_CUDA(e, e->d_binowner.ensure((size_t)n_bins + 64));
    KC_CUDA(e, cudaMemcpyAsync(e->d_binowner.p, e->h_binowner.data(), n_bins, cudaMemcpyHostToDevice, e->stream));
    // 16 bits per own position, at most 2^29 bits = 64 MB: the probes (one per foreign position) must hit L2
    filter_bits = 1u << 20;
    while (filter_bits < (1u << 29) && (unsigned long long)filter_bits < 16ull * std::max<unsigned long long>(own_positions, 1))
      filter_bits <<= 1;
    KC_CUDA(e, e->d_filter.ensure((size_t)filter_bits / 8 + 64));
    KC_CUDA(e, e->d_keepmask.ensure(((R + kSxTile - 1) / kSxTile + 1) * kL1Threads * 2));
  }
  const RowOwner owner{n_shards > 1 ? e->d_binowner.as<uint8_t>() : nullptr, shard};
  // An upload still in flight (kc_set_proteins from host buffers, pair order = input order): one level-1 pass per
  // upload chunk, each waiting for its own chunk only; the level-1 sort of chunk u runs while chunk u + 1
  // crosses PCIe.  A tile also reads the first bytes of the next one (the k-mer halo): pass u stops one tile
  // short of its chunk's end, the last pass takes the rest.
  static_assert(kc_engine::kUploadChunks <= kSxMaxPass, "one level-1 pass per upload chunk");
  int n_pass = 1;
  uint32_t pass_end[kc_engine::kUploadChunks] = {};
  if (e->n_chunks > 1 && n_shards <= 1 && !e->cfg.cross_class_only) {
    n_pass = e->n_chunks;
    for (int u = 0; u < n_pass; ++u) {
      const uint64_t t = e->chunk_byte[u + 1] / kSxTile;
      pass_end[u] = (uint32_t)(t ? t - 1 : 0);
    }
  } else {
    KC_CUDA(e, wait_upload(e) == KC_OK ? cudaSuccess : cudaErrorUnknown);
  }
  SxPlan plan = sx_make_plan(E, R, e->num_sm, n_pass, pass_end);
  if (e->cfg.census_merge == 77u) plan.ballots = 0;  // (A/B switch while tuning; see profiles/r2_history.md)
  const uint32_t n_tiles = (uint32_t)((R + kSxTile - 1) / kSxTile);
  const uint32_t D1 = plan.d1(), D2 = plan.d2(), NB = plan.n_buckets();
  const uint64_t n_h1 = (uint64_t)plan.h1_stride() * plan.n_pass, n_h2 = (uint64_t)NB * plan.c2;
  KC_CUDA(e, e->d_rec_a.ensure((E + 64) * 8));
  KC_CUDA(e, e->d_rec_b.ensure((E + 64) * 8));
  KC_CUDA(e, e->d_h1.ensure((n_h1 + 2) * 4));
  KC_CUDA(e, e->d_h2.ensure((n_h2 + 2) * 4));
  KC_CUDA(e, e->d_huge_list.ensure(((uint64_t)NB + 2) * 4));
  KC_CUDA(e, e->d_mid_list.ensure(((uint64_t)NB + 2) * 4));
  KC_CUDA(e, e->d_tile_row.ensure(((uint64_t)n_tiles + 2) * 4));
  KC_CUDA(e, e->d_rowlen.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_entries.ensure((E + E / 2 + 64) * 16));
  KC_CUDA(e, e->d_runs.ensure((E / 2 + 64) * 16));
  KC_CUDA(e, e->d_run_cnt.ensure(((uint64_t)n_bins + 2) * 4));
  KC_CUDA(e, e->d_rowlen_p.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rowcap.ensure(((uint64_t)n + 2) * 4));
  KC_CUDA(e, e->d_bin_cnt.ensure(((uint64_t)n_bins + 2) * 4));
  KC_CUDA(e, e->d_ids.ensure((E + 64) * 4));
  KC_CUDA(e, e->d_col.ensure((E + 64) * 4));
  KC_CUDA(e, e->d_suf.ensure((E + 64) * 8));
  if (e->cfg.want_blosum) KC_CUDA(e, e->d_sufss.ensure(E + 64));
  KC_CUDA(e, e->d_vocab_h.ensure((E / 2 + 2) * 4));
  KC_CUDA(e, e->d_freq_h.ensure((E / 2 + 2) * 4));
  KC_CUDA(e, e->d_self_h.ensure(E / 2 + 16));
  KC_CUDA(e, e->d_rowwork.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rowwork64.ensure(((uint64_t)n + 1) * 8));
  KC_CUDA(e, e->d_rowinl.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rowmaxlen.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_workprefix.ensure(((uint64_t)n + 2) * 8));
  int rc = ensure_scan(e, std::max<uint64_t>(std::max<uint64_t>(n_h1, n_h2), (uint64_t)n) + 1);
  if (rc) return rc;
  e->have_plist = false;
  e->n_slices = 0;
  const uint8_t* res = e->sres();
  const uint32_t* soff = e->d_soff.as<uint32_t>();
  uint32_t* h1 = e->d_h1.as<uint32_t>();
  uint32_t* h2 = e->d_h2.as<uint32_t>();
  uint2* rec_a = e->d_rec_a.as<uint2>();
  uint2* rec_b = e->d_rec_b.as<uint2>();
  uint32_t* rowcap = e->d_rowcap.as<uint32_t>();
  uint32_t* bin_cnt = e->d_bin_cnt.as<uint32_t>();
  uint4* ent = e->d_entries.as<uint4>();
  const bool k5 = e->cfg.k == 5;

  mark(e, EV_I0);
  KC_CUDA(e, cudaMemsetAsync(ds, 0, sizeof(DeviceScalars), e->stream));
  KC_CUDA(e, cudaMemsetAsync(bin_cnt, 0, ((uint64_t)n_bins + 1) * 4, e->stream));
  mark(e, EV_IC0);
  // entry capacity of every row: its k-mer positions
  e->launches += exclusive_scan(SxPosIn{soff, (uint32_t)e->cfg.k}, SxExclOutTail{rowcap, n}, n, e->scan, e->stream);
  KC_LAUNCH(e, sx_tile_rows_kernel, (n_tiles + 256) / 256, 256, 0, soff, n, (uint32_t)R, n_tiles,
            e->d_tile_row.as<uint32_t>());
  const SxKeep keep{n_shards > 1 ? e->d_filter.as<uint32_t>() : nullptr, filter_bits / 32u - 1u, owner,
                    n_shards > 1 ? e->d_keepmask.as<uint16_t>() : nullptr};
  if (n_shards > 1) {  // the filter of the k-mers this rank's rows hold: its two row blocks, tile by tile
    KC_CUDA(e, cudaMemsetAsync(e->d_filter.p, 0, (size_t)filter_bits / 8, e->stream));
    for (int part = 0; part < 2; ++part) {
      const uint32_t blk = part == 0 ? shard : 2 * n_shards - 1 - shard;
      const uint32_t r_lo = bounds[blk], r_hi = bounds[blk + 1];
      if (r_hi <= r_lo) continue;
      const uint64_t b_lo = e->cfg.cross_class_only ? e->h_soff[r_lo] : e->h_off[r_lo];
      const uint64_t b_hi = e->cfg.cross_class_only ? e->h_soff[r_hi] : e->h_off[r_hi];
      if (b_hi <= b_lo) continue;
      const uint32_t t_lo = (uint32_t)(b_lo / kSxTile), t_hi = (uint32_t)((b_hi + kSxTile - 1) / kSxTile);
      const uint32_t fgrid = std::min<uint32_t>(t_hi - t_lo, (uint32_t)e->num_sm * 2u);
      if (k5)
        KC_LAUNCH(e, sx_filter_build_kernel<5>, fgrid, kL1Threads, 0, res, (uint32_t)R, soff, e->d_tile_row.as<uint32_t>(),
                  t_lo, t_hi, keep);
      else
        KC_LAUNCH(e, sx_filter_build_kernel<7>, fgrid, kL1Threads, 0, res, (uint32_t)R, soff, e->d_tile_row.as<uint32_t>(),
                  t_lo, t_hi, keep);
    }
  }
  // level 1: count, scan, scatter (residue stream -> rec_a, partitioned by the top b1 hash bits)
  {
    const uint32_t* trow = e->d_tile_row.as<uint32_t>();
    const size_t smem_c = (size_t)D1 * 4 + kSxTile + 64 + 256;
    const size_t smem_s = sx_scatter_smem(D1, kL1Threads / 32);
    const uint64_t n_hp = (uint64_t)D1 * plan.g1;  // histogram entries of one pass (+ 1: the end of its region)
#define KC_L1(KK, SH)                                                                                              \
  do {                                                                                                            \
    KC_CUDA(e, cudaFuncSetAttribute((sx_l1_count_kernel<KK, SH>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));   \
    KC_CUDA(e, cudaFuncSetAttribute((sx_l1_scatter_kernel<KK, SH>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s)); \
    for (uint32_t u = 0; u < plan.n_pass; ++u) {                                                                  \
      if (n_pass > 1) KC_CUDA(e, cudaStreamWaitEvent(e->stream, e->chunk_ev[u], 0));                              \
      uint32_t* hp = h1 + (size_t)u * plan.h1_stride();                                                           \
      KC_LAUNCH(e, (sx_l1_count_kernel<KK, SH>), plan.g1, kL1Threads, smem_c, res, (uint32_t)R, soff, trow, plan, u, keep, h1); \
      e->launches += exclusive_scan(U32In{hp}, SxExclOutTailBase{hp, n_hp, u ? hp - 1 : nullptr}, n_hp, e->scan, e->stream);    \
      KC_LAUNCH(e, (sx_l1_scatter_kernel<KK, SH>), plan.g1, kL1Threads, smem_s, res, (uint32_t)R, soff, trow, plan, u, keep,    \
                h1, rec_a);                                                                                       \
    }                                                                                                             \
    if (n_pass > 1) e->n_chunks = 0;                                                                              \
  } while (0)
    if (k5) {
      if (n_shards > 1) KC_L1(5, true); else KC_L1(5, false);
    } else {
      if (n_shards > 1) KC_L1(7, true); else KC_L1(7, false);
    }
#undef KC_L1
  }
  // level 2: rec_a -> rec_b, every level-1 partition by the next b2 hash bits
  {
    const size_t smem_c = (size_t)D2 * 4;
    const size_t smem_s = sx_scatter_smem(D2, kL2Threads / 32);
    KC_CUDA(e, cudaFuncSetAttribute(sx_l2_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    KC_LAUNCH(e, sx_l2_count_kernel, D1 * plan.c2, kL2Threads, smem_c, rec_a, h1, plan, h2);
    e->launches += exclusive_scan(U32In{h2}, SxExclOutTail{h2, n_h2}, n_h2, e->scan, e->stream);
    KC_LAUNCH(e, sx_l2_scatter_kernel, D1 * plan.c2, kL2Threads, smem_s, rec_a, h1, plan, h2, rec_b);
  }
  mark(e, EV_IC1);
  // buckets: sort, census, ids, postings, entries
  {
    SxBucketArgs A{};
    A.rec = rec_b;
    A.scratch = rec_a;
    A.h2 = h2;
    A.plan = plan;
    A.first_after = e->cfg.cross_class_only ? e->d_first_after.as<uint32_t>() : nullptr;
    A.k = e->cfg.k;
    A.col = e->d_col.as<uint32_t>();
    A.entries = ent;
    A.rowcap_prefix = rowcap;
    A.bin_cursor = bin_cnt;
    A.vocab = e->d_vocab_h.as<uint32_t>();
    A.freq = e->d_freq_h.as<uint32_t>();
    A.selfscore = e->d_self_h.as<uint8_t>();
    A.ss3 = e->d_ss3.as<uint8_t>();
    A.g = &ds->bg;
    A.n_incid = &ds->n_incid;
    A.mid_list = e->d_mid_list.as<uint32_t>();
    A.mid_cnt = &ds->sx_mid_cnt;
    A.huge_list = e->d_huge_list.as<uint32_t>();
    A.huge_cnt = &ds->sx_huge_cnt;
    const bool small = e->cfg.bucket_cap == 512u;  // tests: tiny capacities push ordinary buckets down every path
    A.mid_cap = small ? 512u : 4096u;
    A.owner = owner;
    // one warp per bucket
#define KC_SXW(CROSS, WCAP)                                                                                      \
  do {                                                                                                           \
    constexpr size_t smem = sx_wb_warp_bytes<WCAP>() * kWbWarps + kWbTableBytes;                                 \
    KC_CUDA(e, cudaFuncSetAttribute((sx_warp_bucket_kernel<CROSS, WCAP>), cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                                 \
    int per_sm = 1;                                                                                              \
    KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (sx_warp_bucket_kernel<CROSS, WCAP>),      \
                                                             kWbWarps * 32, smem));                              \
    const uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>((NB + kWbWarps - 1) / kWbWarps,               \
                                                                     (uint32_t)(e->num_sm * std::max(per_sm, 1)))); \
    KC_LAUNCH(e, (sx_warp_bucket_kernel<CROSS, WCAP>), grid, kWbWarps * 32, smem, A);                            \
  } while (0)
    if (small) {
      if (A.first_after) KC_SXW(true, 64u); else KC_SXW(false, 64u);
    } else {
      if (A.first_after) KC_SXW(true, 512u); else KC_SXW(false, 512u);
    }
#undef KC_SXW
    // buckets beyond a warp's slice: one CTA each (the list is short on ordinary sets)
#define KC_SXB(CROSS, CAP)                                                                                       \
  do {                                                                                                           \
    constexpr size_t smem = sx_bucket_smem<CAP>();                                                               \
    KC_CUDA(e, cudaFuncSetAttribute((sx_bucket_kernel<CROSS, CAP>), cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                                 \
    int per_sm = 1;                                                                                              \
    KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (sx_bucket_kernel<CROSS, CAP>), CAP / 8, smem)); \
    const uint32_t grid = std::min<uint32_t>(NB, (uint32_t)(e->num_sm * std::max(per_sm, 1)));                    \
    KC_LAUNCH(e, (sx_bucket_kernel<CROSS, CAP>), grid, CAP / 8, smem, A);                                        \
  } while (0)
    if (small) {
      if (A.first_after) KC_SXB(true, 512u); else KC_SXB(false, 512u);
    } else {
      if (A.first_after) KC_SXB(true, 4096u); else KC_SXB(false, 4096u);
    }
#undef KC_SXB
    // buckets beyond a CTA's shared memory (k-mers with thousands of holders); the list is usually empty
    const uint32_t hgrid = std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)e->num_sm, std::max<uint32_t>(e->sx_huge_last, 8u)));
    if (A.first_after)
      KC_LAUNCH(e, sx_huge_kernel<true>, hgrid, kHugeThreads, 0, A);
    else
      KC_LAUNCH(e, sx_huge_kernel<false>, hgrid, kHugeThreads, 0, A);
  }
  // entry bins -> rows (laid out by capacity: row r starts at rowcap[r])
  KC_CUDA(e, cudaFuncSetAttribute(rows_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinSmemBytes));
  if (n_bins)
    KC_LAUNCH(e, rows_finalize_kernel, std::min<uint32_t>(n_bins, (uint32_t)e->num_sm * 3), kFinThreads, kFinSmemBytes, ent,
              rowcap, bin_cnt, n, 0u, n_bins, e->d_runs.as<uint4>(), e->d_run_cnt.as<uint32_t>(),
              e->d_rowlen.as<uint32_t>(), e->d_rowlen_p.as<uint32_t>(), e->d_ids.as<uint32_t>(), e->d_suf.as<uint2>(),
              e->cfg.want_blosum ? e->d_sufss.as<uint8_t>() : nullptr, e->d_rowwork64.as<unsigned long long>(),
              e->d_rowwork.as<uint32_t>(), e->d_rowinl.as<uint32_t>(), e->d_rowmaxlen.as<uint32_t>());
  if (n_shards <= 1)
    e->launches += exclusive_scan(WorkIn{e->d_rowwork.as<uint32_t>(), e->d_rowlen.as<uint32_t>()},
                                  U64ExclOutWithTail{e->d_workprefix.as<unsigned long long>(), n}, n, e->scan, e->stream);
  mark(e, EV_I1);
  DeviceScalars hs{};
  uint32_t n_records = 0;
  KC_CUDA(e, cudaMemcpyAsync(&hs, ds, sizeof(hs), cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaMemcpyAsync(&n_records, h2 + n_h2, 4, cudaMemcpyDeviceToHost, e->stream));
  // the host's own row tables, while the kernels above run (a deferred staging: kc_engine::pend_off)
  const int rc_staged = ensure_staged(e);
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  KC_CUDA(e, cudaGetLastError());
  if (rc_staged) return rc_staged;
  e->index_records = n_records;
  e->sx_huge_last = hs.sx_huge_cnt;
  e->sx_mid_last = hs.sx_mid_cnt;
  e->sx_max_bucket = hs.bg.max_bucket;
  // (sharded: totals over the k-mers whose first holder is one of this rank's rows; they add up over the ranks)
  e->istats.n_positions = own_positions;
  e->istats.n_incidences = hs.n_incid;
  e->istats.n_distinct = hs.bg.n_distinct;
  e->istats.n_repeated = hs.bg.n_repeated;
  e->istats.n_singleton = hs.bg.n_distinct - hs.bg.n_repeated;
  e->istats.nnz = hs.bg.nnz;
  e->v_local = E / 2 + 1;  // ids are placed by capacity (stream_index.cuh): the id SPACE, with holes
  e->multi_total = hs.bg.multi_total;
  e->work_total = hs.bg.work_total;
  e->ishard = shard;
  e->ishards = n_shards;
  e->block_bounds = bounds;
  e->n_own_rows = n_own_rows;
  if (stats) *stats = e->istats;
  e->bucketed = true;
  e->streamed = true;
  e->canonical_ready = false;
  e->have_index = true;
  return KC_OK;
}

// The canonical view of a partitioned index: ascending-k-mer ids, sorted id rows, kmer_freq by
// canonical id — the universe-table stages of index.cuh, unsliced, run on demand for the readback
// / lookup entry points (never on the hot path).  d_pk still holds every row's sorted distinct
// k-mers; it is rewritten to canonical ids in place exactly like the table build does.
static int ensure_canonical(kc_engine* e) {
  if (e->ishards > 1)
    return fail(e, KC_EINVAL, "the readback / lookup entry points need a whole index: kc_build_index, not a shard");
  if (!e->bucketed || e->canonical_ready) return KC_OK;
  const uint32_t n = (uint32_t)e->n;
  const uint64_t W = e->n_words, V = e->istats.n_repeated;
  DeviceScalars* ds = e->ds;
  KC_CUDA(e, e->d_seen.ensure(W * 8));
  KC_CUDA(e, e->d_dict.ensure(W * 8));
  KC_CUDA(e, e->d_zero.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rowlen_c.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_islo_c.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_vocab.ensure((V + 1) * 4));
  KC_CUDA(e, e->d_freq.ensure((V + 1) * 4));
  KC_CUDA(e, e->d_self.ensure(V + 16));
  KC_CUDA(e, e->d_pk.ensure((e->R + 64) * 4));
  KC_CUDA(e, e->d_ndist.ensure(((uint64_t)n + 1) * 4));
  int rc = ensure_scan(e, std::max<uint64_t>(W, (uint64_t)n + 1));
  if (rc) return rc;
  {  // the rows' sorted distinct k-mers (the partitioned build does not write them)
    DBuf none;
    std::swap(none, e->d_ksplit);
    rc = e->cfg.k == 5 ? run_extract_census<5>(e) : run_extract_census<7>(e);
    std::swap(none, e->d_ksplit);
    if (rc) return rc;
  }
  KC_CUDA(e, cudaMemsetAsync(e->d_seen.p, 0, W * 8, e->stream));
  KC_CUDA(e, cudaMemsetAsync(e->d_zero.p, 0, ((uint64_t)n + 1) * 4, e->stream));
  KC_CUDA(e, cudaMemsetAsync(e->d_rowlen_c.p, 0, ((uint64_t)n + 1) * 4, e->stream));
  KC_CUDA(e, cudaMemsetAsync(e->d_freq.p, 0, (V + 1) * 4, e->stream));
  const uint32_t grid = blocks_for(n, 8, e->num_sm * 8);
  KC_LAUNCH(e, (census_pass_kernel<32, false>), grid, 256, 0, e->d_pk.as<uint32_t>(), e->d_pstart.as<uint32_t>(),
            e->d_zero.as<uint32_t>(), e->d_ndist.as<uint32_t>(), n, e->d_seen.as<uint32_t>());
  e->launches += exclusive_scan(PopcIn<1>{e->d_seen.as<uint32_t>()},
                                DictOut<1>{e->d_seen.as<uint32_t>(), e->d_dict.as<uint2>()}, W, e->scan, e->stream);
  unsigned long long v_check = 0;
  KC_CUDA(e, cudaMemcpyAsync(&v_check, &ds->scan_total, 8, cudaMemcpyDeviceToHost, e->stream));
  if (V)
    KC_LAUNCH(e, expand_bitmap_kernel, blocks_for(W, 256, e->num_sm * 16), 256, 0, e->d_dict.as<uint2>(), W,
              e->d_vocab.as<uint32_t>(), e->d_self.as<uint8_t>(), e->cfg.k);
  KC_LAUNCH(e, ids_freq_kernel<32>, grid, 256, 0, e->d_dict.as<uint2>(), e->d_pstart.as<uint32_t>(),
            e->d_zero.as<uint32_t>(), e->d_ndist.as<uint32_t>(), n, e->d_pk.as<uint32_t>(),
            e->d_rowlen_c.as<uint32_t>(), e->d_islo_c.as<uint32_t>(), nullptr, e->d_freq.as<uint32_t>(), &ds->nnz);
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  KC_CUDA(e, cudaGetLastError());
  if (v_check != V) return fail(e, KC_ECUDA, "partitioned index and canonical view disagree on the vocabulary size");
  e->canonical_ready = true;
  return KC_OK;
}

// =========================================================================================
extern "C" {

int kc_abi_version(void) { return KC_ABI_VERSION; }

uint32_t kc_sample_position(uint64_t seed, uint32_t protein, uint32_t n_positions, uint32_t x) {
  return n_positions ? sample_perm(sample_key(seed, protein), n_positions, x % n_positions) : 0u;
}

int kc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int kc_create(const kc_config* cfg, kc_engine** out) {
  if (!cfg || !out) return KC_EINVAL;
  *out = nullptr;
  if (cfg->k != 5 && cfg->k != 7) return KC_EINVAL;  // src/tree.rs:104 panics on anything else
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return KC_ENODEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) return KC_ENODEVICE;
  kc_engine* e = new kc_engine();
  e->cfg = *cfg;
  e->dev = cfg->device;
  if (cudaSetDevice(e->dev) != cudaSuccess) {
    delete e;
    return KC_ENODEVICE;
  }
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, e->dev);
  e->num_sm = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : kNumSM;
  e->smem_optin = prop.sharedMemPerBlockOptin;
  e->total_mem = prop.totalGlobalMem;
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete e;
    return KC_ECUDA;
  }
  e->own_stream = true;
  for (int i = 0; i < EV_COUNT; ++i) cudaEventCreate(&e->ev[i]);
  cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < kc_engine::kUploadChunks; ++i) cudaEventCreateWithFlags(&e->chunk_ev[i], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->main_ev, cudaEventDisableTiming);
  // residue LUT: src/protein.rs:9-13 order, everything else -> 20 (src/protein.rs:49-54)
  uint8_t lut[256];
  std::memset(lut, 20, sizeof(lut));
  const char* alphabet = "CSTAGPDEQNHRKMILVWYF*";
  for (int i = 0; i < 21; ++i) lut[(unsigned char)alphabet[i]] = (uint8_t)i;
  if (cudaMemcpyToSymbol(c_residue_lut, lut, 256) != cudaSuccess ||
      e->d_scalars.ensure(sizeof(DeviceScalars)) != cudaSuccess) {
    kc_destroy(e);
    return KC_ECUDA;
  }
  e->ds = e->d_scalars.as<DeviceScalars>();
  cudaMemset(e->ds, 0, sizeof(DeviceScalars));
  if (e->d_ss3.ensure(9261 + 64) != cudaSuccess) {
    kc_destroy(e);
    return KC_ENOMEM;
  }
  sx_ss3_kernel<<<(9261 + 255) / 256, 256, 0, e->stream>>>(e->d_ss3.as<uint8_t>());
  e->universe = pow21(cfg->k);
  e->n_words = ((uint64_t)e->universe + 31) / 32;
  *out = e;
  return KC_OK;
}

void kc_destroy(kc_engine* e) {
  if (!e) return;
  cudaSetDevice(e->dev);
  cudaDeviceSynchronize();
  DBuf* all[] = {&e->d_res, &e->d_off, &e->d_kpos, &e->d_pstart, &e->d_plen, &e->d_orig, &e->d_rank,
                 &e->d_first_after, &e->d_long, &e->d_huge, &e->d_cta, &e->d_vlong, &e->d_huge_off, &e->d_huge_scratch, &e->d_pk,
                 &e->d_ndist, &e->d_rowlen, &e->d_seen, &e->d_dict, &e->d_vocab, &e->d_freq,
                 &e->d_self, &e->d_colptr, &e->d_cursor, &e->d_col, &e->d_suf, &e->d_sufss, &e->d_rowwork, &e->d_lists,
                 &e->d_colscratch, &e->d_workprefix, &e->d_ksplit, &e->d_isplit, &e->d_rowwork64, &e->d_rowinl, &e->d_rowmaxlen, &e->d_psplit, &e->d_rowbase, &e->d_plist, &e->d_pss,
                 &e->d_rec, &e->d_entries, &e->d_bin_cnt, &e->d_rowcap, &e->d_bucket_cnt, &e->d_ids,
                 &e->d_vocab_h, &e->d_freq_h, &e->d_self_h, &e->d_zero, &e->d_rowlen_c, &e->d_islo_c, &e->d_filter, &e->d_binowner, &e->d_runs, &e->d_run_cnt, &e->d_rowlen_p, &e->d_rowbin, &e->d_rowsafe, &e->d_rowlogh, &e->d_edges, &e->d_keys_a, &e->d_keys_b,
                 &e->d_vals_a, &e->d_vals_b, &e->d_hist, &e->d_edges_sorted, &e->d_scalars, &e->d_scan_tiles,
                 &e->d_tmp, &e->d_sres_own, &e->d_soff, &e->d_rec_a, &e->d_rec_b, &e->d_h1, &e->d_h2,
                 &e->d_huge_list, &e->d_mid_list, &e->d_tile_row, &e->d_ss3, &e->d_comm, &e->d_gather, &e->d_keepmask};
  if (e->comm && nccl_api().ok()) nccl_api().CommDestroy(e->comm);
  e->comm = nullptr;
  for (DBuf* b : all) b->release();
  for (int i = 0; i < EV_COUNT; ++i)
    if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  for (int i = 0; i < kc_engine::kUploadChunks; ++i)
    if (e->chunk_ev[i]) cudaEventDestroy(e->chunk_ev[i]);
  if (e->main_ev) cudaEventDestroy(e->main_ev);
  delete e;
}

const char* kc_last_error(const kc_engine* e) { return e ? e->err.c_str() : "null engine"; }

int kc_set_stream(kc_engine* e, void* cuda_stream) {
  if (!e) return KC_EINVAL;
  cudaSetDevice(e->dev);
  if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
  e->n_chunks = 0;
  cudaStreamSynchronize(e->stream);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  e->own_stream = false;
  if (cuda_stream) {
    e->stream = (cudaStream_t)cuda_stream;
  } else {
    KC_CUDA(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->own_stream = true;
  }
  return KC_OK;
}

// dist: every rank uploads 1 / world of the residue stream and the slices are all-gathered over NVLink
static int set_proteins_host(kc_engine* e, const uint8_t* residues, const uint64_t* offsets, const uint32_t* class_id,
                             uint64_t n, bool dist) {
  if (!e || !offsets || (n && !class_id)) return KC_EINVAL;
  KC_CUDA(e, cudaSetDevice(e->dev));
  // (a chunked upload of the previous call may still be in flight into d_res; the host buffers of THIS call
  // must stay valid until the next kc_build_index* / kc_extract_kmers returns: the copies are asynchronous)
  if (int rcw = wait_upload(e)) return rcw;
  e->have_proteins = e->have_index = e->have_pairs = false;
  // What sizes the copies is checked first (they are queued at once and run while the host does the rest);
  // the engine is only marked as holding proteins after the whole offsets array has been validated below.
  if (offsets[0] != 0) return fail(e, KC_EINVAL, "offsets[0] must be 0");
  const uint64_t R = offsets[n];
  if (R && !residues) return fail(e, KC_EINVAL, "null residues");
  if (R >= 0xFFFF0000ull) return fail(e, KC_ETOOLARGE, "more than 2^32-65536 residues");
  if (n >= 0xFFFFFFF0ull) return fail(e, KC_ETOOLARGE, "too many proteins");
  e->n = n;
  e->pend_off = nullptr;
  e->pend_cls = nullptr;
  mark(e, EV_H2D0);
  size_t padded = padded_res_bytes(R);
  const uint64_t world = dist ? (uint64_t)e->cworld : 1;
  const uint64_t per = ((R + world - 1) / world + 255) & ~255ull;  // slice of the all-gather
  if (dist) padded = std::max<size_t>(padded, (size_t)(per * world) + kSxTile);
  KC_CUDA(e, e->d_res.ensure(padded));
  // the offsets first: the one H2D copy engine serves the copies in issue order, and the layout
  // kernel (main stream) must not wait behind the whole residue stream
  KC_CUDA(e, e->d_off.ensure((n + 1) * 8));
  KC_CUDA(e, cudaMemcpyAsync(e->d_off.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, e->stream));
  e->n_chunks = 0;
  constexpr int C = kc_engine::kUploadChunks;
  if (dist && world > 1) {
    NcclApi& nc = nccl_api();
    if (!e->comm || !nc.ok()) return fail(e, KC_EINVAL, "kc_comm_init first");
    const uint64_t lo = std::min<uint64_t>(R, per * (uint64_t)e->crank), hi = std::min<uint64_t>(R, lo + per);
    if (hi > lo)
      KC_CUDA(e, cudaMemcpyAsync(e->d_res.as<uint8_t>() + lo, residues + lo, hi - lo, cudaMemcpyHostToDevice, e->stream));
    const ncclResult_t nr = nc.AllGather(e->d_res.as<uint8_t>() + per * (uint64_t)e->crank, e->d_res.p, per, ncclUint8,
                                         e->comm, e->stream);
    if (nr != ncclSuccess) return fail(e, KC_ECUDA, std::string("ncclAllGather: ") + nc.GetErrorString(nr));
    KC_CUDA(e, cudaMemsetAsync(e->d_res.as<uint8_t>() + R, 0, padded - R, e->stream));
  } else if (!e->cfg.cross_class_only && e->copy_stream && e->cfg.no_upload_overlap != 1u &&
             ((R >= (32u << 20) && n >= 64 * C) || (e->cfg.no_upload_overlap == 2u && R >= 4ull * C * kSxTile))) {
    // pair order = input order.  The stream is cut at tile borders of the streaming build (chunk c = the bytes
    // [chunk_byte[c], chunk_byte[c+1])); chunk_row[c] = the rows that lie entirely inside the chunks before c
    // (what the row-chunked extract kernels of the other builds may read after event c - 1).
    // The copy stream waits for what the main stream still does with the old residues.
    KC_CUDA(e, cudaEventRecord(e->main_ev, e->stream));
    KC_CUDA(e, cudaStreamWaitEvent(e->copy_stream, e->main_ev, 0));
    e->chunk_row[0] = 0;
    e->chunk_byte[0] = 0;
    for (int c = 1; c <= C; ++c) {
      const uint64_t cut = c == C ? R : std::min<uint64_t>(R, (R / C * c + kSxTile - 1) / kSxTile * kSxTile);
      e->chunk_byte[c] = std::max(cut, e->chunk_byte[c - 1]);
      // (the offsets are validated below; whatever they hold, the rows stay monotone and within [0, n])
      const uint64_t r = c == C ? n : (uint64_t)(std::upper_bound(offsets, offsets + n + 1, e->chunk_byte[c]) - offsets);
      e->chunk_row[c] = std::max(e->chunk_row[c - 1], (uint32_t)std::min<uint64_t>(n, c == C ? n : (r ? r - 1 : 0)));
    }
    for (int c = 0; c < C; ++c) {
      const uint64_t b0 = e->chunk_byte[c], b1 = e->chunk_byte[c + 1];
      if (b1 > b0)
        KC_CUDA(e, cudaMemcpyAsync(e->d_res.as<uint8_t>() + b0, residues + b0, b1 - b0, cudaMemcpyHostToDevice,
                                   e->copy_stream));
      if (c + 1 == C) KC_CUDA(e, cudaMemsetAsync(e->d_res.as<uint8_t>() + R, 0, padded - R, e->copy_stream));
      KC_CUDA(e, cudaEventRecord(e->chunk_ev[c], e->copy_stream));
    }
    e->n_chunks = C;
  } else {
    if (R) KC_CUDA(e, cudaMemcpyAsync(e->d_res.p, residues, R, cudaMemcpyHostToDevice, e->stream));
    KC_CUDA(e, cudaMemsetAsync(e->d_res.as<uint8_t>() + R, 0, padded - R, e->stream));
  }
  e->R = R;
  int rc = KC_OK;
  if (!e->cfg.cross_class_only) {
    // pair order = input order: the row tables come straight from the caller's offsets (one pass in parallel
    // slabs while the upload runs; it also checks that they do not decrease), the device derives the row
    // layout from d_off, and the engine's own copies of offsets / classes are deferred (kc_engine::pend_off)
    rc = stage_layout_host(e, offsets);
    if (rc == KC_OK) rc = stage_layout_device(e);
    if (rc == KC_OK) rc = stage_lists_device(e);
    if (rc == KC_OK) {
      e->pend_off = offsets;
      e->pend_cls = class_id;
      e->have_proteins = true;
      e->have_index = e->have_pairs = false;
    }
  } else {
    copy_host_arrays(e, offsets, class_id);
    rc = stage_layout(e);
  }
  if (e->n_chunks) {
    cudaEventRecord(e->ev[EV_H2D1], e->copy_stream);
    e->ev_set[EV_H2D1] = true;
  } else {
    mark(e, EV_H2D1);
  }
  return rc;
}

int kc_set_proteins(kc_engine* e, const uint8_t* residues, const uint64_t* offsets, const uint32_t* class_id,
                    uint64_t n) {
  return set_proteins_host(e, residues, offsets, class_id, n, false);
}

int kc_set_proteins_device(kc_engine* e, const uint8_t* d_residues, const uint64_t* d_offsets,
                           const uint32_t* d_class_id, uint64_t n) {
  if (!e || !d_offsets || (n && !d_class_id)) return KC_EINVAL;
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (int rcw = wait_upload(e)) return rcw;
  e->have_proteins = e->have_index = e->have_pairs = false;
  e->pend_off = nullptr;
  e->pend_cls = nullptr;
  e->n = n;
  e->h_off.resize(n + 1);
  e->h_cls.resize(n);
  // the layout tables (pair order, length classes) are derived on the host from the offsets
  // and classes: 12 bytes per protein come back, the residues stay in HBM
  KC_CUDA(e, cudaMemcpyAsync(e->h_off.data(), d_offsets, (n + 1) * 8, cudaMemcpyDeviceToHost, e->stream));
  if (n) KC_CUDA(e, cudaMemcpyAsync(e->h_cls.data(), d_class_id, n * 4, cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  if (e->h_off[0] != 0) return fail(e, KC_EINVAL, "offsets[0] must be 0");
  for (uint64_t p = 0; p < n; ++p)
    if (e->h_off[p + 1] < e->h_off[p]) return fail(e, KC_EINVAL, "offsets must be non-decreasing");
  const uint64_t R = e->h_off[n];
  if (R >= 0xFFFF0000ull) return fail(e, KC_ETOOLARGE, "more than 2^32-65536 residues");
  mark(e, EV_H2D0);
  const size_t padded = padded_res_bytes(R);
  KC_CUDA(e, e->d_res.ensure(padded));
  if (R) KC_CUDA(e, cudaMemcpyAsync(e->d_res.p, d_residues, R, cudaMemcpyDeviceToDevice, e->stream));
  KC_CUDA(e, cudaMemsetAsync(e->d_res.as<uint8_t>() + R, 0, padded - R, e->stream));
  KC_CUDA(e, e->d_off.ensure((n + 1) * 8));
  KC_CUDA(e, cudaMemcpyAsync(e->d_off.p, d_offsets, (n + 1) * 8, cudaMemcpyDeviceToDevice, e->stream));
  int rc = stage_layout(e);
  mark(e, EV_H2D1);
  return rc;
}

int kc_set_proteins_device_residues(kc_engine* e, const uint8_t* d_residues, const uint64_t* offsets,
                                    const uint32_t* class_id, uint64_t n) {
  if (!e || !offsets || (n && !class_id)) return KC_EINVAL;
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (int rcw = wait_upload(e)) return rcw;
  e->have_proteins = e->have_index = e->have_pairs = false;
  e->pend_off = nullptr;
  e->pend_cls = nullptr;
  if (offsets[0] != 0) return fail(e, KC_EINVAL, "offsets[0] must be 0");
  for (uint64_t p = 0; p < n; ++p)
    if (offsets[p + 1] < offsets[p]) return fail(e, KC_EINVAL, "offsets must be non-decreasing");
  e->n = n;
  e->h_off.assign(offsets, offsets + n + 1);
  e->h_cls.assign(class_id, class_id + n);
  const uint64_t R = offsets[n];
  if (R && !d_residues) return fail(e, KC_EINVAL, "null residues");
  if (R >= 0xFFFF0000ull) return fail(e, KC_ETOOLARGE, "more than 2^32-65536 residues");
  mark(e, EV_H2D0);
  const size_t padded = padded_res_bytes(R);
  KC_CUDA(e, e->d_res.ensure(padded));
  KC_CUDA(e, e->d_off.ensure((n + 1) * 8));
  KC_CUDA(e, cudaMemcpyAsync(e->d_off.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, e->stream));
  if (R) KC_CUDA(e, cudaMemcpyAsync(e->d_res.p, d_residues, R, cudaMemcpyDeviceToDevice, e->stream));
  KC_CUDA(e, cudaMemsetAsync(e->d_res.as<uint8_t>() + R, 0, padded - R, e->stream));
  int rc = stage_layout(e);
  mark(e, EV_H2D1);
  return rc;
}

int kc_extract_kmers(kc_engine* e, uint32_t* kmers_out, uint64_t capacity, uint64_t* n_positions) {
  if (!e) return KC_EINVAL;
  if (!e->have_proteins) return fail(e, KC_EINVAL, "kc_set_proteins first");
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (int rcs = ensure_staged(e)) return rcs;
  const uint64_t n = e->n;
  const int k = e->cfg.k;
  std::vector<unsigned long long> kpos(n + 1, 0);
  for (uint64_t p = 0; p < n; ++p) {
    const uint64_t len = e->h_off[p + 1] - e->h_off[p];
    const uint64_t every = e->cfg.sample_every > 1 ? e->cfg.sample_every : 1;
    kpos[p + 1] = kpos[p] + (len >= (uint64_t)k ? (len - k + 1) / every : 0);
  }
  const uint64_t npos = kpos[n];
  if (n_positions) *n_positions = npos;
  if (kmers_out && capacity < npos) return fail(e, KC_EINVAL, "kmers_out capacity too small");
  if (int rcw = wait_upload(e)) return rcw;
  KC_CUDA(e, e->d_kpos.ensure((n + 1) * 8));
  KC_CUDA(e, cudaMemcpyAsync(e->d_kpos.p, kpos.data(), (n + 1) * 8, cudaMemcpyHostToDevice, e->stream));
  KC_CUDA(e, e->d_tmp.ensure(std::max<uint64_t>(npos, 4) * 4));
  mark(e, EV_X0);
  int rc = k == 5 ? run_positions<5>(e, e->d_tmp.as<uint32_t>()) : run_positions<7>(e, e->d_tmp.as<uint32_t>());
  mark(e, EV_X1);
  if (rc) return rc;
  KC_CUDA(e, cudaGetLastError());
  if (kmers_out && npos)
    KC_CUDA(e, cudaMemcpyAsync(kmers_out, e->d_tmp.p, npos * 4, cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  return KC_OK;
}

int kc_build_index(kc_engine* e, kc_index_stats* stats) { return kc_build_index_shard(e, 0, 1, stats); }

int kc_index_shard_info(kc_engine* e, uint32_t info[4]) {
  if (!e || !info) return KC_EINVAL;
  if (!e->have_index) return fail(e, KC_EINVAL, "kc_build_index first");
  info[0] = e->ishard;
  info[1] = e->ishards;
  info[2] = e->ishards > 1 ? (uint32_t)e->block_bounds.size() - 1 : 1u;
  info[3] = e->ishards > 1 ? e->n_own_rows : (uint32_t)e->n;
  return KC_OK;
}

int kc_index_flavour(kc_engine* e) {
  if (!e || !e->have_index) return -1;
  return e->streamed ? 1 : (e->bucketed ? (int)e->cap_hint : 0);
}

int kc_index_shard_blocks(kc_engine* e, uint32_t* bounds, uint32_t capacity) {
  if (!e || !bounds) return KC_EINVAL;
  if (!e->have_index) return fail(e, KC_EINVAL, "kc_build_index first");
  if (e->ishards <= 1) {
    if (capacity < 2) return fail(e, KC_EINVAL, "capacity too small");
    bounds[0] = 0;
    bounds[1] = (uint32_t)e->n;
    return KC_OK;
  }
  if (capacity < e->block_bounds.size()) return fail(e, KC_EINVAL, "capacity too small");
  for (size_t i = 0; i < e->block_bounds.size(); ++i) bounds[i] = e->block_bounds[i];
  return KC_OK;
}

static int build_index_shard_impl(kc_engine* e, uint32_t shard, uint32_t n_shards, kc_index_stats* stats);

int kc_build_index_shard(kc_engine* e, uint32_t shard, uint32_t n_shards, kc_index_stats* stats) {
  const int rc = build_index_shard_impl(e, shard, n_shards, stats);
  // the caller's offsets / classes are only borrowed until this call returns (kc_engine::pend_off; the hot
  // builds have made the copies behind their kernels already)
  if (e) ensure_staged(e);
  return rc;
}

static int build_index_shard_impl(kc_engine* e, uint32_t shard, uint32_t n_shards, kc_index_stats* stats) {
  if (!e || n_shards == 0 || shard >= n_shards) return KC_EINVAL;
  if (n_shards > 255) return fail(e, KC_EINVAL, "at most 255 shards (the row blocks' owner is one byte)");
  if (!e->have_proteins) return fail(e, KC_EINVAL, "kc_set_proteins first");
  KC_CUDA(e, cudaSetDevice(e->dev));
  const uint32_t n = (uint32_t)e->n;
  const uint64_t R = e->R, W = e->n_words;
  DeviceScalars* ds = e->ds;
  e->have_index = e->have_pairs = false;
  e->bucketed = e->canonical_ready = false;
  e->ishard = 0;
  e->ishards = 1;
  e->streamed = false;
  {
    // Which build (kc_config.index_build): the streaming partitioned build (stream_index.cuh) by default.
    // It has no subsampling mode (Protein::new_with_rand_fivemers is dead code in the reference): that runs
    // the table build.
    // Sharded (multi-GPU) builds: round 1's bucket build is the faster one there (its extract pass dedups per
    // row while it scans the foreign rows; measured at 8 shards: index 8.3 ms against 12.2 ms), so it is tried
    // first; when one of its buckets overflows (a k-mer with thousands of holders) the rank builds the SAME
    // shard with the streaming build, which has no capacity to overflow.  Both cut the same row blocks and
    // give the same per-rank results, so ranks may differ in which build they ran (a fallback to
    // the replicated table build left ranks scoring incompatible row partitions).
    uint32_t want = e->cfg.index_build;
    const bool subsampled = e->cfg.sample_every > 1;
    if (want == KC_INDEX_AUTO) want = subsampled ? KC_INDEX_TABLE : (n_shards > 1 ? KC_INDEX_BUCKET : KC_INDEX_STREAM);
    if (want == KC_INDEX_STREAM && (subsampled || n >= (1u << 24))) want = subsampled ? KC_INDEX_TABLE : KC_INDEX_BUCKET;
    if (want == KC_INDEX_STREAM && n > 0) return build_index_stream(e, shard, n_shards, stats);
    // (the bucket build remembers per protein-set signature that its buckets overflowed)
    const bool known_overflow = e->cap_hint == 0 && e->cap_hint_n == e->n && e->cap_hint_R == e->R;
    if (want == KC_INDEX_BUCKET && known_overflow && !subsampled && n > 0 && n < (1u << 24) &&
        (n_shards > 1 || e->cfg.index_build != KC_INDEX_BUCKET))
      return build_index_stream(e, shard, n_shards, stats);
    if (want == KC_INDEX_BUCKET && n > 0 && n <= (1u << 24)) {
      bool overflow = false;
      int rc = build_index_bucketed(e, shard, n_shards, stats, &overflow);
      if (rc != KC_OK || !overflow) return rc;
      // k-mers (or clumps of them) with more holders than a shared-memory bucket takes
      e->cap_hint = 0;
      e->cap_hint_n = e->n;
      e->cap_hint_R = e->R;
      if (!subsampled && n < (1u << 24) && (n_shards > 1 || e->cfg.index_build != KC_INDEX_BUCKET))
        return build_index_stream(e, shard, n_shards, stats);
      // (an explicit KC_INDEX_BUCKET on one GPU keeps round 1's behaviour: the table build)
    }
  }
  KC_CUDA(e, e->d_pk.ensure((R + 64) * 4));
  KC_CUDA(e, e->d_ndist.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rowlen.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_seen.ensure(W * 8));  // 2 bits per k-mer
  // L2 blocking plan: slices of 2^slice_shift k-mers such that the randomly accessed part of
  // every sliced pass (census state / dictionary + freq / cursor + postings) is <= ~64 MB
  {
    const unsigned long long n_positions_est = e->n_pos_unsampled;
    const double footprint = std::max((double)e->universe / 4.0 * 1.5, (double)n_positions_est * 5.0);
    uint32_t want = (uint32_t)std::min(64.0, std::ceil(footprint / (64.0 * 1024 * 1024)));
    if (e->cfg.index_slices) want = e->cfg.index_slices;
    uint32_t shift = 31;
    while (shift > 12 && ((((uint64_t)e->universe - 1) >> shift) + 1) < want) --shift;
    e->slice_shift = shift;
    e->n_slices = (uint32_t)((((uint64_t)e->universe - 1) >> shift) + 1);
  }
  const uint32_t P = e->n_slices;
  KC_CUDA(e, e->d_ksplit.ensure(((uint64_t)P + 1) * std::max<uint64_t>(n, 1) * 4));
  KC_CUDA(e, e->d_isplit.ensure(((uint64_t)P + 1) * std::max<uint64_t>(n, 1) * 4));
  KC_CUDA(e, e->d_psplit.ensure(((uint64_t)P + 1) * std::max<uint64_t>(n, 1) * 4));
  KC_CUDA(e, e->d_rowbase.ensure(((uint64_t)n + 2) * 8));
  e->have_plist = false;
  KC_CUDA(e, e->d_rowwork64.ensure(((uint64_t)n + 1) * 8));
  KC_CUDA(e, e->d_rowinl.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_rowmaxlen.ensure(((uint64_t)n + 1) * 4));
  if (n) {
    KC_CUDA(e, cudaMemsetAsync(e->d_rowinl.p, 0, (uint64_t)n * 4, e->stream));
    KC_CUDA(e, cudaMemsetAsync(e->d_rowmaxlen.p, 0, (uint64_t)n * 4, e->stream));
  }
  KC_CUDA(e, e->d_dict.ensure(W * 8));
  int rc = ensure_scan(e, std::max<uint64_t>(W, (uint64_t)n + 1));
  if (rc) return rc;

  mark(e, EV_I0);
  KC_CUDA(e, cudaMemsetAsync(e->d_seen.p, 0, W * 8, e->stream));
  KC_CUDA(e, cudaMemsetAsync(ds, 0, sizeof(DeviceScalars), e->stream));
  const unsigned long long n_positions = e->h_pospref[n];  // positions are known on the host

  // K1-K3: extract, per-protein dedup, census bitmaps
  mark(e, EV_IC0);
  rc = e->cfg.k == 5 ? run_extract_census<5>(e) : run_extract_census<7>(e);
  if (rc) return rc;
  const bool narrow = P >= 4;  // few entries per (row, slice): 8 lanes per row instead of 32
  const uint32_t pass_grid = blocks_for(n, narrow ? 32 : 8, e->num_sm * 8);
  auto split = [&](DBuf& b, uint32_t q) { return b.as<uint32_t>() + (size_t)q * n; };
  // the census state is only 2 bits per k-mer: it can take coarser slices (longer row runs)
  const uint32_t cmerge = std::max<uint32_t>(1u, e->cfg.census_merge);
  for (uint32_t q = 0; q < P && n; q += cmerge) {
    const uint32_t q1 = std::min(P, q + cmerge);
    const bool preread = e->universe <= (1u << 24);  // hot k-mers (k=5): skip marks that are no-ops
    const uint32_t cgrid = narrow && cmerge < 4 ? pass_grid : blocks_for(n, 8, e->num_sm * 8);
#define KC_CENSUS(G, PRE)                                                                              \
  KC_LAUNCH(e, (census_pass_kernel<G, PRE>), cgrid, 256, 0, e->d_pk.as<uint32_t>(), e->d_pstart.as<uint32_t>(), \
            split(e->d_ksplit, q), split(e->d_ksplit, q1), n, e->d_seen.as<uint32_t>())
    if (narrow && cmerge < 4) {
      if (preread) KC_CENSUS(8, true); else KC_CENSUS(8, false);
    } else {
      if (preread) KC_CENSUS(32, true); else KC_CENSUS(32, false);
    }
#undef KC_CENSUS
  }
  mark(e, EV_IC1);
  // K4: rank dictionary over "held by >= 2 proteins"
  e->launches += exclusive_scan(PopcIn<1>{e->d_seen.as<uint32_t>()},
                                DictOut<1>{e->d_seen.as<uint32_t>(), e->d_dict.as<uint2>()}, W, e->scan, e->stream);
  KC_CUDA(e, cudaMemcpyAsync(&ds->n_repeated, &ds->scan_total, 8, cudaMemcpyDeviceToDevice, e->stream));
  KC_LAUNCH(e, popc_reduce_kernel, blocks_for(2 * W, 256 * 8, e->num_sm * 8), 256, 0, e->d_seen.as<uint32_t>(),
            2 * W, &ds->n_distinct);
  DeviceScalars hs{};
  KC_CUDA(e, cudaMemcpyAsync(&hs, ds, sizeof(hs), cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  KC_CUDA(e, cudaGetLastError());
  const uint64_t V = hs.n_repeated, n_incid = hs.n_incid;
  if (V >= 0xFFFFFFF0ull || n_incid >= 0xFFFFFFF0ull) return fail(e, KC_ETOOLARGE, "index too large");

  KC_CUDA(e, e->d_vocab.ensure((V + 1) * 4));
  KC_CUDA(e, e->d_freq.ensure((V + 1) * 4));
  KC_CUDA(e, e->d_self.ensure(V + 16));
  KC_CUDA(e, e->d_colptr.ensure((V + 2) * 4));
  KC_CUDA(e, e->d_cursor.ensure((V + 2) * 4));
  KC_CUDA(e, e->d_col.ensure((n_incid + 64) * 4));
  KC_CUDA(e, e->d_suf.ensure((R + 64) * 8));
  if (e->cfg.want_blosum) KC_CUDA(e, e->d_sufss.ensure(R + 64));
  KC_CUDA(e, e->d_rowwork.ensure(((uint64_t)n + 1) * 4));
  KC_CUDA(e, e->d_workprefix.ensure(((uint64_t)n + 2) * 8));
  const uint64_t list_cap = n_incid / 33 + 16;
  KC_CUDA(e, e->d_lists.ensure(list_cap * 3 * 4));
  rc = ensure_scan(e, std::max<uint64_t>(V + 1, (uint64_t)n + 1));
  if (rc) return rc;
  KC_CUDA(e, cudaMemsetAsync(e->d_freq.p, 0, (V + 1) * 4, e->stream));

  if (V) {
    KC_LAUNCH(e, expand_bitmap_kernel, blocks_for(W, 256, e->num_sm * 16), 256, 0, e->d_dict.as<uint2>(), W,
              e->d_vocab.as<uint32_t>(), e->d_self.as<uint8_t>(), e->cfg.k);
  }
  // K5: ids (in place over the distinct k-mers), row lengths, kmer_freq — one launch per slice
  if (n) KC_CUDA(e, cudaMemsetAsync(e->d_rowlen.p, 0, (uint64_t)n * 4, e->stream));
  for (uint32_t q = 0; q < P && n; ++q) {
    uint32_t* last = q + 1 == P ? split(e->d_isplit, P) : nullptr;
    if (narrow)
      KC_LAUNCH(e, ids_freq_kernel<8>, pass_grid, 256, 0, e->d_dict.as<uint2>(), e->d_pstart.as<uint32_t>(),
                split(e->d_ksplit, q), split(e->d_ksplit, q + 1), n, e->d_pk.as<uint32_t>(),
                e->d_rowlen.as<uint32_t>(), split(e->d_isplit, q), last, e->d_freq.as<uint32_t>(), &ds->nnz);
    else
      KC_LAUNCH(e, ids_freq_kernel<32>, pass_grid, 256, 0, e->d_dict.as<uint2>(), e->d_pstart.as<uint32_t>(),
                split(e->d_ksplit, q), split(e->d_ksplit, q + 1), n, e->d_pk.as<uint32_t>(),
                e->d_rowlen.as<uint32_t>(), split(e->d_isplit, q), last, e->d_freq.as<uint32_t>(), &ds->nnz);
  }
  // postings: colptr = exclusive scan of freq, fill, sort every column by protein rank
  e->launches += exclusive_scan(U32In{e->d_freq.as<uint32_t>()}, ColptrOut{e->d_colptr.as<uint32_t>()}, V + 1,
                                e->scan, e->stream);
  KC_CUDA(e, cudaMemcpyAsync(e->d_cursor.p, e->d_colptr.p, (V + 1) * 4, cudaMemcpyDeviceToDevice, e->stream));
  if (n && V) {
    for (uint32_t q = 0; q < P; ++q) {
      if (narrow)
        KC_LAUNCH(e, postings_fill_kernel<8>, pass_grid, 256, 0, e->d_pstart.as<uint32_t>(), split(e->d_isplit, q),
                  split(e->d_isplit, q + 1), n, e->d_pk.as<uint32_t>(), e->d_cursor.as<uint32_t>(),
                  e->d_col.as<uint32_t>());
      else
        KC_LAUNCH(e, postings_fill_kernel<32>, pass_grid, 256, 0, e->d_pstart.as<uint32_t>(), split(e->d_isplit, q),
                  split(e->d_isplit, q + 1), n, e->d_pk.as<uint32_t>(), e->d_cursor.as<uint32_t>(),
                  e->d_col.as<uint32_t>());
    }
    uint32_t* lists = e->d_lists.as<uint32_t>();
    KC_LAUNCH(e, postings_sort_small_kernel, blocks_for(V, 256, e->num_sm * 8), 256, 0,
              e->d_colptr.as<uint32_t>(), (uint32_t)V, e->d_col.as<uint32_t>(), lists, lists + list_cap,
              lists + 2 * list_cap, ds->list_counts);
    if (n > 32) {
      KC_LAUNCH(e, postings_sort_mid_kernel, e->num_sm * 8, 128, 0, e->d_colptr.as<uint32_t>(), lists,
                ds->list_counts, e->d_col.as<uint32_t>());
    }
    if (n > kColWarpMax) {
      const uint32_t np2 = std::min<uint32_t>(next_pow2_u32(n), kColBlockMax);
      const size_t smem = (size_t)np2 * 4;
      KC_CUDA(e, cudaFuncSetAttribute(postings_sort_big_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      KC_LAUNCH(e, postings_sort_big_kernel<false>, e->num_sm, 512, smem, e->d_colptr.as<uint32_t>(),
                lists + list_cap, ds->list_counts, 1, e->d_col.as<uint32_t>(), nullptr, 0u);
    }
    if (n > kColBlockMax) {
      const uint32_t stride = next_pow2_u32(n);
      const uint32_t ctas = 32;
      KC_CUDA(e, e->d_colscratch.ensure((size_t)stride * ctas * 4));
      KC_LAUNCH(e, postings_sort_big_kernel<true>, ctas, 512, 0, e->d_colptr.as<uint32_t>(), lists + 2 * list_cap,
                ds->list_counts, 2, e->d_col.as<uint32_t>(), e->d_colscratch.as<uint32_t>(), stride);
    }
    KC_LAUNCH(e, multi_edge_total_kernel, blocks_for(V, 256 * 4, e->num_sm * 4), 256, 0, e->d_freq.as<uint32_t>(),
              (uint32_t)V, &ds->multi_total);
    KC_CUDA(e, cudaMemsetAsync(e->d_rowwork64.p, 0, (uint64_t)n * 8, e->stream));
    const uint32_t* fa = e->cfg.cross_class_only ? e->d_first_after.as<uint32_t>() : nullptr;
    for (uint32_t q = 0; q < P; ++q) {
      if (narrow)
        KC_LAUNCH(e, suffix_ranges_kernel<8>, pass_grid, 256, 0, e->d_pstart.as<uint32_t>(), split(e->d_isplit, q),
                  split(e->d_isplit, q + 1), n, e->d_pk.as<uint32_t>(), e->d_colptr.as<uint32_t>(),
                  e->d_col.as<uint32_t>(), fa, e->d_self.as<uint8_t>(), e->d_suf.as<uint2>(),
                  e->cfg.want_blosum ? e->d_sufss.as<uint8_t>() : nullptr, e->d_rowwork64.as<unsigned long long>(),
                  e->d_rowinl.as<uint32_t>(), e->d_rowmaxlen.as<uint32_t>(), split(e->d_psplit, q), &ds->work_total);
      else
        KC_LAUNCH(e, suffix_ranges_kernel<32>, pass_grid, 256, 0, e->d_pstart.as<uint32_t>(), split(e->d_isplit, q),
                  split(e->d_isplit, q + 1), n, e->d_pk.as<uint32_t>(), e->d_colptr.as<uint32_t>(),
                  e->d_col.as<uint32_t>(), fa, e->d_self.as<uint8_t>(), e->d_suf.as<uint2>(),
                  e->cfg.want_blosum ? e->d_sufss.as<uint8_t>() : nullptr, e->d_rowwork64.as<unsigned long long>(),
                  e->d_rowinl.as<uint32_t>(), e->d_rowmaxlen.as<uint32_t>(), split(e->d_psplit, q), &ds->work_total);
    }
    KC_LAUNCH(e, clamp_rowwork_kernel, (n + 255) / 256, 256, 0, e->d_rowwork64.as<unsigned long long>(), n,
              e->d_rowwork.as<uint32_t>());
  } else if (n) {
    KC_CUDA(e, cudaMemsetAsync(e->d_rowwork.p, 0, (uint64_t)n * 4, e->stream));
  }
  // work prefix for the shard partition
  if (n) {
    e->launches += exclusive_scan(WorkIn{e->d_rowwork.as<uint32_t>(), e->d_rowlen.as<uint32_t>()},
                                  U64ExclOutWithTail{e->d_workprefix.as<unsigned long long>(), n}, n, e->scan,
                                  e->stream);
  }
  KC_CUDA(e, cudaMemcpyAsync(&hs, ds, sizeof(hs), cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  KC_CUDA(e, cudaGetLastError());
  // Materialised multi-edge lists for the stream pair kernel (opt-in, kc_config.pair_lists): 4 bytes
  // (+1 with BLOSUM) per multi-edge.  Measured on synth_1m_k7: the fill costs 12 ms, the stream
  // kernel saves 3 ms over the gather kernels, so the default is the gather kernels.
  if (n && V && hs.work_total && e->cfg.pair_lists) {
    // budget: a quarter of the device memory (no cudaMemGetInfo here: it stalls for milliseconds)
    const unsigned long long need = hs.work_total * (4ull + (e->cfg.want_blosum ? 1ull : 0ull)) + (1ull << 20);
    if (need <= (unsigned long long)e->total_mem / 4) {
      KC_CUDA(e, e->d_plist.ensure((hs.work_total + 64) * 4));
      if (e->cfg.want_blosum) KC_CUDA(e, e->d_pss.ensure(hs.work_total + 64));
      e->launches += exclusive_scan(U64In{e->d_rowwork64.as<unsigned long long>()},
                                    U64ExclOutWithTail{e->d_rowbase.as<unsigned long long>(), n}, n, e->scan,
                                    e->stream);
      for (uint32_t q = 0; q < P; ++q) {
#define KC_PFILL(G)                                                                                          \
  KC_LAUNCH(e, products_fill_kernel<G>, pass_grid, 256, 0, e->d_pstart.as<uint32_t>(), split(e->d_isplit, q),   \
            split(e->d_isplit, q + 1), n, e->d_suf.as<uint2>(),                                                  \
            e->cfg.want_blosum ? e->d_sufss.as<uint8_t>() : nullptr, e->d_col.as<uint32_t>(),                  \
            e->d_rowbase.as<unsigned long long>(), e->d_rowwork64.as<unsigned long long>(), split(e->d_psplit, q), \
            e->d_plist.as<uint32_t>(), e->cfg.want_blosum ? e->d_pss.as<uint8_t>() : nullptr)
        if (narrow) KC_PFILL(8); else KC_PFILL(32);
#undef KC_PFILL
      }
      e->have_plist = true;
    }
  }
  mark(e, EV_I1);
  KC_CUDA(e, cudaGetLastError());
  e->istats.n_positions = n_positions;
  e->istats.n_incidences = n_incid;
  e->istats.n_distinct = hs.n_distinct;
  e->istats.n_repeated = V;
  e->v_local = V;
  e->istats.n_singleton = hs.n_distinct - V;
  e->istats.nnz = hs.nnz;
  e->multi_total = hs.multi_total;
  e->work_total = hs.work_total;
  if (stats) *stats = e->istats;
  e->have_index = true;
  return KC_OK;
}

int kc_get_distinct_kmers(kc_engine* e, uint32_t* out, uint64_t capacity) {
  if (!e || !out) return KC_EINVAL;
  if (!e->have_index) return fail(e, KC_EINVAL, "kc_build_index first");
  if (capacity < e->istats.n_distinct) return fail(e, KC_EINVAL, "capacity too small");
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (int rc = ensure_canonical(e)) return rc;
  const uint64_t W = e->n_words, D = e->istats.n_distinct;
  DBuf dict1, list;
  cudaError_t a = dict1.ensure(W * 8), b = list.ensure((D + 1) * 4);
  if (a != cudaSuccess || b != cudaSuccess) {
    dict1.release();
    list.release();
    return fail(e, KC_ENOMEM, "out of device memory");
  }
  e->launches += exclusive_scan(PopcIn<0>{e->d_seen.as<uint32_t>()},
                                DictOut<0>{e->d_seen.as<uint32_t>(), dict1.as<uint2>()}, W, e->scan, e->stream);
  KC_LAUNCH(e, expand_bitmap_kernel, blocks_for(W, 256, e->num_sm * 16), 256, 0, dict1.as<uint2>(), W,
            list.as<uint32_t>(), nullptr, e->cfg.k);
  cudaError_t rc = cudaMemcpyAsync(out, list.p, D * 4, cudaMemcpyDeviceToHost, e->stream);
  if (rc == cudaSuccess) rc = cudaStreamSynchronize(e->stream);
  dict1.release();
  list.release();
  KC_CUDA(e, rc);
  return KC_OK;
}

int kc_get_vocab(kc_engine* e, uint32_t* kmers_out, uint32_t* freq_out, uint64_t capacity) {
  if (!e) return KC_EINVAL;
  if (!e->have_index) return fail(e, KC_EINVAL, "kc_build_index first");
  const uint64_t V = e->istats.n_repeated;
  if (capacity < V) return fail(e, KC_EINVAL, "capacity too small");
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (int rc = ensure_canonical(e)) return rc;
  if (kmers_out && V) KC_CUDA(e, cudaMemcpyAsync(kmers_out, e->d_vocab.p, V * 4, cudaMemcpyDeviceToHost, e->stream));
  if (freq_out && V) KC_CUDA(e, cudaMemcpyAsync(freq_out, e->d_freq.p, V * 4, cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  return KC_OK;
}

int kc_get_protein_ids(kc_engine* e, uint64_t* row_offsets, uint32_t* ids_out, uint64_t capacity) {
  if (!e || !row_offsets) return KC_EINVAL;
  if (!e->have_index) return fail(e, KC_EINVAL, "kc_build_index first");
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (int rc = ensure_canonical(e)) return rc;
  const uint32_t n = (uint32_t)e->n;
  std::vector<uint32_t> rowlen(n);
  if (n) KC_CUDA(e, cudaMemcpyAsync(rowlen.data(), e->canon_rowlen(), (uint64_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  row_offsets[0] = 0;
  for (uint32_t p = 0; p < n; ++p) row_offsets[p + 1] = row_offsets[p] + rowlen[e->rank_of(p)];
  const uint64_t nnz = row_offsets[n];
  if (!ids_out) return KC_OK;
  if (capacity < nnz) return fail(e, KC_EINVAL, "capacity too small");
  if (!nnz) return KC_OK;
  DBuf d_ro, d_out;
  cudaError_t a = d_ro.ensure(((uint64_t)n + 1) * 8), b = d_out.ensure(nnz * 4);
  if (a != cudaSuccess || b != cudaSuccess) {
    d_ro.release();
    d_out.release();
    return fail(e, KC_ENOMEM, "out of device memory");
  }
  cudaMemcpyAsync(d_ro.p, row_offsets, ((uint64_t)n + 1) * 8, cudaMemcpyHostToDevice, e->stream);
  KC_LAUNCH(e, compact_rows_kernel, blocks_for(n, 8, e->num_sm * 8), 256, 0, e->d_pstart.as<uint32_t>(),
            e->canon_rowlen(), d_ro.as<unsigned long long>(), e->d_rank.as<uint32_t>(), n,
            e->d_pk.as<uint32_t>(), d_out.as<uint32_t>());
  cudaError_t rc = cudaMemcpyAsync(ids_out, d_out.p, nnz * 4, cudaMemcpyDeviceToHost, e->stream);
  if (rc == cudaSuccess) rc = cudaStreamSynchronize(e->stream);
  d_ro.release();
  d_out.release();
  KC_CUDA(e, rc);
  return KC_OK;
}

int kc_lookup_kmers(kc_engine* e, const uint32_t* kmers, uint64_t n, uint32_t* ids_out) {
  if (!e || (n && (!kmers || !ids_out))) return KC_EINVAL;
  if (!e->have_index) return fail(e, KC_EINVAL, "kc_build_index first");
  if (!n) return KC_OK;
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (int rc = ensure_canonical(e)) return rc;
  KC_CUDA(e, e->d_tmp.ensure(n * 8));
  uint32_t* d_in = e->d_tmp.as<uint32_t>();
  uint32_t* d_out = d_in + n;
  KC_CUDA(e, cudaMemcpyAsync(d_in, kmers, n * 4, cudaMemcpyHostToDevice, e->stream));
  KC_LAUNCH(e, lookup_kernel, blocks_for(n, 256, e->num_sm * 8), 256, 0, e->d_dict.as<uint2>(), d_in, n,
            e->universe, d_out);
  KC_CUDA(e, cudaMemcpyAsync(ids_out, d_out, n * 4, cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  return KC_OK;
}

int kc_get_pair_index(kc_engine* e, uint32_t* kmers_out, uint32_t* freq_out, uint8_t* self_out,
                      uint64_t capacity_vocab, uint64_t* row_offsets, uint32_t* ids_out, uint64_t capacity_ids) {
  if (!e) return KC_EINVAL;
  if (!e->have_index) return fail(e, KC_EINVAL, "kc_build_index first");
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (e->ishards > 1) return fail(e, KC_EINVAL, "kc_get_pair_index needs a whole index: kc_build_index, not a shard");
  const uint64_t V = e->istats.n_repeated;
  const uint32_t n = (uint32_t)e->n;
  if ((kmers_out || freq_out || self_out) && capacity_vocab < V) return fail(e, KC_EINVAL, "capacity too small");
  // The streaming build places ids by capacity (holes in the id space): this entry point presents them
  // densely, as a bijection onto [0, n_repeated) like boomphf's (src/main.rs:139-147).  map = rank among the used ids.
  DBuf d_map, d_vd, d_fd, d_sd;
  const bool sparse = e->streamed;
  auto free_tmp = [&]() {
    d_map.release();
    d_vd.release();
    d_fd.release();
    d_sd.release();
  };
  if (sparse) {
    const uint64_t slots = e->v_local;
    cudaError_t a = d_map.ensure((slots + 2) * 4), b2 = d_vd.ensure((V + 1) * 4), c = d_fd.ensure((V + 1) * 4),
                d = d_sd.ensure(V + 16);
    if (a != cudaSuccess || b2 != cudaSuccess || c != cudaSuccess || d != cudaSuccess) {
      free_tmp();
      return fail(e, KC_ENOMEM, "out of device memory");
    }
    if (int rc = ensure_scan(e, slots + 1)) {
      free_tmp();
      return rc;
    }
    cudaMemsetAsync(d_map.p, 0, (slots + 2) * 4, e->stream);
    if (n)
      KC_LAUNCH(e, sx_mark_ids_kernel, blocks_for(n, 8, e->num_sm * 8), 256, 0, e->pair_rowptr(),
                e->d_rowlen.as<uint32_t>(), n, e->pair_ids(), d_map.as<uint32_t>());
    e->launches += exclusive_scan(U32In{d_map.as<uint32_t>()}, SxExclOutTail{d_map.as<uint32_t>(), slots + 1}, slots + 1,
                                  e->scan, e->stream);
    KC_LAUNCH(e, sx_gather_vocab_kernel, blocks_for(slots, 256, e->num_sm * 8), 256, 0, d_map.as<uint32_t>(), slots,
              e->d_vocab_h.as<uint32_t>(), e->d_freq_h.as<uint32_t>(), e->d_self_h.as<uint8_t>(), d_vd.as<uint32_t>(),
              d_fd.as<uint32_t>(), d_sd.as<uint8_t>());
  }
  const void* vb = sparse ? d_vd.p : (e->bucketed ? e->d_vocab_h.p : e->d_vocab.p);
  const void* fb = sparse ? d_fd.p : (e->bucketed ? e->d_freq_h.p : e->d_freq.p);
  const void* sb = sparse ? d_sd.p : (const void*)e->pair_self();
  cudaError_t crc = cudaSuccess;
  if (kmers_out && V) crc = cudaMemcpyAsync(kmers_out, vb, V * 4, cudaMemcpyDeviceToHost, e->stream);
  if (crc == cudaSuccess && freq_out && V) crc = cudaMemcpyAsync(freq_out, fb, V * 4, cudaMemcpyDeviceToHost, e->stream);
  if (crc == cudaSuccess && self_out && V) crc = cudaMemcpyAsync(self_out, sb, V, cudaMemcpyDeviceToHost, e->stream);
  if (crc == cudaSuccess) crc = cudaStreamSynchronize(e->stream);
  if (crc != cudaSuccess || !row_offsets) {
    free_tmp();
    KC_CUDA(e, crc);
    return KC_OK;
  }
  std::vector<uint32_t> rowlen(n);
  if (n) crc = cudaMemcpyAsync(rowlen.data(), e->d_rowlen.p, (uint64_t)n * 4, cudaMemcpyDeviceToHost, e->stream);
  if (crc == cudaSuccess) crc = cudaStreamSynchronize(e->stream);
  if (crc != cudaSuccess) {
    free_tmp();
    KC_CUDA(e, crc);
  }
  row_offsets[0] = 0;
  for (uint32_t p = 0; p < n; ++p) row_offsets[p + 1] = row_offsets[p] + rowlen[e->rank_of(p)];
  const uint64_t nnz = row_offsets[n];
  if (!ids_out || !nnz) {
    free_tmp();
    return KC_OK;
  }
  if (capacity_ids < nnz) {
    free_tmp();
    return fail(e, KC_EINVAL, "capacity too small");
  }
  DBuf d_ro, d_out;
  cudaError_t a = d_ro.ensure(((uint64_t)n + 1) * 8), b = d_out.ensure(nnz * 4);
  if (a != cudaSuccess || b != cudaSuccess) {
    d_ro.release();
    d_out.release();
    free_tmp();
    return fail(e, KC_ENOMEM, "out of device memory");
  }
  cudaMemcpyAsync(d_ro.p, row_offsets, ((uint64_t)n + 1) * 8, cudaMemcpyHostToDevice, e->stream);
  KC_LAUNCH(e, compact_rows_kernel, blocks_for(n, 8, e->num_sm * 8), 256, 0, e->pair_rowptr(),
            e->d_rowlen.as<uint32_t>(), d_ro.as<unsigned long long>(), e->d_rank.as<uint32_t>(), n, e->pair_ids(),
            d_out.as<uint32_t>());
  if (sparse)
    KC_LAUNCH(e, sx_remap_ids_kernel, blocks_for(nnz, 256, e->num_sm * 8), 256, 0, d_map.as<uint32_t>(),
              d_out.as<uint32_t>(), nnz);
  cudaError_t rc = cudaMemcpyAsync(ids_out, d_out.p, nnz * 4, cudaMemcpyDeviceToHost, e->stream);
  if (rc == cudaSuccess) rc = cudaStreamSynchronize(e->stream);
  d_ro.release();
  d_out.release();
  free_tmp();
  KC_CUDA(e, rc);
  return KC_OK;
}

int kc_score_pairs(kc_engine* e, kc_pair_stats* stats) { return kc_score_pairs_shard(e, 0, 1, stats); }

int kc_score_pairs_shard(kc_engine* e, uint32_t shard, uint32_t n_shards, kc_pair_stats* stats) {
  if (!e || n_shards == 0 || shard >= n_shards) return KC_EINVAL;
  if (!e->have_index) return fail(e, KC_EINVAL, "kc_build_index first");
  if (e->ishards > 1 && (shard != e->ishard || n_shards != e->ishards))
    return fail(e, KC_EINVAL, "the index was built for another shard (kc_build_index_shard)");
  KC_CUDA(e, cudaSetDevice(e->dev));
  const uint32_t n = (uint32_t)e->n;
  DeviceScalars* ds = e->ds;
  e->have_pairs = false;
  e->n_edges = 0;
  kc_pair_stats ps{};
  ps.n_multi_edges = e->multi_total;
  if (n == 0 || e->v_local == 0) {
    e->pstats = ps;
    if (stats) *stats = ps;
    e->have_pairs = true;
    mark(e, EV_P0);
    mark(e, EV_PK0);
    mark(e, EV_PK1);
    mark(e, EV_P1);
    mark(e, EV_E1);
    return KC_OK;
  }
  KC_CUDA(e, e->d_rowbin.ensure((uint64_t)n + 64));
  KC_CUDA(e, e->d_rowsafe.ensure((uint64_t)n + 64));
  KC_CUDA(e, e->d_rowlogh.ensure((uint64_t)n + 64));
  if (e->edge_cap == 0) {
    e->edge_cap = e->cfg.max_edges ? e->cfg.max_edges : (1ull << 22);
    KC_CUDA(e, e->d_edges.ensure(e->edge_cap * 16));
  }
  // dense accumulators: u16 unless some row could exceed it
  const uint32_t max_rowlen = e->max_plen;
  const bool wide = max_rowlen >= 65535u;
  const size_t dense_budget = std::min<size_t>(e->smem_optin ? e->smem_optin : 48 * 1024, 200 * 1024);
  const uint32_t dense_cap_cols = (uint32_t)((dense_budget - 64) / (wide ? 4 : 2)) & ~7u;
  const uint32_t dense_cols = std::min<uint32_t>(dense_cap_cols, (n + 7u) & ~7u);
  const size_t dense_smem = (size_t)(wide ? dense_cols : (dense_cols + 1) / 2) * 4 + 16;
  // packed hash slots: protein rank in the high bits (2^key_bits > n so no key is all ones)
  uint32_t key_bits = 1;
  while ((1ull << key_bits) <= (uint64_t)n) ++key_bits;
  const uint32_t count_bits = 32 - key_bits;

  mark(e, EV_P0);
  for (int attempt = 0;; ++attempt) {
    KC_CUDA(e, cudaMemsetAsync(&ds->edge_cursor, 0,
                               sizeof(DeviceScalars) - offsetof(DeviceScalars, edge_cursor), e->stream));
    if (e->ishards <= 1)  // (sharded index: the rows were fixed when the index was built, RowOwner below)
      KC_LAUNCH(e, shard_bounds_kernel, 1, 32, 0, e->d_workprefix.as<unsigned long long>(), n, shard, n_shards,
                ds->shard_rows);
    KC_LAUNCH(e, classify_rows_kernel, (n + 255) / 256, 256, 0, e->d_rowwork.as<uint32_t>(),
              e->pair_rowlen(), e->d_rowinl.as<uint32_t>(), e->d_rowmaxlen.as<uint32_t>(),
              e->cfg.cross_class_only ? e->d_first_after.as<uint32_t>() : nullptr, n,
              ds->shard_rows, dense_cols, count_bits, e->d_rowbin.as<uint8_t>(), e->d_rowsafe.as<uint8_t>(),
              e->d_rowlogh.as<uint8_t>(), ds->bin_counts, e->owner(),
              // trust the partner bound only when the tiles took most of the multi-edges (input with locality)
              !e->bucketed ? 0 : ((double)e->work_total <= 0.6 * (double)e->multi_total ? 1 : 2),
              e->cfg.want_blosum != 0 && !e->have_plist);
    EdgeSink sink{e->d_edges.as<uint4>(), &ds->edge_cursor, e->edge_cap, e->cfg.threshold,
                  e->cfg.want_blosum ? kUnscored : 0u};
    mark(e, EV_PK0);
    int rc;
    if (e->bucketed) {  // the pairs inside one 64-row bin: dense shared-memory tiles over the run records
      const uint32_t n_bins = (n + kBinRows - 1) >> kBinRowsLog;
      const uint32_t tgrid = std::min<uint32_t>(n_bins, (uint32_t)e->num_sm * 4);
      const uint32_t* fa = e->cfg.cross_class_only ? e->d_first_after.as<uint32_t>() : nullptr;
      const uint32_t* bounds = e->ishards > 1 ? nullptr : ds->shard_rows;
#define KC_TILES(SC, CR)                                                                                        \
  do {                                                                                                          \
    KC_CUDA(e, cudaFuncSetAttribute((pairs_tile_kernel<SC, CR>), cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                    (int)kTileSmemBytes));                                                      \
    KC_LAUNCH(e, (pairs_tile_kernel<SC, CR>), tgrid, kTileThreads, kTileSmemBytes, e->d_runs.as<uint4>(),       \
              e->d_rowcap.as<uint32_t>(), e->d_run_cnt.as<uint32_t>(), n, n_bins, fa, bounds, e->owner(), sink, \
              &ds->pc);                                                                                         \
  } while (0)
      if (e->cfg.want_blosum) {
        if (fa) KC_TILES(true, true); else KC_TILES(true, false);
      } else {
        if (fa) KC_TILES(false, true); else KC_TILES(false, false);
      }
#undef KC_TILES
    }
    if (e->have_plist) {
#define KC_STREAM(SC)                                                                                         \
  do {                                                                                                        \
    constexpr size_t smem = (size_t)kStreamWarps * stream_warp_words<SC>() * 4;                               \
    KC_CUDA(e, cudaFuncSetAttribute(pairs_stream_kernel<SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int per_sm = 1;                                                                                           \
    KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pairs_stream_kernel<SC>, kStreamWarps * 32, smem)); \
    if (per_sm < 1) per_sm = 1;                                                                               \
    KC_LAUNCH(e, pairs_stream_kernel<SC>, (uint32_t)(e->num_sm * per_sm), kStreamWarps * 32, smem,            \
              e->d_rowbase.as<unsigned long long>(), e->d_rowwork.as<uint32_t>(), e->d_plist.as<uint32_t>(),  \
              SC ? e->d_pss.as<uint8_t>() : nullptr, e->d_rowbin.as<uint8_t>(), e->d_rowsafe.as<uint8_t>(), n, \
              count_bits, &ds->row_cursor[kBinMain], &ds->n_overflow, ds->bin_counts, sink, &ds->pc);         \
  } while (0)
      if (e->cfg.want_blosum) KC_STREAM(true); else KC_STREAM(false);
#undef KC_STREAM
    } else     if (e->cfg.want_blosum) {
#define KC_SCORED(LOGH, CAPV, BINV, WARPSV)                                                                       \
  do {                                                                                                            \
    auto kern = pairs_main_scored_kernel<LOGH, CAPV, BINV, WARPSV>;                                               \
    constexpr size_t smem = (size_t)WARPSV * scored_warp_words<LOGH, CAPV>() * 4;                                 \
    KC_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    int per_sm = 1;                                                                                               \
    KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPSV * 32, smem));                  \
    if (per_sm < 1) per_sm = 1;                                                                                   \
    KC_LAUNCH(e, kern, (uint32_t)(e->num_sm * per_sm), WARPSV * 32, smem, e->pair_rowptr(), e->pair_rowlen(),     \
              e->d_suf.as<uint2>(), e->d_sufss.as<uint8_t>(), e->d_col.as<uint32_t>(), e->d_rowbin.as<uint8_t>(), \
              e->d_rowsafe.as<uint8_t>(), n, &ds->row_cursor[BINV], &ds->n_overflow, ds->bin_counts, sink, &ds->pc); \
  } while (0)
      KC_SCORED(kMainSLogH, kMainSCap, kBinMainS, 7);
      KC_SCORED(kMainLogHMax, kMainCap, kBinMain, 5);
#undef KC_SCORED
    } else {
      constexpr size_t smem = (size_t)kMainWarps * ((1u << kMainLogHMax) + 2 * kIdxPerWarp + kMainCap / 2 + 4) * 4;
      KC_CUDA(e, cudaFuncSetAttribute(pairs_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 1;
      KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pairs_main_kernel, kMainWarps * 32, smem));
      if (per_sm < 1) per_sm = 1;
      KC_LAUNCH(e, pairs_main_kernel, (uint32_t)(e->num_sm * per_sm), kMainWarps * 32, smem,
                e->pair_rowptr(), e->pair_rowlen(), e->d_suf.as<uint2>(),
                e->d_col.as<uint32_t>(), e->d_rowbin.as<uint8_t>(), e->d_rowsafe.as<uint8_t>(),
                e->d_rowlogh.as<uint8_t>(), n, count_bits, &ds->row_cursor[kBinMain], &ds->n_overflow,
                ds->bin_counts, sink, &ds->pc);
    }
    if ((rc = launch_packed<8, 1, 4>(e, kBinPack8, count_bits, sink))) return rc;
    if ((rc = launch_packed<9, 1, 4>(e, kBinPack9, count_bits, sink))) return rc;
    if ((rc = launch_packed<10, 1, 4>(e, kBinPack10, count_bits, sink))) return rc;
    if ((rc = launch_packed<11, 1, 4>(e, kBinPack11, count_bits, sink))) return rc;
    if ((rc = launch_packed<12, 4, 4>(e, kBinPack12, count_bits, sink))) return rc;
    if ((rc = launch_packed<13, 8, 8>(e, kBinPack13, count_bits, sink))) return rc;
    if ((rc = launch_packed<14, 8, 8>(e, kBinPack14, count_bits, sink))) return rc;
    if ((rc = launch_hash<14, 8, 8>(e, kBinWide, sink))) return rc;
    {
      int per_sm = 1;
      if (wide) {
        KC_CUDA(e, cudaFuncSetAttribute(pairs_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)dense_smem));
        KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pairs_dense_kernel<true>, kDenseThreads, dense_smem));
      } else {
        KC_CUDA(e, cudaFuncSetAttribute(pairs_dense_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)dense_smem));
        KC_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pairs_dense_kernel<false>, kDenseThreads, dense_smem));
      }
      if (per_sm < 1) per_sm = 1;
      const uint32_t grid = (uint32_t)(e->num_sm * per_sm);
      const uint32_t* fa = e->cfg.cross_class_only ? e->d_first_after.as<uint32_t>() : nullptr;
      if (wide)
        KC_LAUNCH(e, pairs_dense_kernel<true>, grid, kDenseThreads, dense_smem, e->pair_rowptr(),
                  e->pair_rowlen(), e->d_suf.as<uint2>(), e->d_col.as<uint32_t>(), fa,
                  e->d_rowbin.as<uint8_t>(), n, dense_cols, &ds->row_cursor[kBinDense], ds->bin_counts, sink, &ds->pc);
      else
        KC_LAUNCH(e, pairs_dense_kernel<false>, grid, kDenseThreads, dense_smem, e->pair_rowptr(),
                  e->pair_rowlen(), e->d_suf.as<uint2>(), e->d_col.as<uint32_t>(), fa,
                  e->d_rowbin.as<uint8_t>(), n, dense_cols, &ds->row_cursor[kBinDense], ds->bin_counts, sink, &ds->pc);
    }
    mark(e, EV_PK1);
    DeviceScalars hs{};
    KC_CUDA(e, cudaMemcpyAsync(&hs, ds, sizeof(hs), cudaMemcpyDeviceToHost, e->stream));
    KC_CUDA(e, cudaStreamSynchronize(e->stream));
    KC_CUDA(e, cudaGetLastError());
    if (hs.n_overflow) {  // rows whose optimistic table overflowed: score them with safely sized tables
      if ((rc = launch_packed<8, 1, 4>(e, kBinRetry + kBinPack8, count_bits, sink))) return rc;
      if ((rc = launch_packed<9, 1, 4>(e, kBinRetry + kBinPack9, count_bits, sink))) return rc;
      if ((rc = launch_packed<10, 1, 4>(e, kBinRetry + kBinPack10, count_bits, sink))) return rc;
      if ((rc = launch_packed<11, 1, 4>(e, kBinRetry + kBinPack11, count_bits, sink))) return rc;
      if ((rc = launch_packed<12, 4, 4>(e, kBinRetry + kBinPack12, count_bits, sink))) return rc;
      if ((rc = launch_packed<13, 8, 8>(e, kBinRetry + kBinPack13, count_bits, sink))) return rc;
      if ((rc = launch_packed<14, 8, 8>(e, kBinRetry + kBinPack14, count_bits, sink))) return rc;
      mark(e, EV_PK1);
      KC_CUDA(e, cudaMemcpyAsync(&hs, ds, sizeof(hs), cudaMemcpyDeviceToHost, e->stream));
      KC_CUDA(e, cudaStreamSynchronize(e->stream));
      KC_CUDA(e, cudaGetLastError());
    }
    if (hs.edge_cursor > e->edge_cap) {  // grow to the exact need and score again
      if (attempt >= 2) return fail(e, KC_ECUDA, "edge buffer did not converge");
      e->edge_cap = hs.edge_cursor + (hs.edge_cursor >> 4) + 1024;
      e->d_edges.release();
      KC_CUDA(e, e->d_edges.ensure(e->edge_cap * 16));
      ps.n_retries++;
      continue;
    }
    ps.n_pairs_kept = hs.pc.n_pairs;
    ps.n_edges_out = hs.pc.n_edges;
    ps.sum_count_out = hs.pc.sum_count;
    ps.n_rows = e->ishards > 1 ? e->n_own_rows : hs.shard_rows[1] - hs.shard_rows[0];
    ps.n_rows_rescored = hs.n_overflow;
    e->n_edges = hs.edge_cursor;
    ps.n_multi_edges_kept = hs.pc.n_multi;
    break;
  }
  mark(e, EV_P1);
  const uint64_t ne = e->n_edges;
  // K9 + canonical order
  if (ne) {
    KC_CUDA(e, e->d_keys_a.ensure(ne * 8));
    KC_CUDA(e, e->d_keys_b.ensure(ne * 8));
    KC_CUDA(e, e->d_vals_a.ensure(ne * 8));
    KC_CUDA(e, e->d_vals_b.ensure(ne * 8));
    const uint32_t nb = (uint32_t)((ne + kRsTile - 1) / kRsTile);
    KC_CUDA(e, e->d_hist.ensure((size_t)nb * 256 * 4));
    int rc = ensure_scan(e, (uint64_t)nb * 256);
    if (rc) return rc;
    KC_CUDA(e, e->d_edges_sorted.ensure(ne * 16));
    KC_LAUNCH(e, finalize_edges_kernel, blocks_for(ne, 256, e->num_sm * 8), 256, 0, e->d_edges.as<uint4>(), ne,
              e->d_orig.as<uint32_t>(), e->d_keys_a.as<unsigned long long>(), e->d_vals_a.as<unsigned long long>());
    int nbits = 1;
    while ((1ull << nbits) < (uint64_t)n) ++nbits;
    int passes[8], np = 0;
    for (int s = 0; s < nbits; s += 8) passes[np++] = s;
    for (int s = 0; s < nbits; s += 8) passes[np++] = 32 + s;
    int in_b = 0;
    e->launches += radix_sort_pairs(e->d_keys_a.as<unsigned long long>(), e->d_vals_a.as<unsigned long long>(),
                                    e->d_keys_b.as<unsigned long long>(), e->d_vals_b.as<unsigned long long>(), ne,
                                    passes, np, e->d_hist.as<uint32_t>(), e->scan, e->stream, &in_b);
    if (e->cfg.want_blosum) {
      // hash size from the longest possible row (a row has at most plen - k + 1 ids)
      const bool small_rows = 10ull * max_rowlen <= 7ull * 1024;
      unsigned long long* skeys = (in_b ? e->d_keys_b : e->d_keys_a).as<unsigned long long>();
      unsigned long long* svals = (in_b ? e->d_vals_b : e->d_vals_a).as<unsigned long long>();
      if (small_rows)
        KC_LAUNCH(e, (edge_blosum_kernel<1024, 22, 4>), blocks_for((ne + 31) / 32, 4, e->num_sm * 10), 128, 0, skeys, svals,
                  ne, e->d_rank.as<uint32_t>(), e->pair_rowptr(), e->d_rowlen.as<uint32_t>(), e->pair_ids(),
                  e->pair_self());
      else
        KC_LAUNCH(e, (edge_blosum_kernel<4096, 20, 2>), blocks_for((ne + 31) / 32, 2, e->num_sm * 5), 64, 0, skeys, svals,
                  ne, e->d_rank.as<uint32_t>(), e->pair_rowptr(), e->d_rowlen.as<uint32_t>(), e->pair_ids(),
                  e->pair_self());
    }
    KC_LAUNCH(e, assemble_edges_kernel, blocks_for(ne, 256, e->num_sm * 8), 256, 0,
              (in_b ? e->d_keys_b : e->d_keys_a).as<unsigned long long>(),
              (in_b ? e->d_vals_b : e->d_vals_a).as<unsigned long long>(), ne, e->d_edges_sorted.as<uint4>());
  }
  mark(e, EV_E1);
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  KC_CUDA(e, cudaGetLastError());
  e->pstats = ps;
  e->have_pairs = true;
  if (stats) *stats = e->pstats;
  return KC_OK;
}

int kc_get_edges(kc_engine* e, kc_edge* out, uint64_t capacity) {
  if (!e) return KC_EINVAL;
  if (!e->have_pairs) return fail(e, KC_EINVAL, "kc_score_pairs first");
  if (capacity < e->n_edges) return fail(e, KC_EINVAL, "capacity too small");
  if (!e->n_edges) return KC_OK;
  if (!out) return KC_EINVAL;
  KC_CUDA(e, cudaSetDevice(e->dev));
  mark(e, EV_D0);
  KC_CUDA(e, cudaMemcpyAsync(out, e->d_edges_sorted.p, e->n_edges * 16, cudaMemcpyDeviceToHost, e->stream));
  mark(e, EV_D1);
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  return KC_OK;
}

int kc_get_edges_device(kc_engine* e, const kc_edge** d_edges_out, uint64_t* n_edges_out) {
  if (!e || !d_edges_out || !n_edges_out) return KC_EINVAL;
  if (!e->have_pairs) return fail(e, KC_EINVAL, "kc_score_pairs first");
  *d_edges_out = e->n_edges ? reinterpret_cast<const kc_edge*>(e->d_edges_sorted.p) : nullptr;
  *n_edges_out = e->n_edges;
  return KC_OK;
}

int kc_get_edge_kmers(kc_engine* e, uint64_t edge_index, uint32_t* kmers_out, uint64_t capacity) {
  if (!e || !kmers_out) return KC_EINVAL;
  if (!e->have_pairs) return fail(e, KC_EINVAL, "kc_score_pairs first");
  if (edge_index >= e->n_edges) return fail(e, KC_EINVAL, "edge index out of range");
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (int rc = ensure_canonical(e)) return rc;
  uint4 ed;
  KC_CUDA(e, cudaMemcpyAsync(&ed, e->d_edges_sorted.as<uint4>() + edge_index, 16, cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  if (capacity < ed.z) return fail(e, KC_EINVAL, "capacity too small");
  KC_CUDA(e, e->d_tmp.ensure((uint64_t)ed.z * 4 + 16));
  KC_LAUNCH(e, shared_kmers_kernel, 1, 32, 0, e->rank_of(ed.x), e->rank_of(ed.y), e->d_pstart.as<uint32_t>(),
            e->canon_rowlen(), e->d_pk.as<uint32_t>(), e->d_vocab.as<uint32_t>(), e->d_tmp.as<uint32_t>(),
            ed.z, &e->ds->n_shared);
  KC_CUDA(e, cudaMemcpyAsync(kmers_out, e->d_tmp.p, (uint64_t)ed.z * 4, cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  return KC_OK;
}

int kc_get_timings(kc_engine* e, kc_timings* out) {
  if (!e || !out) return KC_EINVAL;
  cudaSetDevice(e->dev);
  cudaStreamSynchronize(e->stream);
  std::memset(out, 0, sizeof(*out));
  out->h2d_ms = elapsed(e, EV_H2D0, EV_H2D1);
  out->extract_ms = elapsed(e, EV_X0, EV_X1);
  out->index_ms = elapsed(e, EV_I0, EV_I1);
  out->pairs_ms = elapsed(e, EV_P0, EV_P1);
  out->edges_ms = elapsed(e, EV_P1, EV_E1);
  out->d2h_ms = elapsed(e, EV_D0, EV_D1);
  out->pair_kernel_ms = elapsed(e, EV_PK0, EV_PK1);
  out->census_kernel_ms = elapsed(e, EV_IC0, EV_IC1);
  out->kernel_launches = e->launches;
  out->index_records = e->index_records;
  out->index_mid_buckets = e->sx_mid_last;
  out->index_huge_buckets = e->sx_huge_last;
  return KC_OK;
}

int kc_reset_timings(kc_engine* e) {
  if (!e) return KC_EINVAL;
  e->launches = 0;
  for (int i = 0; i < EV_COUNT; ++i) e->ev_set[i] = false;
  return KC_OK;
}

int kc_bitset_pair_counts(kc_engine* e, const uint32_t* rows, uint32_t n_rows, uint32_t* counts_out) {
  if (!e || (n_rows && (!rows || !counts_out))) return KC_EINVAL;
  if (!e->have_index) return fail(e, KC_EINVAL, "kc_build_index first");
  if (!n_rows) return KC_OK;
  for (uint32_t i = 0; i < n_rows; ++i)
    if (rows[i] >= e->n) return fail(e, KC_EINVAL, "row out of range");
  KC_CUDA(e, cudaSetDevice(e->dev));
  if (int rc = ensure_canonical(e)) return rc;
  const uint32_t words = (uint32_t)((e->istats.n_repeated + 31) / 32) + 1;
  std::vector<uint32_t> ranks(n_rows);
  for (uint32_t i = 0; i < n_rows; ++i) ranks[i] = e->rank_of(rows[i]);
  DBuf d_rows, d_bits, d_counts;
  cudaError_t a = d_rows.ensure((size_t)n_rows * 4), b = d_bits.ensure((size_t)n_rows * words * 4),
              c = d_counts.ensure((size_t)n_rows * n_rows * 4);
  if (a != cudaSuccess || b != cudaSuccess || c != cudaSuccess) {
    d_rows.release();
    d_bits.release();
    d_counts.release();
    return fail(e, KC_ENOMEM, "out of device memory");
  }
  cudaMemcpyAsync(d_rows.p, ranks.data(), (size_t)n_rows * 4, cudaMemcpyHostToDevice, e->stream);
  cudaMemsetAsync(d_bits.p, 0, (size_t)n_rows * words * 4, e->stream);
  KC_LAUNCH(e, bitset_fill_kernel, blocks_for(n_rows, 8, e->num_sm * 8), 256, 0, d_rows.as<uint32_t>(), n_rows,
            e->d_pstart.as<uint32_t>(), e->canon_rowlen(), e->d_pk.as<uint32_t>(), words,
            d_bits.as<uint32_t>());
  KC_LAUNCH(e, bitset_pairs_kernel, dim3((n_rows + kBitCols - 1) / kBitCols, (n_rows + kBitRows - 1) / kBitRows),
            kBitWarps * 32, 0, d_bits.as<uint32_t>(), n_rows, words, d_counts.as<uint32_t>());
  e->bitset_word_ops = (double)n_rows * n_rows * words;
  cudaError_t rc = cudaMemcpyAsync(counts_out, d_counts.p, (size_t)n_rows * n_rows * 4, cudaMemcpyDeviceToHost,
                                   e->stream);
  if (rc == cudaSuccess) rc = cudaStreamSynchronize(e->stream);
  d_rows.release();
  d_bits.release();
  d_counts.release();
  KC_CUDA(e, rc);
  return KC_OK;
}

// POPC issue-rate microbenchmark: the measured denominator of the bitset path's roofline.
int kc_popc_microbench(kc_engine* e, uint32_t iters, double* gpopc_per_s) {
  if (!e || !gpopc_per_s || iters == 0) return KC_EINVAL;
  KC_CUDA(e, cudaSetDevice(e->dev));
  KC_CUDA(e, e->d_tmp.ensure(64));
  const uint32_t grid = (uint32_t)e->num_sm * 8u;
  cudaEvent_t a, b;
  KC_CUDA(e, cudaEventCreate(&a));
  KC_CUDA(e, cudaEventCreate(&b));
  KC_LAUNCH(e, popc_microbench_kernel, grid, 256, 0, iters / 8 + 1, 1u, e->d_tmp.as<uint32_t>());  // warm-up
  cudaEventRecord(a, e->stream);
  KC_LAUNCH(e, popc_microbench_kernel, grid, 256, 0, iters, 2u, e->d_tmp.as<uint32_t>());
  cudaEventRecord(b, e->stream);
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *gpopc_per_s = ms > 0.f ? (double)grid * 256.0 * 8.0 * iters / (ms * 1e-3) / 1e9 : 0.0;
  return KC_OK;
}

// ---- multi-GPU entry points (dist.cuh) ---------------------------------------------------------
int kc_comm_unique_id(uint8_t* id) {
  if (!id) return KC_EINVAL;
  NcclApi& nc = nccl_api();
  if (!nc.ok()) return KC_ENODEVICE;
  ncclUniqueId u;
  if (nc.GetUniqueId(&u) != ncclSuccess) return KC_ECUDA;
  static_assert(sizeof(u) == KC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  std::memcpy(id, &u, sizeof(u));
  return KC_OK;
}

int kc_comm_init(kc_engine* e, const uint8_t* id, int rank, int world) {
  if (!e || !id || world < 1 || rank < 0 || rank >= world) return KC_EINVAL;
  if (world > 255) return fail(e, KC_EINVAL, "at most 255 ranks (the row blocks' owner is one byte)");
  KC_CUDA(e, cudaSetDevice(e->dev));
  NcclApi& nc = nccl_api();
  if (!nc.ok()) return fail(e, KC_ENODEVICE, "NCCL is not available: " + nc.err);
  if (e->comm) {
    nc.CommDestroy(e->comm);
    e->comm = nullptr;
  }
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  const ncclResult_t nr = nc.CommInitRank(&e->comm, world, u, rank);
  if (nr != ncclSuccess) {
    e->comm = nullptr;
    return fail(e, KC_ECUDA, std::string("ncclCommInitRank: ") + nc.GetErrorString(nr));
  }
  e->crank = rank;
  e->cworld = world;
  KC_CUDA(e, e->d_comm.ensure(4096));
  return KC_OK;
}

int kc_comm_info(kc_engine* e, int* rank, int* world) {
  if (!e) return KC_EINVAL;
  if (rank) *rank = e->comm ? e->crank : 0;
  if (world) *world = e->comm ? e->cworld : 1;
  return KC_OK;
}

int kc_set_proteins_dist(kc_engine* e, const uint8_t* residues, const uint64_t* offsets, const uint32_t* class_id,
                         uint64_t n) {
  if (!e) return KC_EINVAL;
  return set_proteins_host(e, residues, offsets, class_id, n, e->comm != nullptr && e->cworld > 1);
}

// sum of `count` u64 values over the ranks, through the engine's device scratch
static int allreduce_u64(kc_engine* e, unsigned long long* vals, int count) {
  if (!e->comm || e->cworld <= 1) return KC_OK;
  NcclApi& nc = nccl_api();
  unsigned long long* d = e->d_comm.as<unsigned long long>();
  KC_CUDA(e, cudaMemcpyAsync(d, vals, (size_t)count * 8, cudaMemcpyHostToDevice, e->stream));
  const ncclResult_t nr = nc.AllReduce(d, d, (size_t)count, ncclUint64, ncclSum, e->comm, e->stream);
  if (nr != ncclSuccess) return fail(e, KC_ECUDA, std::string("ncclAllReduce: ") + nc.GetErrorString(nr));
  KC_CUDA(e, cudaMemcpyAsync(vals, d, (size_t)count * 8, cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  return KC_OK;
}

int kc_build_index_dist(kc_engine* e, kc_index_stats* stats) {
  if (!e) return KC_EINVAL;
  kc_index_stats st{};
  int rc = kc_build_index_shard(e, (uint32_t)(e->comm ? e->crank : 0), (uint32_t)(e->comm ? e->cworld : 1), &st);
  // Every rank must take the same branch below (a collective follows): whether the build sharded the index
  // is a function of the configuration only (kc_config.index_build, subsampling), never of the data.
  unsigned long long v[8] = {st.n_positions, st.n_incidences, st.n_distinct, st.n_singleton, st.n_repeated, st.nnz,
                             (unsigned long long)(rc != KC_OK), (unsigned long long)(e->ishards > 1)};
  if (e->comm && e->cworld > 1) {
    const bool sharded = e->ishards > 1 || rc != KC_OK;
    if (!sharded) {  // a whole index on every rank: only the error flag travels
      for (int i = 0; i < 6; ++i) v[i] = 0;
    }
    const kc_index_stats mine = st;
    const int rc2 = allreduce_u64(e, v, 8);
    if (rc2) return rc2;
    if (v[6]) return rc != KC_OK ? rc : fail(e, KC_ECUDA, "kc_build_index_dist failed on another rank");
    if (sharded) {
      st.n_positions = v[0], st.n_incidences = v[1], st.n_distinct = v[2];
      st.n_singleton = v[3], st.n_repeated = v[4], st.nnz = v[5];
    } else {
      st = mine;
    }
  } else if (rc != KC_OK) {
    return rc;
  }
  if (stats) *stats = st;
  return KC_OK;
}

int kc_score_pairs_dist(kc_engine* e, kc_pair_stats* stats) {
  if (!e) return KC_EINVAL;
  kc_pair_stats ps{};
  const int rc = kc_score_pairs_shard(e, (uint32_t)(e->comm ? e->crank : 0), (uint32_t)(e->comm ? e->cworld : 1), &ps);
  if (e->comm && e->cworld > 1) {
    unsigned long long v[8] = {ps.n_multi_edges_kept, ps.n_pairs_kept, ps.n_edges_out, ps.sum_count_out, ps.n_rows,
                               e->ishards > 1 ? ps.n_multi_edges : 0ull, (unsigned long long)(rc != KC_OK),
                               ps.n_rows_rescored};
    const int rc2 = allreduce_u64(e, v, 8);
    if (rc2) return rc2;
    if (v[6]) return rc != KC_OK ? rc : fail(e, KC_ECUDA, "kc_score_pairs_dist failed on another rank");
    // (the engine keeps the rank-local numbers: kc_get_edges / kc_gather_edges size by them)
    ps.n_multi_edges_kept = v[0], ps.n_pairs_kept = v[1], ps.n_edges_out = v[2], ps.sum_count_out = v[3];
    ps.n_rows = v[4];
    if (e->ishards > 1) ps.n_multi_edges = v[5];
    ps.n_rows_rescored = v[7];
  } else if (rc != KC_OK) {
    return rc;
  }
  if (stats) *stats = ps;
  return KC_OK;
}

// The edge lists of all ranks as one list sorted by (a, b).  shared == 0: device-to-device over NVLink into a
// buffer on rank 0, one D2H copy into rank 0's `out` (the other ranks pass out = NULL).  shared != 0: `out` is
// ONE host buffer mapped by every rank (one process with a thread per GPU, or POSIX shared memory across
// processes): every rank copies its own runs straight to their final place over its own PCIe link.
static int gather_edges_impl(kc_engine* e, kc_edge* out, uint64_t capacity, uint64_t* n_total, int shared) {
  if (!e) return KC_EINVAL;
  if (!e->have_pairs) return fail(e, KC_EINVAL, "kc_score_pairs first");
  KC_CUDA(e, cudaSetDevice(e->dev));
  const int world = e->comm ? e->cworld : 1, rank = e->comm ? e->crank : 0;
  const uint64_t ne = e->n_edges;
  if (world == 1) {
    if (n_total) *n_total = ne;
    return kc_get_edges(e, out, capacity);
  }
  NcclApi& nc = nccl_api();
  // every rank's list = the run of its early block, then the run of its late block (both sorted by a, b)
  unsigned long long n_low = ne;
  unsigned long long* d = e->d_comm.as<unsigned long long>();  // [0..1]: my runs, [16 .. 16 + 2 world): all
  if (e->ishards > 1 && ne) {
    const uint32_t hi_start = e->block_bounds[e->block_bounds.size() - 2 - (size_t)rank];
    // (class-major pair order: rows are ranks of the pair order, the list is sorted by input index; one run)
    if (!e->cfg.cross_class_only) {
      KC_LAUNCH(e, edge_lower_bound_kernel, 1, 32, 0, e->d_edges_sorted.as<uint4>(), ne, hi_start, d);
      KC_CUDA(e, cudaMemcpyAsync(&n_low, d, 8, cudaMemcpyDeviceToHost, e->stream));
      KC_CUDA(e, cudaStreamSynchronize(e->stream));
    }
  }
  unsigned long long mine[2] = {n_low, ne - n_low};
  KC_CUDA(e, cudaMemcpyAsync(d, mine, 16, cudaMemcpyHostToDevice, e->stream));
  ncclResult_t nr = nc.AllGather(d, d + 16, 2, ncclUint64, e->comm, e->stream);
  if (nr != ncclSuccess) return fail(e, KC_ECUDA, std::string("ncclAllGather: ") + nc.GetErrorString(nr));
  std::vector<unsigned long long> runs(2 * (size_t)world);
  KC_CUDA(e, cudaMemcpyAsync(runs.data(), d + 16, 16 * (size_t)world, cudaMemcpyDeviceToHost, e->stream));
  KC_CUDA(e, cudaStreamSynchronize(e->stream));
  // block order: early runs of ranks 0 .. world-1, then late runs of ranks world-1 .. 0
  std::vector<unsigned long long> place(2 * (size_t)world);
  unsigned long long total = 0;
  for (int r = 0; r < world; ++r) {
    place[2 * r] = total;
    total += runs[2 * r];
  }
  for (int r = world - 1; r >= 0; --r) {
    place[2 * r + 1] = total;
    total += runs[2 * r + 1];
  }
  if (n_total) *n_total = total;
  const uint4* my = e->d_edges_sorted.as<uint4>();
  if (shared) {
    // (every rank passes the same buffer and sees the same total: they all return here together)
    if (total > capacity || (total && !out)) return fail(e, KC_EINVAL, "capacity too small");
    mark(e, EV_D0);
    if (mine[0]) KC_CUDA(e, cudaMemcpyAsync(out + place[2 * rank], my, mine[0] * 16, cudaMemcpyDeviceToHost, e->stream));
    if (mine[1])
      KC_CUDA(e, cudaMemcpyAsync(out + place[2 * rank + 1], my + mine[0], mine[1] * 16, cudaMemcpyDeviceToHost, e->stream));
    mark(e, EV_D1);
    KC_CUDA(e, cudaStreamSynchronize(e->stream));
    unsigned long long flag[1] = {0};  // barrier: every rank's copies have landed
    if (int rc = allreduce_u64(e, flag, 1)) return rc;
  } else {
    if (rank == 0) KC_CUDA(e, e->d_gather.ensure(std::max<unsigned long long>(total, 1) * 16));
    uint4* g = e->d_gather.as<uint4>();
    nr = nc.GroupStart();
    for (int part = 0; part < 2 && nr == ncclSuccess; ++part) {
      if (rank != 0) {
        if (mine[part]) nr = nc.Send(my + (part ? mine[0] : 0), mine[part] * 16, ncclUint8, 0, e->comm, e->stream);
      } else {
        for (int src = 1; src < world && nr == ncclSuccess; ++src)
          if (runs[2 * src + part])
            nr = nc.Recv(g + place[2 * src + part], runs[2 * src + part] * 16, ncclUint8, src, e->comm, e->stream);
      }
    }
    const ncclResult_t nr2 = nc.GroupEnd();
    if (nr == ncclSuccess) nr = nr2;
    if (nr != ncclSuccess) return fail(e, KC_ECUDA, std::string("edge gather (ncclSend/ncclRecv): ") + nc.GetErrorString(nr));
    if (rank == 0) {
      if (mine[0]) KC_CUDA(e, cudaMemcpyAsync(g + place[0], my, mine[0] * 16, cudaMemcpyDeviceToDevice, e->stream));
      if (mine[1]) KC_CUDA(e, cudaMemcpyAsync(g + place[1], my + mine[0], mine[1] * 16, cudaMemcpyDeviceToDevice, e->stream));
      // (checked only now: rank 0 has to take part in the exchange whatever its buffer holds)
      if (total > capacity || (total && !out)) {
        cudaStreamSynchronize(e->stream);
        return fail(e, KC_EINVAL, "capacity too small");
      }
      mark(e, EV_D0);
      if (total) KC_CUDA(e, cudaMemcpyAsync(out, g, total * 16, cudaMemcpyDeviceToHost, e->stream));
      mark(e, EV_D1);
    }
    KC_CUDA(e, cudaStreamSynchronize(e->stream));
  }
  // a replicated index scored in contiguous shards is in block order as well; the class-major pair order is
  // not (blocks of the pair order, list sorted by input index): restore (a, b) on the host
  if (e->cfg.cross_class_only && (rank == 0) && total) {
    std::sort(out, out + total, [](const kc_edge& x, const kc_edge& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
  }
  return KC_OK;
}

int kc_gather_edges(kc_engine* e, kc_edge* out, uint64_t capacity, uint64_t* n_total) {
  return gather_edges_impl(e, out, capacity, n_total, 0);
}
int kc_gather_edges_shared(kc_engine* e, kc_edge* shared_out, uint64_t capacity, uint64_t* n_total) {
  return gather_edges_impl(e, shared_out, capacity, n_total, 1);
}

}  // extern "C"

