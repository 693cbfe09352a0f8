// kc_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A plain C++17 restatement of the reference's k-mer clustering hot path
// (Isabella136/uniprot_kmer_based_clustering).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; the
// product (uniprot_kmer_based_clustering_b200/) never links, imports or calls it.
//
// PARITY STATUS: the reference has no tests and cannot be built here (needs nightly
// Rust + crates.io); it therefore has no golden vectors of its own.  This oracle is
// pinned against (1) the survey-time known answers of SURVEY.md §8c (independent
// numpy/scipy restatement, tests/golden/arg_golden.json), (2) a second, literal
// Python model of the reference's multigraph data structures (oracle/ref_model.py)
// on small inputs.  Everything at k=7, BLOSUM scores and MPHF ids are "parity
// unpinned" by the reference itself (dead / absent code); see DESIGN.md.
//
// Every stage cites the reference file:line it follows (paths relative to the
// reference root).  Two pair-stage modes exist on purpose:
//   mode 0 "literal": enumerate one multi-edge per (repeated k-mer, protein pair)
//          like Graph::new, filter by class like remove_uninteresting_edges, group
//          parallel edges like combine_edges, threshold like align_and_output_pairs.
//   mode 1 "fast"   : row-wise accumulation (multi-threaded); used for large inputs
//          and as the timed CPU baseline.  Tests check mode 0 == mode 1.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// src/protein.rs:9-13 — residue order defines the codes 0..20; '*' (20) is also the
// bucket for every byte that is not one of the 20 letters (src/protein.rs:49-54).
const char kAlphabet[22] = "CSTAGPDEQNHRKMILVWYF*";

// src/blosum.rs:8-30 diagonal (B62[r][r]) in the same residue order; the score of a
// shared k-mer is the sum of its residues' self-scores, code 20 scoring 0
// (framework-defined extension).
const int kSelfScore[21] = {9, 4, 5, 4, 6, 7, 6, 5, 5, 6, 8, 5, 5, 5, 4, 4, 4, 11, 7, 6, 0};

struct Edge {
  uint32_t a, b, count;
  int32_t blosum;
};

struct IndexStats {
  uint64_t n_positions, n_incidences, n_distinct, n_singleton, n_repeated, nnz;
};

struct PairStats {
  uint64_t n_multi_edges;       // src/graph/mod.rs:51  "Number of total edges"
  uint64_t n_multi_edges_kept;  // src/graph/mod.rs:695 "Number of edges now" (class filter)
  uint64_t n_pairs_kept;        // src/graph/mod.rs:545 "Number of edges now" (combine)
  uint64_t n_edges_out;         // src/graph/mod.rs:242 survivors of the threshold
  uint64_t sum_count_out;
};

// Position sampler of the optional subsampling mode (Protein::new_with_rand_fivemers,
// src/protein.rs:77-104: a tenth of the start positions, without replacement).  The reference
// draws from a thread-local RNG (not reproducible); the framework defines a counter-based
// sampler instead (include/kc_b200.h, kc_sample_position) and this is its restatement.
uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
uint32_t sample_position(uint64_t seed, uint32_t protein, uint32_t n, uint32_t x) {
  const uint32_t key = mix32((uint32_t)seed ^ mix32(protein ^ (uint32_t)(seed >> 32) ^ 0x9E3779B9u));
  uint32_t half = 1;
  while ((1ull << (2 * half)) < (uint64_t)n) ++half;
  const uint32_t mask = (1u << half) - 1u;
  uint32_t y = x;
  do {
    uint32_t L = y >> half, R = y & mask;
    for (uint32_t round = 0; round < 4; ++round) {
      const uint32_t t = L ^ (mix32(R ^ key ^ (round * 0x9E3779B9u)) & mask);
      L = R;
      R = t;
    }
    y = (L << half) | R;
  } while (y >= n);
  return y;
}

struct Oracle {
  int k = 5;
  int threads = 1;
  uint32_t sample_every = 0;
  uint64_t sample_seed = 0;
  uint64_t n = 0;
  std::vector<uint8_t> res;
  std::vector<uint64_t> off;
  std::vector<uint32_t> cls;
  uint8_t lut[256];

  std::vector<uint64_t> kpos_off;  // per protein, exclusive prefix of positions
  std::vector<uint32_t> kmers;     // one per position (duplicates kept)

  std::vector<uint32_t> distinct, distinct_cnt;  // census, sorted by k-mer
  std::vector<uint32_t> vocab, freq;             // repeated k-mers (sorted) and #proteins
  std::vector<uint64_t> row_off;                 // CSR of A[protein, repeated id]
  std::vector<uint32_t> ids;                     // sorted ascending within a row
  std::vector<uint64_t> col_off;                 // CSC
  std::vector<uint32_t> holders;                 // proteins ascending within a column
  IndexStats istats{};

  std::vector<Edge> edges;
  PairStats pstats{};
  std::string err;
  double t_extract = 0, t_index = 0, t_pairs = 0;
};

template <class F>
void parallel_for(int threads, uint64_t n, uint64_t grain, F f) {
  // the reference's own idiom: workers pull the next index from a shared cursor
  // (src/main.rs:90-95, src/graph/mod.rs:161).
  if (threads <= 1 || n <= grain) {
    f(0, 0, n);
    return;
  }
  std::atomic<uint64_t> cursor{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      for (;;) {
        uint64_t lo = cursor.fetch_add(grain);
        if (lo >= n) break;
        f(t, lo, std::min(n, lo + grain));
      }
    });
  for (auto& th : pool) th.join();
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int self_score_of(uint32_t kmer, int k) {
  int s = 0;
  for (int i = 0; i < k; ++i) {
    s += kSelfScore[kmer % 21];
    kmer /= 21;
  }
  return s;
}

}  // namespace

extern "C" {

Oracle* ko_create(int k, int threads) {
  if (k != 5 && k != 7) return nullptr;  // src/tree.rs:104 panics on any other size
  Oracle* o = new Oracle();
  o->k = k;
  o->threads = threads < 1 ? 1 : threads;
  std::memset(o->lut, 20, sizeof(o->lut));  // unwrap_or(20), src/protein.rs:50-51
  for (int i = 0; i < 21; ++i) o->lut[(unsigned char)kAlphabet[i]] = (uint8_t)i;
  return o;
}

void ko_destroy(Oracle* o) { delete o; }
void ko_set_sampling(Oracle* o, uint32_t every, uint64_t seed) {
  o->sample_every = every;
  o->sample_seed = seed;
}
const char* ko_last_error(Oracle* o) { return o->err.c_str(); }

int ko_set_proteins(Oracle* o, const uint8_t* residues, const uint64_t* offsets,
                    const uint32_t* class_id, uint64_t n) {
  o->n = n;
  o->off.assign(offsets, offsets + n + 1);
  o->res.assign(residues, residues + offsets[n]);
  o->cls.assign(class_id, class_id + n);
  return 0;
}

// Protein::new, src/protein.rs:107-132: one k-mer per start position, in order,
// duplicates kept; create_five_mer src/protein.rs:29-37 (first residue most
// significant, base 21).  Proteins shorter than k yield nothing.
int ko_extract(Oracle* o) {
  double t0 = now_s();
  const int k = o->k;
  o->kpos_off.assign(o->n + 1, 0);
  const uint64_t every = o->sample_every > 1 ? o->sample_every : 1;
  for (uint64_t p = 0; p < o->n; ++p) {
    uint64_t len = o->off[p + 1] - o->off[p];
    o->kpos_off[p + 1] = o->kpos_off[p] + (len >= (uint64_t)k ? (len - k + 1) / every : 0);
  }
  o->kmers.resize(o->kpos_off[o->n]);
  parallel_for(o->threads, o->n, 256, [&](int, uint64_t lo, uint64_t hi) {
    for (uint64_t p = lo; p < hi; ++p) {
      const uint8_t* s = o->res.data() + o->off[p];
      uint64_t npos = o->kpos_off[p + 1] - o->kpos_off[p];
      uint32_t* out = o->kmers.data() + o->kpos_off[p];
      const uint64_t len = o->off[p + 1] - o->off[p];
      for (uint64_t i = 0; i < npos; ++i) {
        const uint64_t pos =
            every > 1 ? sample_position(o->sample_seed, (uint32_t)p, (uint32_t)(len - k + 1), (uint32_t)i) : i;
        uint32_t v = 0;
        for (int j = 0; j < k; ++j) v = v * 21u + o->lut[s[pos + j]];
        out[i] = v;
      }
    }
  });
  o->t_extract = now_s() - t0;
  return 0;
}

uint64_t ko_n_positions(Oracle* o) { return o->kmers.size(); }
void ko_get_kmers(Oracle* o, uint32_t* out) {
  std::memcpy(out, o->kmers.data(), o->kmers.size() * 4);
}

// Census + split + id assignment + rewrite: src/main.rs:84-199.
int ko_build_index(Oracle* o, IndexStats* st) {
  double t0 = now_s();
  const uint64_t n = o->n;
  // per-protein sort + dedup (src/main.rs:100-102)
  std::vector<uint64_t> dcount(n + 1, 0);
  std::vector<uint32_t> sorted(o->kmers);
  parallel_for(o->threads, n, 256, [&](int, uint64_t lo, uint64_t hi) {
    for (uint64_t p = lo; p < hi; ++p) {
      uint32_t* b = sorted.data() + o->kpos_off[p];
      uint32_t* e = sorted.data() + o->kpos_off[p + 1];
      std::sort(b, e);
      dcount[p + 1] = std::unique(b, e) - b;
    }
  });
  for (uint64_t p = 0; p < n; ++p) dcount[p + 1] += dcount[p];
  std::vector<uint32_t> inc(dcount[n]);
  parallel_for(o->threads, n, 256, [&](int, uint64_t lo, uint64_t hi) {
    for (uint64_t p = lo; p < hi; ++p)
      std::memcpy(inc.data() + dcount[p], sorted.data() + o->kpos_off[p],
                  (dcount[p + 1] - dcount[p]) * 4);
  });
  // global census: the reference keeps one sorted (kmer, n_proteins) list
  // (merge_sort, src/main.rs:23-48); sort + run-length gives the same list.
  {
    std::vector<uint32_t> all(inc);
    const int T = o->threads;
    if (T > 1 && all.size() > (1u << 20)) {
      std::vector<uint64_t> cut(T + 1);
      for (int t = 0; t <= T; ++t) cut[t] = all.size() * (uint64_t)t / T;
      std::vector<std::thread> pool;
      for (int t = 0; t < T; ++t)
        pool.emplace_back([&, t] { std::sort(all.begin() + cut[t], all.begin() + cut[t + 1]); });
      for (auto& th : pool) th.join();
      for (int step = 1; step < T; step *= 2) {
        std::vector<std::thread> mp;
        for (int t = 0; t + step < T; t += 2 * step)
          mp.emplace_back([&, t, step] {
            std::inplace_merge(all.begin() + cut[t], all.begin() + cut[t + step],
                               all.begin() + cut[std::min(T, t + 2 * step)]);
          });
        for (auto& th : mp) th.join();
      }
    } else {
      std::sort(all.begin(), all.end());
    }
    o->distinct.clear();
    o->distinct_cnt.clear();
    for (uint64_t i = 0; i < all.size();) {
      uint64_t j = i;
      while (j < all.size() && all[j] == all[i]) ++j;
      o->distinct.push_back(all[i]);
      o->distinct_cnt.push_back((uint32_t)(j - i));
      i = j;
    }
  }
  // split unique / repeated (src/main.rs:127-137).  boomphf ids are an arbitrary
  // bijection; canonical id = rank among repeated k-mers, ascending.
  o->vocab.clear();
  o->freq.clear();
  uint64_t singles = 0;
  for (size_t i = 0; i < o->distinct.size(); ++i) {
    if (o->distinct_cnt[i] == 1)
      ++singles;
    else {
      o->vocab.push_back(o->distinct[i]);
      o->freq.push_back(o->distinct_cnt[i]);
    }
  }
  // rewrite: remove_unique_five_mers + modify_hash_five_mer (src/protein.rs:151-174);
  // id lists are kept sorted (update_graph_edges sorts them anyway,
  // src/graph/vertex.rs:82-85).
  o->row_off.assign(n + 1, 0);
  std::vector<uint32_t> tmp(inc.size());
  parallel_for(o->threads, n, 256, [&](int, uint64_t lo, uint64_t hi) {
    for (uint64_t p = lo; p < hi; ++p) {
      uint64_t w = dcount[p];
      for (uint64_t i = dcount[p]; i < dcount[p + 1]; ++i) {
        auto it = std::lower_bound(o->vocab.begin(), o->vocab.end(), inc[i]);
        if (it != o->vocab.end() && *it == inc[i]) tmp[w++] = (uint32_t)(it - o->vocab.begin());
      }
      o->row_off[p + 1] = w - dcount[p];
    }
  });
  for (uint64_t p = 0; p < n; ++p) o->row_off[p + 1] += o->row_off[p];
  o->ids.resize(o->row_off[n]);
  for (uint64_t p = 0; p < n; ++p)
    std::memcpy(o->ids.data() + o->row_off[p], tmp.data() + dcount[p],
                (o->row_off[p + 1] - o->row_off[p]) * 4);
  // inverted index (what times_kmer_visited + the triangular edge layout encode,
  // src/graph/vertex.rs:92-136): holders of each id in ascending protein order.
  const uint64_t V = o->vocab.size();
  o->col_off.assign(V + 1, 0);
  for (uint32_t id : o->ids) o->col_off[id + 1]++;
  for (uint64_t v = 0; v < V; ++v) o->col_off[v + 1] += o->col_off[v];
  o->holders.resize(o->ids.size());
  {
    std::vector<uint64_t> cur(o->col_off.begin(), o->col_off.end() - 1);
    for (uint64_t p = 0; p < n; ++p)
      for (uint64_t i = o->row_off[p]; i < o->row_off[p + 1]; ++i)
        o->holders[cur[o->ids[i]]++] = (uint32_t)p;
  }
  o->istats = {o->kmers.size(), inc.size(), o->distinct.size(), singles, V, o->ids.size()};
  if (st) *st = o->istats;
  o->t_index = now_s() - t0;
  return 0;
}

uint64_t ko_n_distinct(Oracle* o) { return o->distinct.size(); }
void ko_get_distinct(Oracle* o, uint32_t* kmers, uint32_t* counts) {
  std::memcpy(kmers, o->distinct.data(), o->distinct.size() * 4);
  if (counts) std::memcpy(counts, o->distinct_cnt.data(), o->distinct_cnt.size() * 4);
}
void ko_get_vocab(Oracle* o, uint32_t* kmers, uint32_t* freq) {
  std::memcpy(kmers, o->vocab.data(), o->vocab.size() * 4);
  if (freq) std::memcpy(freq, o->freq.data(), o->freq.size() * 4);
}
void ko_get_protein_ids(Oracle* o, uint64_t* row_off, uint32_t* ids) {
  std::memcpy(row_off, o->row_off.data(), o->row_off.size() * 8);
  std::memcpy(ids, o->ids.data(), o->ids.size() * 4);
}

static int32_t blosum_by_intersection(const Oracle* o, const std::vector<int>& ss, uint32_t a,
                                      uint32_t b) {
  uint64_t i = o->row_off[a], ie = o->row_off[a + 1], j = o->row_off[b], je = o->row_off[b + 1];
  int32_t s = 0;
  while (i < ie && j < je) {
    if (o->ids[i] < o->ids[j])
      ++i;
    else if (o->ids[i] > o->ids[j])
      ++j;
    else {
      s += ss[o->ids[i]];
      ++i;
      ++j;
    }
  }
  return s;
}

// Graph::new -> remove_uninteresting_edges -> combine_edges -> threshold.
// rank/world select a subset of rows for the multi-GPU tests: only pairs whose lower
// protein a satisfies row_lo <= a < row_hi are produced (row_hi == 0 means all).
int ko_score_pairs(Oracle* o, uint32_t threshold, int cross_class_only, int want_blosum, int mode,
                   uint64_t row_lo, uint64_t row_hi, PairStats* st) {
  double t0 = now_s();
  const uint64_t n = o->n, V = o->vocab.size();
  if (row_hi == 0) row_hi = n;
  std::vector<int> ss(V);
  for (uint64_t v = 0; v < V; ++v) ss[v] = self_score_of(o->vocab[v], o->k);
  PairStats ps{};
  // edge count per k-mer f(f-1)/2 and its sum: src/graph/mod.rs:44-51
  for (uint64_t v = 0; v < V; ++v) ps.n_multi_edges += (uint64_t)o->freq[v] * (o->freq[v] - 1) / 2;
  o->edges.clear();

  if (mode == 0) {
    // one multi-edge per (k-mer, arrival i < arrival j): src/graph/vertex.rs:92-136
    std::vector<uint64_t> keys;
    for (uint64_t v = 0; v < V; ++v) {
      const uint32_t* h = o->holders.data() + o->col_off[v];
      const uint64_t f = o->col_off[v + 1] - o->col_off[v];
      for (uint64_t i = 0; i < f; ++i) {
        if (h[i] < row_lo || h[i] >= row_hi) continue;
        for (uint64_t j = i + 1; j < f; ++j) {
          // remove_uninteresting_edges: keep when the class strings differ,
          // src/graph/mod.rs:580-587
          if (cross_class_only && o->cls[h[i]] == o->cls[h[j]]) continue;
          keys.push_back(((uint64_t)h[i] << 32) | h[j]);
          if (keys.size() > (3ull << 30)) {
            o->err = "literal mode: too many multi-edges, use mode 1";
            return 1;
          }
        }
      }
    }
    ps.n_multi_edges_kept = keys.size();
    // combine_edges: one edge per protein pair carrying all shared k-mers,
    // src/graph/mod.rs:393-440, src/graph/edge.rs:56-85
    std::sort(keys.begin(), keys.end());
    for (uint64_t i = 0; i < keys.size();) {
      uint64_t j = i;
      while (j < keys.size() && keys[j] == keys[i]) ++j;
      ps.n_pairs_kept++;
      uint32_t cnt = (uint32_t)(j - i);
      // align_and_output_pairs: `len() <= 10 -> continue`, src/graph/mod.rs:242
      if (cnt > threshold) {
        Edge e{(uint32_t)(keys[i] >> 32), (uint32_t)keys[i], cnt, 0};
        if (want_blosum) e.blosum = blosum_by_intersection(o, ss, e.a, e.b);
        o->edges.push_back(e);
        ps.sum_count_out += cnt;
      }
      i = j;
    }
  } else {
    const int T = o->threads;
    std::vector<std::vector<Edge>> out(T);
    std::vector<PairStats> tps(T);
    std::vector<std::vector<uint32_t>> cnt(T), touched(T);
    std::vector<std::vector<int32_t>> sc(T);
    for (int t = 0; t < T; ++t) {
      cnt[t].assign(n, 0);
      if (want_blosum) sc[t].assign(n, 0);
    }
    parallel_for(T, row_hi - row_lo, 16, [&](int t, uint64_t lo, uint64_t hi) {
      auto& c = cnt[t];
      auto& tl = touched[t];
      for (uint64_t a = row_lo + lo; a < row_lo + hi; ++a) {
        tl.clear();
        for (uint64_t i = o->row_off[a]; i < o->row_off[a + 1]; ++i) {
          const uint32_t id = o->ids[i];
          const uint32_t* hb = o->holders.data() + o->col_off[id];
          const uint32_t* he = o->holders.data() + o->col_off[id + 1];
          const uint32_t* h = std::upper_bound(hb, he, (uint32_t)a);
          for (; h < he; ++h) {
            if (c[*h]++ == 0) tl.push_back(*h);
            if (want_blosum) sc[t][*h] += ss[id];
          }
        }
        for (uint32_t b : tl) {
          const uint32_t v = c[b];
          const int32_t s = want_blosum ? sc[t][b] : 0;
          c[b] = 0;
          if (want_blosum) sc[t][b] = 0;
          if (cross_class_only && o->cls[a] == o->cls[b]) continue;
          tps[t].n_multi_edges_kept += v;
          tps[t].n_pairs_kept++;
          if (v > threshold) {
            out[t].push_back(Edge{(uint32_t)a, b, v, s});
            tps[t].sum_count_out += v;
          }
        }
      }
    });
    for (int t = 0; t < T; ++t) {
      ps.n_multi_edges_kept += tps[t].n_multi_edges_kept;
      ps.n_pairs_kept += tps[t].n_pairs_kept;
      ps.sum_count_out += tps[t].sum_count_out;
      o->edges.insert(o->edges.end(), out[t].begin(), out[t].end());
    }
    std::sort(o->edges.begin(), o->edges.end(), [](const Edge& x, const Edge& y) {
      return x.a != y.a ? x.a < y.a : x.b < y.b;
    });
  }
  ps.n_edges_out = o->edges.size();
  o->pstats = ps;
  if (st) *st = ps;
  o->t_pairs = now_s() - t0;
  return 0;
}

uint64_t ko_n_edges(Oracle* o) { return o->edges.size(); }
void ko_get_edges(Oracle* o, Edge* out) {
  std::memcpy(out, o->edges.data(), o->edges.size() * sizeof(Edge));
}
void ko_get_times(Oracle* o, double* t3) {
  t3[0] = o->t_extract;
  t3[1] = o->t_index;
  t3[2] = o->t_pairs;
}

// BLOSUM62 self-score of one packed k-mer (exposed so tests can pin the table).
int ko_self_score(uint32_t kmer, int k) { return self_score_of(kmer, k); }

}  // extern "C"
