// cli.cpp — `kmer_cluster <fasta> <threads>`: the reference's command line
// (`cargo run --release -- <fasta> <threads>`, src/main.rs:50-239) over the B200 engine.
// Progress lines and the five parity counters are printed on stderr with the reference's
// wording (src/main.rs:51,74,124,151,201,214; src/graph/mod.rs:50-51,545,550,695,250-251).
// The DIAMOND hand-off (src/graph/mod.rs:253-317) is out of scope; the surviving pairs are
// written to stdout as TSV instead of the Debug dump.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kc_b200.h"
#include "../../include/kc_host.h"

static void die(const char* what, kc_engine* e, int rc) {
  std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, e ? kc_last_error(e) : "");
  std::exit(101);  // the reference panics (exit status 101)
}

int main(int argc, char** argv) {
  std::fprintf(stderr, "We start main\n");
  // Arguments required: input, threads (src/main.rs:54-60); options may follow
  if (argc < 3) {
    std::fprintf(stderr, "Requires two command line arguments: input and thread\n");
    return 101;
  }
  const char* input = argv[1];
  char* endp = nullptr;
  const long threads = std::strtol(argv[2], &endp, 10);
  if (!endp || *endp || threads < 0) {
    std::fprintf(stderr, "threads argument should be of type int\n");
    return 101;
  }
  kc_config cfg{};
  cfg.k = 5;
  cfg.device = 0;
  cfg.threshold = 10;
  cfg.cross_class_only = 1;
  cfg.want_blosum = 0;
  bool list_kmers = false, want_tree = false;
  int n_gpus = 1;
  const char* handoff_dir = nullptr;
  for (int i = 3; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--k" && i + 1 < argc) cfg.k = std::atoi(argv[++i]);
    else if (a == "--threshold" && i + 1 < argc) cfg.threshold = (uint32_t)std::atoi(argv[++i]);
    else if (a == "--device" && i + 1 < argc) cfg.device = std::atoi(argv[++i]);
    else if (a == "--all-classes") cfg.cross_class_only = 0;
    else if (a == "--gpus" && i + 1 < argc) n_gpus = std::atoi(argv[++i]);  // one engine (one NCCL rank) per GPU
    else if (a == "--blosum") cfg.want_blosum = 1;
    else if (a == "--kmers") list_kmers = true;
    else if (a == "--tree") want_tree = true;
    else if (a == "--handoff" && i + 1 < argc) handoff_dir = argv[++i];  // fasta_files/, db_files/, blastp_output.tsv
    else if (a == "--sample-every" && i + 1 < argc) cfg.sample_every = (uint32_t)std::atoi(argv[++i]);
    else if (a == "--seed" && i + 1 < argc) cfg.sample_seed = std::strtoull(argv[++i], nullptr, 0);
    else if (a == "--index" && i + 1 < argc) {  // stream | bucket | table (kc_config.index_build)
      const std::string v = argv[++i];
      cfg.index_build = v == "stream" ? KC_INDEX_STREAM : v == "bucket" ? KC_INDEX_BUCKET : v == "table" ? KC_INDEX_TABLE : 99u;
      if (cfg.index_build == 99u) {
        std::fprintf(stderr, "--index takes stream, bucket or table\n");
        return 101;
      }
    }
    else {
      std::fprintf(stderr, "unknown option %s\n", a.c_str());
      return 101;
    }
  }
  kc_fasta* fa = nullptr;
  int rc = kc_fasta_parse_file(input, (int)threads, &fa);
  if (rc != KC_OK) {
    std::fprintf(stderr, "input argument should refer to an existing fasta file\n");
    return 101;
  }
  std::fprintf(stderr, "We created Protein structs\n");
  if (n_gpus < 1 || n_gpus > kc_device_count()) {
    std::fprintf(stderr, "--gpus %d: %d CUDA devices are visible\n", n_gpus, kc_device_count());
    return 101;
  }
  if (n_gpus > 1 && (list_kmers || want_tree)) {
    std::fprintf(stderr, "--kmers and --tree need the whole index on one GPU (--gpus 1)\n");
    return 101;
  }
  // one engine per GPU; with more than one, a host thread per engine and NCCL below the C ABI (kc_comm_*)
  std::vector<kc_engine*> engines((size_t)n_gpus, nullptr);
  for (int g = 0; g < n_gpus; ++g) {
    kc_config c = cfg;
    c.device = n_gpus > 1 ? g : cfg.device;
    rc = kc_create(&c, &engines[(size_t)g]);
    if (rc != KC_OK) {
      std::fprintf(stderr, "kc_create failed (%d): no usable CUDA device or bad option\n", rc);
      return 101;
    }
  }
  kc_engine* e = engines[0];
  const uint64_t n = kc_fasta_n_proteins(fa);
  if (kc_fasta_n_missing_class(fa)) {  // the reference panics on such a record (src/protein.rs:137)
    std::fprintf(stderr, "%llu record ids have fewer than 4 '|'-separated fields: no AMR class (src/protein.rs:135-138)\n",
                 (unsigned long long)kc_fasta_n_missing_class(fa));
    return 101;
  }
  kc_index_stats is{};
  kc_pair_stats ps{};
  double secs = 0.0;
  std::vector<int> rcs((size_t)n_gpus, 0);
  uint8_t comm_id[KC_COMM_ID_BYTES] = {0};
  if (n_gpus > 1 && kc_comm_unique_id(comm_id) != KC_OK) {
    std::fprintf(stderr, "NCCL is not available (libnccl.so.2)\n");
    return 101;
  }
  auto per_rank = [&](int g) {
    kc_engine* eg = engines[(size_t)g];
    int r = KC_OK;
    if (n_gpus > 1) r = kc_comm_init(eg, comm_id, g, n_gpus);
    if (!r) r = kc_set_proteins_dist(eg, kc_fasta_residues(fa), kc_fasta_offsets(fa), kc_fasta_class_ids(fa), n);
    kc_index_stats ig{};
    if (!r) r = kc_build_index_dist(eg, &ig);
    const auto t0 = std::chrono::steady_clock::now();
    kc_pair_stats pg{};
    if (!r) r = kc_score_pairs_dist(eg, &pg);
    if (g == 0) {
      is = ig;
      ps = pg;
      secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    rcs[(size_t)g] = r;
  };
  {
    std::vector<std::thread> pool;
    for (int g = 1; g < n_gpus; ++g) pool.emplace_back(per_rank, g);
    per_rank(0);
    for (auto& th : pool) th.join();
  }
  for (int g = 0; g < n_gpus; ++g)
    if (rcs[(size_t)g]) die("index / pair stage", engines[(size_t)g], rcs[(size_t)g]);
  std::fprintf(stderr, "We combined k-mers\nWe found unique k-mers\nWe made unique hash\nWe can make a graph\n");
  std::fprintf(stderr, "Number of %dmers found in at least two proteins: %llu\n", cfg.k,
               (unsigned long long)is.n_repeated);
  std::fprintf(stderr, "Number of total edges: %llu\n", (unsigned long long)ps.n_multi_edges);
  if (cfg.cross_class_only) {
    std::fprintf(stderr, "Remove edges without diverging AMR labels\n");
    std::fprintf(stderr, "Number of edges now: %llu\n", (unsigned long long)ps.n_multi_edges_kept);
  }
  std::fprintf(stderr, "Combine edges with the same two vertices\n");
  std::fprintf(stderr, "Number of edges now: %llu\n", (unsigned long long)ps.n_pairs_kept);
  std::fprintf(stderr, "Graph construction and refinement time: %g seconds\n", secs);
  std::vector<kc_edge> edges(ps.n_edges_out);
  if (n_gpus == 1) {
    rc = kc_get_edges(e, edges.data(), edges.size());
    if (rc) die("kc_get_edges", e, rc);
  } else {  // every rank copies its own sorted runs to their place in the one list
    auto gather = [&](int g) {
      uint64_t total = 0;
      rcs[(size_t)g] = kc_gather_edges_shared(engines[(size_t)g], edges.data(), edges.size(), &total);
    };
    std::vector<std::thread> pool;
    for (int g = 1; g < n_gpus; ++g) pool.emplace_back(gather, g);
    gather(0);
    for (auto& th : pool) th.join();
    for (int g = 0; g < n_gpus; ++g)
      if (rcs[(size_t)g]) die("kc_gather_edges_shared", engines[(size_t)g], rcs[(size_t)g]);
  }
  std::printf("a\tb\tid_a\tid_b\tkmers_in_common%s%s\n", cfg.want_blosum ? "\tblosum" : "",
              list_kmers ? "\tkmers" : "");
  std::vector<uint32_t> kms;
  for (size_t i = 0; i < edges.size(); ++i) {
    const kc_edge& ed = edges[i];
    std::fprintf(stderr, "Cross-checking:\n\treference protein:%s\n\tquery protein:%s\n\tkmers in common:%u\n",
                 kc_fasta_id(fa, ed.a), kc_fasta_id(fa, ed.b), ed.count);
    std::printf("%u\t%u\t%s\t%s\t%u", ed.a, ed.b, kc_fasta_id(fa, ed.a), kc_fasta_id(fa, ed.b), ed.count);
    if (cfg.want_blosum) std::printf("\t%d", ed.blosum);
    if (list_kmers) {
      kms.resize(ed.count);
      rc = kc_get_edge_kmers(e, i, kms.data(), kms.size());
      if (rc) die("kc_get_edge_kmers", e, rc);
      std::printf("\t");
      for (size_t j = 0; j < kms.size(); ++j) std::printf(j ? ",%u" : "%u", kms[j]);
    }
    std::printf("\n");
  }
  if (handoff_dir) {  // what align_and_output_pairs leaves for DIAMOND (src/graph/mod.rs:253-261,273-280,304-317)
    uint64_t n_files = 0;
    rc = kc_write_handoff(fa, edges.data(), edges.size(), handoff_dir, &n_files);
    if (rc) die("kc_write_handoff", e, rc);
    std::fprintf(stderr, "Wrote %llu fasta files and blastp_output.tsv (header) under %s\n",
                 (unsigned long long)n_files, handoff_dir);
  }
  if (want_tree) {
    // src/tree.rs (host): Tree::new + add_protein in input order over the per-protein id lists
    std::vector<uint64_t> ro(n + 1);
    std::vector<uint32_t> ids(is.nnz);
    rc = kc_get_protein_ids(e, ro.data(), ids.data(), ids.size());
    if (rc) die("kc_get_protein_ids", e, rc);
    kc_tree* tree = nullptr;
    rc = kc_tree_build(ro.data(), ids.data(), n, (uint32_t)is.n_repeated, &tree);
    if (rc) die("kc_tree_build", e, rc);
    std::vector<uint32_t> cluster(n);
    uint32_t n_clusters = 0;
    kc_tree_clusters(tree, cluster.data(), &n_clusters);
    std::fprintf(stderr, "Tree: %u top-level clusters, %llu merges, %llu proteins without k-mers in common\n",
                 n_clusters, (unsigned long long)kc_tree_n_merges(tree),
                 (unsigned long long)kc_tree_n_no_common(tree));
    std::printf("#protein\tcluster\n");
    for (uint64_t p = 0; p < n; ++p) std::printf("#%s\t%u\n", kc_fasta_id(fa, p), cluster[p]);
    kc_tree_free(tree);
  }
  for (kc_engine* eg : engines) kc_destroy(eg);
  kc_fasta_free(fa);
  return 0;
}
