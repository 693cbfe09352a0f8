/* kc_b200.h — C ABI of the B200-native k-mer clustering engine.
 *
 * Drop-in boundary for the hot path of Isabella136/uniprot_kmer_based_clustering.
 * The reference is a single Rust binary with no FFI of its own; the functions below are
 * what an FFI crate for that path binds, one entry point per reference stage.  Citations
 * (file:line) are relative to the reference root.  INTEGRATION.md shows the Rust
 * `extern "C"` block and the safe wrappers that keep the reference's module API.
 *
 * Conventions
 *   - One opaque engine per GPU (one process per GPU; the engine owns all device memory).
 *   - Every call returns 0 (KC_OK) or a KC_E* code; the message is kc_last_error(engine).
 *     Nothing panics or throws across the boundary (the reference panics, e.g.
 *     src/main.rs:55-63; the wrappers turn codes back into panics where parity wants it).
 *   - Host output goes into caller-allocated buffers sized from the stats structs.
 *   - Calls on one engine come from one host thread at a time.
 *   - There is no CPU fallback: without a CUDA device kc_create fails with KC_ENODEVICE.
 *   - Protein indices are 0-based in the order given to kc_set_proteins (FASTA order,
 *     SURVEY C1).  k-mer ids are canonical: rank of the k-mer among the repeated k-mers
 *     in ascending k-mer order (replaces boomphf ids).
 */
#ifndef KC_B200_H_
#define KC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KC_ABI_VERSION 2

enum {
  KC_OK = 0,
  KC_EINVAL = 1,    /* bad argument / call order */
  KC_ENODEVICE = 2, /* no usable CUDA device */
  KC_ECUDA = 3,     /* CUDA runtime error, see kc_last_error */
  KC_ENOMEM = 4,    /* device or host allocation failed */
  KC_ETOOLARGE = 5  /* input exceeds a documented limit (2^32-1 residues, 2^32-1 nnz) */
};

typedef struct kc_engine kc_engine;

typedef struct kc_config {
  int32_t k;                /* 5 (reference, src/protein.rs:29-37) or 7 (src/tree.rs:96-101) */
  int32_t device;           /* CUDA device ordinal */
  uint32_t threshold;       /* emit pairs with count > threshold; reference 10, src/graph/mod.rs:242 */
  int32_t cross_class_only; /* 1 = reference (remove_uninteresting_edges, src/graph/mod.rs:580-587) */
  int32_t want_blosum;      /* 1 = fill kc_edge.blosum (src/blosum.rs table, framework-defined score) */
  uint32_t sample_every;    /* 0/1 = every k-mer position (reference's live path); d > 1 = keep
                             * floor(positions / d) random start positions per protein, without
                             * replacement (Protein::new_with_rand_fivemers, d = 10, src/protein.rs:77-104) */
  uint64_t max_edges;       /* device edge-buffer capacity; 0 = automatic (grows and retries) */
  uint64_t sample_seed;     /* seed of the counter-based position sampler (kc_sample_position) */
  /* Tuning knobs (ABI 2; all 0 = the engine's own choice).  They are per engine: nothing on the product
   * path reads the process environment. */
  uint32_t index_build;     /* KC_INDEX_*: which build kc_build_index runs */
  uint32_t bucket_cap;      /* streaming build: records a shared-memory bucket takes, 4096 (default) or 512
                             * (tests: pushes ordinary buckets through the global-memory path) */
  uint32_t index_slices;    /* table build: number of L2 slices of the k-mer universe (0 = from the footprint) */
  uint32_t census_merge;    /* table build: slices merged per census launch (0/1 = none) */
  uint32_t pair_lists;      /* table build: 1 = materialise the multi-edge lists for the stream pair kernel */
  uint32_t no_upload_overlap; /* 0 = a large residue stream is uploaded in chunks on a copy stream and the index
                             * build starts on the first chunk; 1 = one upload on the main stream; 2 = chunked
                             * at any size (tests) */
} kc_config;

enum {
  KC_INDEX_AUTO = 0,   /* the streaming partitioned build, unless subsampling is on (then the table build) */
  KC_INDEX_STREAM = 1, /* stream_index.cuh: stable two-level partition + one shared-memory pass per bucket */
  KC_INDEX_BUCKET = 2, /* bucket.cuh: round 1's append-by-cursor partitioned build (kept for comparison) */
  KC_INDEX_TABLE = 3   /* index.cuh: universe-sized tables, L2-sliced */
};

/* src/main.rs:84-147 census + split; nnz = sum over proteins of |get_five_hash()| */
typedef struct kc_index_stats {
  uint64_t n_positions;  /* sum of max(0, len-k+1): |get_five_mers()| over all proteins */
  uint64_t n_incidences; /* sum of per-protein distinct k-mers (src/main.rs:100-102) */
  uint64_t n_distinct;   /* census length (src/main.rs:138) */
  uint64_t n_singleton;  /* five_mer_unique (src/main.rs:131-133) */
  uint64_t n_repeated;   /* five_mer_repeat, printed at src/graph/mod.rs:50 */
  uint64_t nnz;          /* sum of kmer_freq (src/main.rs:187-193) */
} kc_index_stats;

typedef struct kc_pair_stats {
  uint64_t n_multi_edges;      /* "Number of total edges", src/graph/mod.rs:51 (whole set) */
  uint64_t n_multi_edges_kept; /* "Number of edges now" after the class filter, :695 */
  uint64_t n_pairs_kept;       /* "Number of edges now" after combine_edges, :545 */
  uint64_t n_edges_out;        /* pairs with count > threshold, :242 */
  uint64_t sum_count_out;      /* sum of their counts */
  uint64_t n_rows;             /* proteins (rows of the pair triangle) this call scored */
  uint64_t n_retries;          /* edge-buffer growth retries (0 in steady state) */
  uint64_t n_rows_rescored;    /* rows whose optimistic on-chip table overflowed and were scored again */
} kc_pair_stats;

/* One surviving pair = KmerEdgeGroup {vertices_key, kmers.len()} (src/graph/edge.rs:48-52),
 * a < b in input order, list sorted by (a, b). */
typedef struct kc_edge {
  uint32_t a, b, count;
  int32_t blosum;
} kc_edge;

/* CUDA-event timings of the last calls (milliseconds, on the engine's stream) */
typedef struct kc_timings {
  float h2d_ms, extract_ms, index_ms, pairs_ms, edges_ms, d2h_ms;
  uint32_t kernel_launches; /* launches of this library's kernels since kc_reset_timings */
  uint32_t reserved;
  float pair_kernel_ms;     /* accumulation kernels only (the roofline kernel family) */
  float census_kernel_ms;   /* extract+dedup+census kernels only (streaming build: the partition kernels) */
  uint64_t index_records;   /* streaming build: (k-mer, row) records partitioned (a sharded build: what it kept) */
  uint32_t index_mid_buckets;  /* buckets that took the CTA path / the global-memory path */
  uint32_t index_huge_buckets;
} kc_timings;

/* The sampler (host-callable, same integer arithmetic as the kernels): the x-th sampled start
 * position, x < floor(n_positions / sample_every), of the protein with input index `protein`.
 * A 4-round Feistel permutation of [0, n_positions) with cycle walking, keyed by (seed, protein):
 * distinct positions, reproducible, order-independent. */
uint32_t kc_sample_position(uint64_t seed, uint32_t protein, uint32_t n_positions, uint32_t x);

int kc_abi_version(void);
int kc_device_count(void);

int kc_create(const kc_config* cfg, kc_engine** out);
void kc_destroy(kc_engine* e);
const char* kc_last_error(const kc_engine* e);
/* run the engine's work on an existing CUDA stream (cudaStream_t); NULL = engine-owned stream */
int kc_set_stream(kc_engine* e, void* cuda_stream);

/* Stage the residue stream: replaces the Vec<Protein> built at src/main.rs:62-72.
 * residues = ASCII bytes of all sequences back to back (no separators, no line breaks),
 * offsets[n+1] delimit proteins, class_id[n] = dictionary id of the AMR class string
 * (Protein::get_amr_class, src/protein.rs:135-138).  Host pointers; copied to HBM ASYNCHRONOUSLY (the upload overlaps
 * the host-side staging and, for page-locked buffers, the first kernels: the streaming index build sorts upload
 * chunk u while chunk u + 1 crosses PCIe), and the engine takes its own host copies of offsets / classes inside the
 * next build, behind that build's kernels: the three buffers must stay valid and unchanged until the next
 * kc_build_index* / kc_extract_kmers on this engine has returned.  offsets[0] must be 0 and
 * the offsets non-decreasing (KC_EINVAL otherwise; a failed call leaves the engine without proteins). */
int kc_set_proteins(kc_engine* e, const uint8_t* residues, const uint64_t* offsets,
                    const uint32_t* class_id, uint64_t n_proteins);
/* Same, inputs already resident in HBM on the engine's device.  The engine COPIES them (device to device, on its
 * stream; the offsets and classes also come back to the host for the row layout): the caller's buffers may be reused
 * once the next kc_build_index* / kc_extract_kmers has returned. */
int kc_set_proteins_device(kc_engine* e, const uint8_t* d_residues, const uint64_t* d_offsets,
                           const uint32_t* d_class_id, uint64_t n_proteins);

/* Same, residue stream resident in HBM (e.g. all-gathered over NVLink from per-rank slices: every
 * rank of a multi-GPU job needs the whole stream, but only 1/world of it has to cross its PCIe link),
 * offsets and classes on the host. */
int kc_set_proteins_device_residues(kc_engine* e, const uint8_t* d_residues, const uint64_t* offsets,
                                    const uint32_t* class_id, uint64_t n_proteins);

/* Protein::new + get_five_mers (src/protein.rs:107-132,141): one packed k-mer per start
 * position, proteins back to back, duplicates kept.  kmers_out may be NULL (compute only);
 * otherwise it receives n_positions u32 values (capacity checked). */
int kc_extract_kmers(kc_engine* e, uint32_t* kmers_out, uint64_t capacity, uint64_t* n_positions);

/* Census, unique/repeated split, perfect k-mer index, per-protein id lists, kmer_freq
 * (src/main.rs:84-199; replaces boomphf Mphf::new at :139-140). */
int kc_build_index(kc_engine* e, kc_index_stats* stats);
/* Multi-GPU: the index of ONE rank's rows of the pair triangle.  The pair order is cut into
 * 2 * n_shards row blocks of equal k-mer positions; rank `shard` owns blocks shard and
 * 2 * n_shards - 1 - shard (one early, one late: a row is scored against the rows after it, so its
 * work falls with its position).  A rank only needs the k-mers its own rows hold, with all
 * their holders, so every rank builds its part from the whole residue stream with no exchange
 * (owner computes): the stats are totals over the k-mers whose first holder is in the block and
 * add up to kc_build_index's over the shards.  Follow with kc_score_pairs_shard(shard, n_shards).
 * The readback / lookup entry points need a whole index.  Which build runs (kc_config.index_build = AUTO): round 1's
 * bucket build, and the streaming build of the SAME shard when one of its buckets overflows; both cut the same row
 * blocks, so ranks may differ in the build they ran.  Only the table build (KC_INDEX_TABLE, or subsampling) builds the
 * whole index on every rank: its stats are the whole-set numbers everywhere and kc_index_shard_info reports
 * n_shards = 1 (kc_score_pairs_shard then cuts work-balanced contiguous row blocks). */
int kc_build_index_shard(kc_engine* e, uint32_t shard, uint32_t n_shards, kc_index_stats* stats);
/* What the engine's current index covers: info[0..3] = {shard, n_shards, n_blocks, rows owned};
 * n_shards == 1 means a whole index (stats are whole-set numbers). */
int kc_index_shard_info(kc_engine* e, uint32_t info[4]);
/* Which build produced the current index: 0 = universe-table build (index.cuh), 1 = streaming partitioned
 * build (stream_index.cuh), else the bucket slot size of round 1's partitioned build (bucket.cuh): 4096, or
 * 8192 after a bucket overflow. */
int kc_index_flavour(kc_engine* e);
/* bounds[n_blocks + 1]: the row blocks of the pair order (block b is owned by rank b if b < n_shards,
 * else by rank n_blocks - 1 - b); a whole index has the one block {0, n}. */
int kc_index_shard_blocks(kc_engine* e, uint32_t* bounds, uint32_t capacity);
/* all distinct k-mers, ascending (the census keys, src/main.rs:138) */
int kc_get_distinct_kmers(kc_engine* e, uint32_t* kmers_out, uint64_t capacity);
/* repeated k-mers ascending (= id order) and kmer_freq[id] (src/main.rs:135,187-193) */
int kc_get_vocab(kc_engine* e, uint32_t* kmers_out, uint32_t* freq_out, uint64_t capacity);
/* Protein::get_five_hash for every protein as CSR: row_offsets[n+1], ids[nnz] ascending
 * within a row (src/protein.rs:146-148; vertex.rs:82-85 sorts them before use) */
int kc_get_protein_ids(kc_engine* e, uint64_t* row_offsets, uint32_t* ids_out, uint64_t capacity);
/* Mphf::hash (src/main.rs:145,192): id of each k-mer, 0xFFFFFFFF if not a repeated k-mer */
int kc_lookup_kmers(kc_engine* e, const uint32_t* kmers, uint64_t n, uint32_t* ids_out);
/* The index exactly as the pair stage reads it: the minimal perfect hash the build assigned
 * (like boomphf's ids at src/main.rs:139-147 it is an arbitrary bijection onto [0, n_repeated);
 * kc_get_vocab / kc_get_protein_ids / kc_lookup_kmers present the canonical ascending-k-mer view).
 * kmers_out[id] / freq_out[id] / self_out[id] (BLOSUM62 self-score) per id, row_offsets[n+1] and
 * ids_out[nnz] = every protein's ids in input protein order, unsorted within a row.
 * Any output pointer may be NULL. */
int kc_get_pair_index(kc_engine* e, uint32_t* kmers_out, uint32_t* freq_out, uint8_t* self_out,
                      uint64_t capacity_vocab, uint64_t* row_offsets, uint32_t* ids_out, uint64_t capacity_ids);

/* Graph::new + remove_uninteresting_edges + combine_edges + the threshold of
 * align_and_output_pairs (src/graph/mod.rs:39-193, 549-697, 322-546, 242) in one pass;
 * the multigraph is never materialised. */
int kc_score_pairs(kc_engine* e, kc_pair_stats* stats);
/* Same for shard `shard` of `n_shards` work-balanced row blocks of the pair triangle
 * (multi-GPU: every rank holds the same index and scores its own shard). */
int kc_score_pairs_shard(kc_engine* e, uint32_t shard, uint32_t n_shards, kc_pair_stats* stats);
/* the surviving pairs of the last kc_score_pairs*, sorted by (a, b) */
int kc_get_edges(kc_engine* e, kc_edge* edges_out, uint64_t capacity);
/* device-resident view of the same list (valid until the next kc_score_pairs* / kc_destroy):
 * lets a multi-GPU host gather edge lists GPU-to-GPU (NCCL) without a host round trip */
int kc_get_edges_device(kc_engine* e, const kc_edge** d_edges_out, uint64_t* n_edges_out);
/* KmerEdgeGroup.kmers (src/graph/edge.rs:49,74) for edge i of the last result, as k-mer
 * VALUES ascending; kmers_out holds edges[i].count entries */
int kc_get_edge_kmers(kc_engine* e, uint64_t edge_index, uint32_t* kmers_out, uint64_t capacity);

/* ---- Multi-GPU: one engine per GPU, one NCCL rank per engine (one process per GPU, or one host thread per
 * GPU).  NCCL lives below this ABI (csrc/dist.cuh; libnccl.so.2 is loaded at run time), so a Rust / C++ host
 * needs no collective library of its own.  The pair triangle is cut into 2 * world row blocks, rank g owns
 * blocks g and 2 * world - 1 - g and computes everything for its rows (kc_build_index_shard /
 * kc_score_pairs_shard); what crosses NVLink is the residue staging (1 / world uploaded per rank, all-gathered),
 * two counter all-reduces and the edge gather.  The reference has no counterpart: it is one process
 * (threads + mutexes, src/main.rs:84-122, src/graph/mod.rs:81-182). */
#define KC_COMM_ID_BYTES 128
/* rank 0 makes the id (ncclGetUniqueId); the host hands the same 128 bytes to every rank */
int kc_comm_unique_id(uint8_t* id);
int kc_comm_init(kc_engine* e, const uint8_t* id, int rank, int world);
int kc_comm_info(kc_engine* e, int* rank, int* world);
/* kc_set_proteins with every rank given the SAME host arrays: the rank uploads 1 / world of the residue
 * stream over its own PCIe link, the slices are all-gathered over NVLink.  Collective. */
int kc_set_proteins_dist(kc_engine* e, const uint8_t* residues, const uint64_t* offsets,
                         const uint32_t* class_id, uint64_t n_proteins);
/* kc_build_index_shard(rank, world) + all-reduce: `stats` are the whole-set numbers on every rank.  Collective. */
int kc_build_index_dist(kc_engine* e, kc_index_stats* stats);
/* kc_score_pairs_shard(rank, world) + all-reduce: whole-job counters on every rank.  Collective. */
int kc_score_pairs_dist(kc_engine* e, kc_pair_stats* stats);
/* The edge lists of all ranks as one list sorted by (a, b), n_total edges.  Collective.
 * kc_gather_edges: device-to-device over NVLink to rank 0, one copy into rank 0's `out` (other ranks: NULL).
 * kc_gather_edges_shared: `shared_out` is ONE host buffer mapped by every rank (threads of one process, or
 * POSIX shared memory): every rank copies its own runs to their final place over its own PCIe link. */
int kc_gather_edges(kc_engine* e, kc_edge* out, uint64_t capacity, uint64_t* n_total);
int kc_gather_edges_shared(kc_engine* e, kc_edge* shared_out, uint64_t capacity, uint64_t* n_total);

int kc_get_timings(kc_engine* e, kc_timings* out);
int kc_reset_timings(kc_engine* e);

/* Dense presence-bitset path (north star item 3): counts[i*n_rows+j] = |K_rows[i] ∩ K_rows[j]| over the
 * repeated-k-mer vocabulary, computed with AND + popcount on per-protein bitsets (the operation of
 * intersect_bitarrays, src/tree.rs:21-45, and of the |c_i ∩ c_j| loop at src/tree.rs:185-216, as a dense
 * all-pairs tile).  A secondary path: the sparse pair stage does ~10^3 times less work on these sets (DESIGN.md
 * §6); the host tree (csrc/tree.cpp) keeps its own sorted-list intersections.  rows = protein indices (host),
 * counts = host. */
int kc_bitset_pair_counts(kc_engine* e, const uint32_t* rows, uint32_t n_rows, uint32_t* counts_out);
/* AND + POPC + ADD issue rate of this GPU in 10^9 operations per second (register-only chains): the measured
 * denominator of the bitset path's roofline (profiles/r2_bitset.md). */
int kc_popc_microbench(kc_engine* e, uint32_t iters, double* gpopc_per_s);

#ifdef __cplusplus
}
#endif
#endif /* KC_B200_H_ */
