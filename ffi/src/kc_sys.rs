//! Raw bindings of include/kc_b200.h (ABI 2) and include/kc_host.h.  One `extern "C"` item per entry point;
//! the comments name the reference code each one replaces (paths relative to the reference root).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const KC_ABI_VERSION: c_int = 2;
pub const KC_OK: c_int = 0;
pub const KC_COMM_ID_BYTES: usize = 128;
pub const KC_INDEX_AUTO: u32 = 0;
pub const KC_INDEX_STREAM: u32 = 1;
pub const KC_INDEX_BUCKET: u32 = 2;
pub const KC_INDEX_TABLE: u32 = 3;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct kc_config {
    pub k: i32,                // 5 (src/protein.rs:29-37) or 7 (src/tree.rs:96-101)
    pub device: i32,
    pub threshold: u32,        // src/graph/mod.rs:242
    pub cross_class_only: i32, // src/graph/mod.rs:580-587
    pub want_blosum: i32,      // src/blosum.rs
    pub sample_every: u32,     // src/protein.rs:77-104
    pub max_edges: u64,
    pub sample_seed: u64,
    pub index_build: u32,      // KC_INDEX_*
    pub bucket_cap: u32,
    pub index_slices: u32,
    pub census_merge: u32,
    pub pair_lists: u32,
    pub no_upload_overlap: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct kc_index_stats {
    pub n_positions: u64,
    pub n_incidences: u64,
    pub n_distinct: u64,
    pub n_singleton: u64,
    pub n_repeated: u64, // printed at src/graph/mod.rs:50
    pub nnz: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct kc_pair_stats {
    pub n_multi_edges: u64,      // src/graph/mod.rs:51
    pub n_multi_edges_kept: u64, // :695
    pub n_pairs_kept: u64,       // :545
    pub n_edges_out: u64,        // :242
    pub sum_count_out: u64,
    pub n_rows: u64,
    pub n_retries: u64,
    pub n_rows_rescored: u64,
}

/// KmerEdgeGroup {vertices_key, kmers.len()} (src/graph/edge.rs:48-52); a < b in input order
#[repr(C)]
#[derive(Clone, Copy, Default, Debug, PartialEq, Eq)]
pub struct kc_edge {
    pub a: u32,
    pub b: u32,
    pub count: u32,
    pub blosum: i32,
}

pub enum kc_engine {}
pub enum kc_fasta {}
pub enum kc_tree {}

extern "C" {
    pub fn kc_abi_version() -> c_int;
    pub fn kc_device_count() -> c_int;
    pub fn kc_create(cfg: *const kc_config, out: *mut *mut kc_engine) -> c_int;
    pub fn kc_destroy(e: *mut kc_engine);
    pub fn kc_last_error(e: *const kc_engine) -> *const c_char;
    pub fn kc_set_stream(e: *mut kc_engine, cuda_stream: *mut c_void) -> c_int;
    // the Vec<Protein> of src/main.rs:62-72, staged to HBM
    pub fn kc_set_proteins(e: *mut kc_engine, residues: *const u8, offsets: *const u64, class_id: *const u32, n: u64) -> c_int;
    // Protein::new + get_five_mers, src/protein.rs:107-132,141
    pub fn kc_extract_kmers(e: *mut kc_engine, out: *mut u32, cap: u64, n_positions: *mut u64) -> c_int;
    // census + split + Mphf::new x2 + rewrite + kmer_freq, src/main.rs:84-199
    pub fn kc_build_index(e: *mut kc_engine, stats: *mut kc_index_stats) -> c_int;
    pub fn kc_build_index_shard(e: *mut kc_engine, shard: u32, n_shards: u32, stats: *mut kc_index_stats) -> c_int;
    pub fn kc_index_shard_info(e: *mut kc_engine, info: *mut u32) -> c_int;
    pub fn kc_index_shard_blocks(e: *mut kc_engine, bounds: *mut u32, cap: u32) -> c_int;
    pub fn kc_index_flavour(e: *mut kc_engine) -> c_int;
    pub fn kc_get_distinct_kmers(e: *mut kc_engine, out: *mut u32, cap: u64) -> c_int;
    pub fn kc_get_vocab(e: *mut kc_engine, kmers: *mut u32, freq: *mut u32, cap: u64) -> c_int;
    pub fn kc_get_protein_ids(e: *mut kc_engine, row_offsets: *mut u64, ids: *mut u32, cap: u64) -> c_int;
    // Mphf::hash, src/main.rs:145,192
    pub fn kc_lookup_kmers(e: *mut kc_engine, kmers: *const u32, n: u64, ids: *mut u32) -> c_int;
    // Graph::new + remove_uninteresting_edges + combine_edges + threshold, src/graph/mod.rs:39-193,549-697,322-546,242
    pub fn kc_score_pairs(e: *mut kc_engine, stats: *mut kc_pair_stats) -> c_int;
    pub fn kc_score_pairs_shard(e: *mut kc_engine, shard: u32, n_shards: u32, stats: *mut kc_pair_stats) -> c_int;
    pub fn kc_get_edges(e: *mut kc_engine, out: *mut kc_edge, cap: u64) -> c_int;
    // KmerEdgeGroup.kmers, src/graph/edge.rs:49,74
    pub fn kc_get_edge_kmers(e: *mut kc_engine, edge: u64, out: *mut u32, cap: u64) -> c_int;
    // multi-GPU: NCCL lives below the ABI (csrc/dist.cuh)
    pub fn kc_comm_unique_id(id: *mut u8) -> c_int;
    pub fn kc_comm_init(e: *mut kc_engine, id: *const u8, rank: c_int, world: c_int) -> c_int;
    pub fn kc_set_proteins_dist(e: *mut kc_engine, residues: *const u8, offsets: *const u64, class_id: *const u32, n: u64) -> c_int;
    pub fn kc_build_index_dist(e: *mut kc_engine, stats: *mut kc_index_stats) -> c_int;
    pub fn kc_score_pairs_dist(e: *mut kc_engine, stats: *mut kc_pair_stats) -> c_int;
    pub fn kc_gather_edges(e: *mut kc_engine, out: *mut kc_edge, cap: u64, n_total: *mut u64) -> c_int;
    pub fn kc_gather_edges_shared(e: *mut kc_engine, shared_out: *mut kc_edge, cap: u64, n_total: *mut u64) -> c_int;
    // host side (kc_host.h): seq_io reader + class dictionary, src/main.rs:62-72, src/protein.rs:109,135-138
    pub fn kc_fasta_parse_file(path: *const c_char, threads: c_int, out: *mut *mut kc_fasta) -> c_int;
    pub fn kc_fasta_free(f: *mut kc_fasta);
    pub fn kc_fasta_n_proteins(f: *const kc_fasta) -> u64;
    pub fn kc_fasta_n_residues(f: *const kc_fasta) -> u64;
    pub fn kc_fasta_residues(f: *const kc_fasta) -> *const u8;
    pub fn kc_fasta_offsets(f: *const kc_fasta) -> *const u64;
    pub fn kc_fasta_class_ids(f: *const kc_fasta) -> *const u32;
    pub fn kc_fasta_n_missing_class(f: *const kc_fasta) -> u64;
    pub fn kc_fasta_class_name(f: *const kc_fasta, class_id: u32) -> *const c_char;
    pub fn kc_fasta_id(f: *const kc_fasta, protein: u64) -> *const c_char;
    // DIAMOND hand-off files, src/graph/mod.rs:202-220,253-261,273-280,304-317
    pub fn kc_write_handoff(f: *const kc_fasta, edges: *const kc_edge, n_edges: u64, dir: *const c_char, n_files: *mut u64) -> c_int;
    // host tree, src/tree.rs:179-385,519-536
    pub fn kc_tree_build(row_offsets: *const u64, ids: *const u32, n: u64, n_ids: u32, out: *mut *mut kc_tree) -> c_int;
    pub fn kc_tree_free(t: *mut kc_tree);
    pub fn kc_tree_clusters(t: *const kc_tree, cluster_of: *mut u32, n_clusters: *mut u32) -> c_int;
}
