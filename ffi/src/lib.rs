//! Safe wrappers that keep the reference's module API (`protein`, `graph::Graph`) over the GPU engine.
//! Errors: every C call returns 0 or a KC_E* code; like the reference (`unwrap` / `expect` / `panic!`,
//! e.g. src/main.rs:55-63, src/graph/mod.rs:205,270,293) the wrappers panic with `kc_last_error`.
pub mod kc_sys;

use kc_sys::*;
use std::ffi::{CStr, CString};
use std::ptr;

fn check(e: *mut kc_engine, rc: i32, what: &str) {
    if rc != KC_OK {
        let msg = if e.is_null() { "".into() } else { unsafe { CStr::from_ptr(kc_last_error(e)) }.to_string_lossy().into_owned() };
        panic!("{what} failed ({rc}): {msg}");
    }
}

/// The staged input: what `Vec<Protein>` holds after src/main.rs:62-72.
pub struct Fasta(*mut kc_fasta);
impl Fasta {
    pub fn from_path(path: &str, threads: u32) -> Fasta {
        let c = CString::new(path).unwrap();
        let mut h = ptr::null_mut();
        let rc = unsafe { kc_fasta_parse_file(c.as_ptr(), threads as i32, &mut h) };
        if rc != KC_OK { panic!("input argument should refer to an existing fasta file"); } // src/main.rs:62-63
        Fasta(h)
    }
    pub fn len(&self) -> usize { unsafe { kc_fasta_n_proteins(self.0) as usize } }
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    pub fn id(&self, p: usize) -> String { unsafe { CStr::from_ptr(kc_fasta_id(self.0, p as u64)) }.to_string_lossy().into_owned() }
    pub fn seq(&self, p: usize) -> String {
        unsafe {
            let off = std::slice::from_raw_parts(kc_fasta_offsets(self.0), self.len() + 1);
            let res = std::slice::from_raw_parts(kc_fasta_residues(self.0), off[self.len()] as usize);
            String::from_utf8_lossy(&res[off[p] as usize..off[p + 1] as usize]).into_owned()
        }
    }
    pub fn raw(&self) -> *mut kc_fasta { self.0 }
}
impl Drop for Fasta { fn drop(&mut self) { unsafe { kc_fasta_free(self.0) } } }

/// One engine per GPU.
pub struct Engine { h: *mut kc_engine, pub k: i32 }
impl Engine {
    pub fn new(cfg: kc_config) -> Engine {
        let mut h = ptr::null_mut();
        let rc = unsafe { kc_create(&cfg, &mut h) };
        if rc != KC_OK { panic!("kc_create failed ({rc}): no usable CUDA device (there is no CPU fallback)"); }
        Engine { h, k: cfg.k }
    }
    /// The engine BORROWS the staged arrays of `fa` (asynchronous upload, host copies made inside the next build):
    /// keep `fa` alive and unchanged until `build_index` has returned (include/kc_b200.h, kc_set_proteins).
    pub fn set_proteins(&mut self, fa: &Fasta) {
        let rc = unsafe { kc_set_proteins(self.h, kc_fasta_residues(fa.0), kc_fasta_offsets(fa.0), kc_fasta_class_ids(fa.0), fa.len() as u64) };
        check(self.h, rc, "kc_set_proteins");
    }
    /// census + split + index + rewrite (src/main.rs:84-199)
    pub fn build_index(&mut self) -> kc_index_stats {
        let mut st = kc_index_stats::default();
        let rc = unsafe { kc_build_index(self.h, &mut st) };
        check(self.h, rc, "kc_build_index");
        st
    }
    /// five_mer_repeat + kmer_freq (src/main.rs:135,187-193)
    pub fn vocab(&mut self, n_repeated: u64) -> (Vec<u32>, Vec<u32>) {
        let (mut v, mut f) = (vec![0u32; n_repeated as usize], vec![0u32; n_repeated as usize]);
        let rc = unsafe { kc_get_vocab(self.h, v.as_mut_ptr(), f.as_mut_ptr(), n_repeated) };
        check(self.h, rc, "kc_get_vocab");
        (v, f)
    }
    pub fn raw(&self) -> *mut kc_engine { self.h }
}
impl Drop for Engine { fn drop(&mut self) { unsafe { kc_destroy(self.h) } } }

pub mod graph {
    use super::*;
    /// Same public surface as the reference's Graph (src/graph/mod.rs:31-40,195,322,549).
    pub struct Graph<'a> { engine: &'a mut Engine, fasta: &'a Fasta, stats: kc_pair_stats, pub edges: Vec<kc_edge> }

    impl<'a> Graph<'a> {
        /// Graph::new(kmer_freq, thread_count, protein_list): all pairs scored in one pass
        pub fn new(kmer_freq: &[u32], _thread_count: usize, engine: &'a mut Engine, fasta: &'a Fasta) -> Graph<'a> {
            let mut stats = kc_pair_stats::default();
            let rc = unsafe { kc_score_pairs(engine.raw(), &mut stats) };
            check(engine.raw(), rc, "kc_score_pairs");
            eprintln!("Number of {}mers found in at least two proteins: {}", engine.k, kmer_freq.len()); // :50
            eprintln!("Number of total edges: {}", stats.n_multi_edges);                                 // :51
            let mut edges = vec![kc_edge::default(); stats.n_edges_out as usize];
            let rc = unsafe { kc_get_edges(engine.raw(), edges.as_mut_ptr(), edges.len() as u64) };
            check(engine.raw(), rc, "kc_get_edges");
            Graph { engine, fasta, stats, edges }
        }
        pub fn remove_uninteresting_edges(&mut self, _thread_count: u32) {          // :549
            eprintln!("Remove edges without diverging AMR labels");
            eprintln!("Number of edges now: {}", self.stats.n_multi_edges_kept);    // :695
        }
        pub fn combine_edges(&mut self, _thread_count: u32) {                       // :322
            eprintln!("Combine edges with the same two vertices");
            eprintln!("Number of edges now: {}", self.stats.n_pairs_kept);          // :545
        }
        /// :195-319; every edge already has more than `threshold` k-mers in common (:242).  Writes the files
        /// the reference leaves for DIAMOND; the makedb / blastp subprocesses (:266-293) attach unchanged.
        pub fn align_and_output_pairs(&self, _thread_count: u32) {
            for e in &self.edges {
                eprintln!("Cross-checking:\n\treference protein:{}\n\tquery protein:{}\n\tkmers in common:{}",
                          self.fasta.id(e.a as usize), self.fasta.id(e.b as usize), e.count);     // :250-251
            }
            let dir = CString::new(".").unwrap();
            let mut n_files = 0u64;
            let rc = unsafe { kc_write_handoff(self.fasta.raw(), self.edges.as_ptr(), self.edges.len() as u64, dir.as_ptr(), &mut n_files) };
            check(self.engine.raw(), rc, "kc_write_handoff");
        }
        /// KmerEdge::get_kmers as k-mer VALUES (src/graph/edge.rs:119-131)
        pub fn edge_kmers(&mut self, i: usize) -> Vec<u32> {
            let mut out = vec![0u32; self.edges[i].count as usize];
            let rc = unsafe { kc_get_edge_kmers(self.engine.raw(), i as u64, out.as_mut_ptr(), out.len() as u64) };
            check(self.engine.raw(), rc, "kc_get_edge_kmers");
            out
        }
    }
}
