//! `cargo run --release -- <fasta> <threads>`: the reference's main (src/main.rs:50-239) over the engine.
use kc_b200::graph::Graph;
use kc_b200::kc_sys::kc_config;
use kc_b200::{Engine, Fasta};
use std::env;

fn main() {
    eprintln!("We start main");
    let args: Vec<String> = env::args().collect();
    if args.len() != 3 { panic!("Requires two command line arguments: input and thread"); }   // :54-57
    let input = &args[1];
    let threads: u32 = args[2].parse().expect("threads argument should be of type int");       // :59-60
    let fasta = Fasta::from_path(input, threads);                                               // :62-72
    eprintln!("We created Protein structs");
    let mut engine = Engine::new(kc_config { k: 5, device: 0, threshold: 10, cross_class_only: 1, ..Default::default() });
    engine.set_proteins(&fasta);
    let stats = engine.build_index();                                                           // :84-199
    eprintln!("We combined k-mers\nWe found unique k-mers\nWe made unique hash\nWe can make a graph");
    let (_repeat, kmer_freq) = engine.vocab(stats.n_repeated);
    let mut graph = Graph::new(&kmer_freq, threads as usize, &mut engine, &fasta);              // :216-218
    graph.remove_uninteresting_edges(threads);                                                  // :224
    graph.combine_edges(threads);                                                               // :226
    graph.align_and_output_pairs(threads);                                                      // :232
    println!("Graph right now: {} edges over the threshold", graph.edges.len());                // (:235 dumps the graph)
}
