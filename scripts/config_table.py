#!/usr/bin/env python
"""one row per BASELINE.json config that fits one GPU (stage timings from the
engine's CUDA events, inputs resident in HBM).  The CPU column is bench.py's cpu_baseline leg
(`bench.oracle_run`, the one sanctioned place outside tests/ that executes oracle/); bit-exactness
is the business of tests/ (`-m gpu`), not of this script.

    python scripts/config_table.py [--skip-oracle-above N]
"""
import argparse
import lzma
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402
import uniprot_kmer_based_clustering_b200 as kc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--skip-oracle-above", type=int, default=300_000)
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
arg = kc.ProteinSet.from_fasta_bytes(lzma.open(os.path.join(ROOT, "tests", "golden", "arg_proteins.fasta.xz")).read(), 8)
configs = [
    ("1 ARG k=5 (cross-class, >10)", arg, 5, True, False),
    ("2 ARG k=7 + BLOSUM (cross-class, >10)", arg, 7, True, True),
    ("3 synth 100 k, k=5 (all pairs, >10, BLOSUM)", kc.ProteinSet.synthetic(100_000, "A", 0xB2000003, threads=16), 5, False, True),
    ("4 synth 1 M, k=7 (all pairs, >10, BLOSUM)", kc.ProteinSet.synthetic(1_000_000, "A", 0xB2000004, threads=16), 7, False, True),
]
threads = os.cpu_count() or 1
print("| Config | GPUs | step ms (index + pairs + edges) | k-mers indexed/s | pairs scored/s | multi-edges/s | CPU restatement "
      f"({threads} threads) |")
print("|---|---|---|---|---|---|---|")
for name, ps, k, cross, blosum in configs:
    with kc.Engine(k, threshold=10, cross_class_only=cross, want_blosum=blosum) as e:
        e.set_protein_set(ps)
        for _ in range(2):
            e.build_index()
            e.score_pairs()
        tot = {}
        steps = 5
        for _ in range(steps):
            e.reset_timings()
            ist = e.build_index()
            pst = e.score_pairs()
            for key, v in e.timings().items():
                tot[key] = tot.get(key, 0.0) + v / steps
        edges = e.get_edges()
    n = ps.n
    pairs = n * (n - 1) // 2
    step = tot["index_ms"] + tot["pairs_ms"] + tot["edges_ms"]
    cpu = "not run (sample in bench.py)"
    if n <= args.skip_oracle_above:
        bench.THRESHOLD = 10
        r = bench.oracle_run(ps, k, cross, n, threads)
        cpu = (f"{r['total_s'] * 1e3:.0f} ms ({r['positions'] / r['index_s']:.3g} k-mers/s, "
               f"{r['pairs'] / r['pairs_s']:.3g} pairs/s)")
    print(f"| {name} | 1 | {step:.2f} ({tot['index_ms']:.2f} + {tot['pairs_ms']:.2f} + {tot['edges_ms']:.2f}) | "
          f"{ist['n_positions'] / (tot['index_ms'] * 1e-3):.3g} | {pairs / ((tot['pairs_ms'] + tot['edges_ms']) * 1e-3):.3g} | "
          f"{pst['n_multi_edges_kept'] / (tot['pair_kernel_ms'] * 1e-3):.3g} | {cpu} |", flush=True)
