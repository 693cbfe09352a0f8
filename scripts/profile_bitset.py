#!/usr/bin/env python
"""The bitset (AND + POPC) path against the measured POPC issue rate: microbenchmark, then kc_bitset_pair_counts
on the first N proteins of the synthetic 20 k set (k = 5), timed and checked against the sparse counts.
    python scripts/profile_bitset.py [n_rows]
    ncu --metrics regex:'smsp__inst_executed_pipe_.*sum$',gpu__time_duration.sum -k regex:'bitset_pairs|popc_micro|pairs_tile' python scripts/profile_bitset.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import uniprot_kmer_based_clustering_b200 as kc  # noqa: E402

n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ps = kc.ProteinSet.synthetic(20000, "A", 0xB2000003, threads=8)
with kc.Engine(5, threshold=10, cross_class_only=False, want_blosum=True) as e:
    peak = max(e.popc_microbench(1 << 14) for _ in range(3))
    e.set_protein_set(ps)
    ist = e.build_index()
    pst = e.score_pairs()  # (runs pairs_tile_kernel too: captured by the same ncu command)
    rows = np.arange(n_rows, dtype=np.uint32)
    e.bitset_pair_counts(rows[:256])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    counts = e.bitset_pair_counts(rows)
    dt = time.perf_counter() - t0
    words = (ist["n_repeated"] + 31) // 32 + 1
    ops = float(n_rows) * n_rows * words
    edges = e.get_edges()
    m = (edges["a"] < n_rows) & (edges["b"] < n_rows)
    assert np.array_equal(counts[edges["a"][m], edges["b"][m]], edges["count"][m])
    print(f"POPC microbenchmark: {peak:.1f} G(AND+POPC+ADD)/s = {peak / 148 / 1.965:.2f} per clock per SM at 1.965 GHz")
    print(f"bitset_pairs: {n_rows} x {n_rows} rows x {words} words = {ops:.3e} word ops; call {dt * 1e3:.1f} ms wall "
          f"(includes the fill, the D2H of {n_rows * n_rows * 4 / 1e6:.0f} MB of counts); see ncu for the kernel time")
