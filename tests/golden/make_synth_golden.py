#!/usr/bin/env python
"""Full-size golden values of the synthetic benchmark configurations, produced by the CPU oracle.

    python tests/golden/make_synth_golden.py [workload ...]    # default: synth_100k_k5 synth_1m_k7

Writes tests/golden/synth_golden.json: for every workload the index counters, the pair counters
(the numbers the reference prints at src/graph/mod.rs:50-51, :695, :545, :242) and the SHA-256 of the
sorted (a, b, count, blosum) edge list as little-endian u32/i32.  The generator (G1, include/kc_host.h)
is all-integer, so the same protein set is rebuilt from (n, law, seed) on any box.
The GPU tests (`-m gpu`) and bench.py (at every N) assert these values: a gathered multi-GPU list is
thereby proven identical to the single-GPU one and to the oracle's.
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (n, law, k, seed, cross_class_only)   — the same table as bench.py
    "synth_1m_k7": (1_000_000, "A", 7, 0xB2000004, False),
    "synth_100k_k5": (100_000, "A", 5, 0xB2000003, False),
    "synth_20k_k5": (20_000, "A", 5, 0xB2000003, False),
    "synth_250k_skew_k7": (250_000, "B", 7, 0xB2000005, False),
}
THRESHOLD = 10


def edge_sha(edges: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(edges).tobytes()).hexdigest()


def main():
    import uniprot_kmer_based_clustering_b200 as kc
    from oracle.oracle import Oracle
    names = sys.argv[1:] or ["synth_100k_k5", "synth_1m_k7"]
    path = os.path.join(ROOT, "tests", "golden", "synth_golden.json")
    out = {}
    if os.path.exists(path):
        with open(path) as fh:
            out = json.load(fh)
    for name in names:
        n, law, k, seed, cross = WORKLOADS[name]
        ps = kc.ProteinSet.synthetic(n, law, seed, threads=os.cpu_count() or 8)
        o = Oracle(k, os.cpu_count() or 8)
        o.set_proteins(ps.residues, ps.offsets, ps.class_id)
        t0 = time.time()
        o.extract_kmers()
        ix = o.build_index()
        pr = o.score_pairs(THRESHOLD, cross, True, mode=1)
        out[name] = {
            "n": n, "law": law, "k": k, "seed": hex(seed), "cross_class_only": cross, "threshold": THRESHOLD,
            "residues_sha256": hashlib.sha256(ps.residues.tobytes()).hexdigest(),
            "index": ix.stats, "pairs": pr.stats,
            "edges_sha256": edge_sha(pr.edges),
            "sum_blosum": int(pr.edges["blosum"].astype(np.int64).sum()),
            "first_edges": [[int(x) for x in e] for e in pr.edges[:3]],
            "oracle_seconds": round(time.time() - t0, 1),
        }
        print(name, json.dumps(out[name]), flush=True)
        with open(path, "w") as fh:
            json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
