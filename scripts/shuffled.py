#!/usr/bin/env python
"""The benchmark set with its rows in RANDOM order (fixed seed): what the pair stage costs when related proteins
are not neighbours in the input (the bin-local tiles of pairs_tile_kernel rely on locality; the reference's
protein order is arbitrary).  Checks the result by mapping the edges back to the generator's order
and comparing with the oracle's golden SHA-256.
    python scripts/shuffled.py [--workload synth_1m_k7] [--steps 3]
"""
import argparse
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import uniprot_kmer_based_clustering_b200 as kc  # noqa: E402
from bench import THRESHOLD, golden_for, make_set  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="synth_1m_k7")
ap.add_argument("--n-proteins", type=int, default=None)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--seed", type=int, default=0x5EED)
args = ap.parse_args()
ps, k, cross = make_set(args.workload, args.n_proteins)
n = ps.n
perm = np.random.default_rng(args.seed).permutation(n)  # new row i = old row perm[i]
lens = np.diff(ps.offsets.astype(np.int64))
new_off = np.zeros(n + 1, dtype=np.uint64)
new_off[1:] = np.cumsum(lens[perm])
# gather the residues row by row (vectorised: index = old start of the row + offset inside the row)
starts = ps.offsets[:-1].astype(np.int64)[perm]
idx = np.repeat(starts - new_off[:-1].astype(np.int64), lens[perm]) + np.arange(int(new_off[n]), dtype=np.int64)
shuf = kc.ProteinSet(ps.residues[idx], new_off, ps.class_id[perm].copy())
out = {}
for name, s in (("ordered", ps), ("shuffled", shuf)):
    with kc.Engine(k, threshold=THRESHOLD, cross_class_only=cross, want_blosum=True) as e:
        e.set_protein_set(s)
        for _ in range(2):
            e.build_index()
            e.score_pairs()
        tot = {}
        for _ in range(args.steps):
            e.reset_timings()
            ist = e.build_index()
            pst = e.score_pairs()
            for key, v in e.timings().items():
                tot[key] = tot.get(key, 0.0) + v
        edges = e.get_edges()
    avg = {key: round(v / args.steps, 3) for key, v in tot.items() if key.endswith("_ms")}
    out[name] = (avg, pst, edges)
    print(name, avg, {key: pst[key] for key in ("n_multi_edges", "n_pairs_kept", "n_edges_out", "n_rows_rescored")})
# map the shuffled edges back: row i of the shuffled set is row perm[i] of the ordered one
ed = out["shuffled"][2]
a, b = perm[ed["a"]], perm[ed["b"]]
lo, hi = np.minimum(a, b), np.maximum(a, b)
back = np.empty(ed.size, dtype=kc.EDGE_DTYPE)
back["a"], back["b"], back["count"], back["blosum"] = lo, hi, ed["count"], ed["blosum"]
back = back[np.lexsort((back["b"], back["a"]))]
same = np.array_equal(back, out["ordered"][2])
gold = golden_for(args.workload, n)
sha = hashlib.sha256(np.ascontiguousarray(back).tobytes()).hexdigest()
print("shuffled edges mapped back == ordered edges:", same, "| golden SHA:", None if gold is None else sha == gold["edges_sha256"])
assert same
