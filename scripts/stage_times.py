#!/usr/bin/env python
"""Stage timings (CUDA events inside the engine) of a few steps of one workload, inputs resident in
HBM.  For quick A/B runs of engine variants (environment switches), not a bench line.

    python scripts/stage_times.py --workload synth_1m_k7 [--steps 3] [--n-proteins N]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import uniprot_kmer_based_clustering_b200 as kc  # noqa: E402
from bench import THRESHOLD, make_set  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="synth_1m_k7")
ap.add_argument("--n-proteins", type=int, default=None)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--cross", action="store_true")
ap.add_argument("--ballots", action="store_true")
ap.add_argument("--index", default="auto", choices=["auto", "stream", "bucket", "table"])
ap.add_argument("--no-blosum", action="store_true")
ap.add_argument("--shard", type=int, default=0)
ap.add_argument("--n-shards", type=int, default=1)
args = ap.parse_args()
ps, k, cross = make_set(args.workload, args.n_proteins)
cross = cross or args.cross
with kc.Engine(k, threshold=THRESHOLD, cross_class_only=cross, want_blosum=not args.no_blosum, index_build=args.index, census_merge=77 if args.ballots else 0) as e:
    e.set_protein_set(ps)
    for _ in range(2):
        ist = e.build_index(args.shard, args.n_shards)
        pst = e.score_pairs(args.shard, args.n_shards)
    torch.cuda.synchronize()
    tot = {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e.reset_timings()
        ist = e.build_index(args.shard, args.n_shards)
        pst = e.score_pairs(args.shard, args.n_shards)
        for key, v in e.timings().items():
            tot[key] = tot.get(key, 0.0) + v
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    avg = {key: round(v / args.steps, 3) for key, v in tot.items()}
    print(args.workload, f"shard {args.shard}/{args.n_shards}", args.index, f"wall {wall:.2f} ms/step", avg)
    print("  index", ist)
    print("  pairs", pst)
