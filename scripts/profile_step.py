#!/usr/bin/env python
"""One profiled step of the hot path (for ncu): warm step outside the profiler range, then
cudaProfilerStart .. one build_index + score_pairs .. cudaProfilerStop.

    ncu --profile-from-start off ... python scripts/profile_step.py --workload synth_1m_k7
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import uniprot_kmer_based_clustering_b200 as kc  # noqa: E402
from bench import THRESHOLD, make_set  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="synth_1m_k7")
ap.add_argument("--n-proteins", type=int, default=None)
ap.add_argument("--cross", action="store_true")
ap.add_argument("--index", default="auto", choices=["auto", "stream", "bucket", "table"])
args = ap.parse_args()
ps, k, cross = make_set(args.workload, args.n_proteins)
cross = cross or args.cross
with kc.Engine(k, threshold=THRESHOLD, cross_class_only=cross, want_blosum=True, index_build=args.index) as e:
    e.set_protein_set(ps)
    e.build_index()
    e.score_pairs()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ist = e.build_index()
    pst = e.score_pairs()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(ist, pst, e.timings())
