#!/usr/bin/env python
"""End-to-end step from pinned HOST buffers (upload + index + pairs + edge readback), host wall clock per call,
for the upload modes of kc_config.no_upload_overlap (0 = chunked + one level-1 pass per chunk, 1 = one upload).
A/B tool, not a bench line.

    python scripts/e2e_times.py [--workload synth_1m_k7] [--steps 5]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import uniprot_kmer_based_clustering_b200 as kc  # noqa: E402
from bench import THRESHOLD, make_set  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="synth_1m_k7")
ap.add_argument("--n-proteins", type=int, default=None)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
ps, k, cross = make_set(args.workload, args.n_proteins)
h_res = torch.from_numpy(ps.residues).pin_memory()
h_off = torch.from_numpy(ps.offsets.view(np.int64)).pin_memory()
h_cls = torch.from_numpy(ps.class_id.view(np.int32)).pin_memory()
h_edges = torch.empty(16_000_000 * 4, dtype=torch.int32).pin_memory()
for mode in (0, 1, 0, 1):
    with kc.Engine(k, threshold=THRESHOLD, cross_class_only=cross, want_blosum=True, no_upload_overlap=mode) as e:
        parts = np.zeros(4)
        tim = {}
        for it in range(args.steps + 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e.set_proteins_ptr(h_res.data_ptr(), h_off.data_ptr(), h_cls.data_ptr(), ps.n, on_device=False)
            t1 = time.perf_counter()
            e.reset_timings()
            ist = e.build_index()
            t2 = time.perf_counter()
            pst = e.score_pairs()
            t3 = time.perf_counter()
            e.get_edges_into(h_edges.data_ptr(), 16_000_000)
            t4 = time.perf_counter()
            if it >= 2:
                parts += np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3]) * 1e3
                for key, v in e.timings().items():
                    tim[key] = tim.get(key, 0.0) + v
        parts /= args.steps
        print(f"no_upload_overlap={mode}: e2e {parts.sum():.2f} ms = set_proteins {parts[0]:.2f} + build_index {parts[1]:.2f} "
              f"+ score_pairs {parts[2]:.2f} + edges {parts[3]:.2f};  engine events: index {tim['index_ms'] / args.steps:.2f} "
              f"partition {tim['census_kernel_ms'] / args.steps:.2f} h2d {tim['h2d_ms'] / args.steps:.2f}")
