#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck  python scripts/sanitize_target.py
    compute-sanitizer --tool racecheck python scripts/sanitize_target.py
The smoke workload (2 000 synthetic proteins, k = 5 and 7, cross-class and all-classes) through every index
build (streaming with both bucket capacities, round 1's partitioned build, the table build), a sharded build
(3 shards, every shard in turn) and the read-back entry points."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import uniprot_kmer_based_clustering_b200 as kc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
tot = 0
for k, seed in ((5, 0xB2000003), (7, 0xB2000004)):
    ps = kc.ProteinSet.synthetic(n, "A", seed, threads=4)
    for cross in (False, True):
        ref = None
        for build, cap in (("stream", 0), ("stream", 512), ("bucket", 0), ("table", 0)):
            with kc.Engine(k, threshold=10, cross_class_only=cross, want_blosum=True, index_build=build,
                           bucket_cap=cap) as e:
                e.set_protein_set(ps)
                e.extract_kmers()
                ist = e.build_index()
                pst = e.score_pairs()
                ed = e.get_edges()
                if ref is None:
                    ref = ed
                assert np.array_equal(ed, ref), (k, cross, build, cap)
                if build == "stream" and cap == 0:
                    e.get_vocab()
                    e.get_pair_index()
                    if ed.size:
                        e.get_edge_kmers(0, int(ed[0]["count"]))
                    parts = []
                    for s in range(3):
                        e.build_index(s, 3)
                        e.score_pairs(s, 3)
                        parts.append(e.get_edges())
                    allp = np.concatenate(parts)
                    allp = allp[np.lexsort((allp["b"], allp["a"]))]
                    assert np.array_equal(allp, ref), "sharded build differs"
                tot += ed.size
print("sanitize target ok:", tot, "edges")
