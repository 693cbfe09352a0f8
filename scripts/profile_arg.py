#!/usr/bin/env python
"""One profiled step on the ARG protein set (tests/golden/arg_proteins.fasta.xz): python scripts/profile_arg.py K [cross]"""
import lzma, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import uniprot_kmer_based_clustering_b200 as kc
k = int(sys.argv[1]); cross = len(sys.argv) < 3 or sys.argv[2] != "all"
ps = kc.ProteinSet.from_fasta_bytes(lzma.open(os.path.join(ROOT, "tests", "golden", "arg_proteins.fasta.xz")).read(), 8)
with kc.Engine(k, threshold=10, cross_class_only=cross, want_blosum=True) as e:
    e.set_protein_set(ps)
    for _ in range(3):
        e.build_index(); e.score_pairs()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    e.reset_timings()
    ist = e.build_index(); pst = e.score_pairs()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(ist, pst, e.timings())
