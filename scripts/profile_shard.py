import os, sys
sys.path.insert(0, "/root/repo")
import torch
import uniprot_kmer_based_clustering_b200 as kc
from bench import THRESHOLD, make_set
ps, k, cross = make_set("synth_1m_k7", None)
sh, ns = int(sys.argv[1]), int(sys.argv[2])
with kc.Engine(k, threshold=THRESHOLD, cross_class_only=cross, want_blosum=True) as e:
    e.set_protein_set(ps)
    for _ in range(2):
        e.build_index(sh, ns); e.score_pairs(sh, ns)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    e.build_index(sh, ns); e.score_pairs(sh, ns)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
