"""B200-native k-mer clustering hot path (drop-in for Isabella136/uniprot_kmer_based_clustering).

The product is the C-ABI shared library `lib/libkc_b200.so` (include/kc_b200.h): hand-written
sm_100a CUDA kernels for k-mer extraction, the perfect k-mer index, all-pairs shared-k-mer
scoring and edge emission.  This package is the thin Python host side over that ABI:

* `Engine`, `ProteinSet`, `cluster`  — plumbing (engine.py)
* `protein`, `graph`, `tree`         — mirror of the reference's Rust module API (same
                                       names, argument meaning and error behaviour)
"""
from ._lib import KcError, LIB_PATH, build, lib  # noqa: F401
from .engine import EDGE_DTYPE, SYNTH_SEEDS, Engine, ProteinSet, cluster  # noqa: F401
from . import graph, protein, tree  # noqa: F401

__all__ = ["Engine", "ProteinSet", "cluster", "KcError", "EDGE_DTYPE", "SYNTH_SEEDS", "build", "lib",
           "LIB_PATH", "protein", "graph", "tree"]
