"""Regenerates tests/golden/tree_golden.json: the LITERAL Python model of the reference's tree
clustering (oracle/tree_model.py, src/tree.rs restated) run on the full ARG protein set with the
oracle's id lists.  Takes ~1.5 min (k=5) + ~2 min (k=7); the C++ host tree must reproduce the
serialised tree bit for bit (tests/test_tree_host.py).

    python tests/golden/make_tree_golden.py
"""
import hashlib
import json
import lzma
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.setrecursionlimit(100000)

from oracle import tree_model as tm  # noqa: E402
from oracle.fasta_ref import parse_fasta_bytes  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    fa = parse_fasta_bytes(lzma.open(os.path.join(HERE, "arg_proteins.fasta.xz")).read())
    out = {}
    for k in (5, 7):
        o = Oracle(k, 4)
        o.set_proteins(fa["residues"], fa["offsets"], fa["class_id"])
        o.extract_kmers()
        ix = o.build_index()
        rows = [ix.ids[int(ix.row_offsets[p]):int(ix.row_offsets[p + 1])] for p in range(len(fa["ids"]))]
        t = tm.build_tree(rows)
        toks = []

        def ser(nd):
            if not nd.children:
                toks.append(nd.protein)
                return
            toks.append(-len(nd.children))
            for c in nd.children:
                ser(c)

        ser(t.root)
        sizes = sorted((len(c) for c in tm.clusters(t)), reverse=True)
        out[f"k{k}"] = {"n_clusters": len(t.root.children), "n_merges": t.log.count("Merging"),
                        "n_no_common": t.log.count("No kmers in common"), "n_tokens": len(toks),
                        "sha_tokens": hashlib.sha256(np.array(toks, dtype="<i8").tobytes()).hexdigest()[:16],
                        "largest_clusters": sizes[:8]}
        print(k, out[f"k{k}"])
    with open(os.path.join(HERE, "tree_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
