"""Regenerates tests/golden/arg_golden.json with an independent numpy/scipy restatement
(NOT the C++ oracle, NOT the CUDA engine) of the reference's hot path, from the ARG
protein set the reference bundles (uniprot_arg.fasta; committed here xz-compressed as a
data fixture because /root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py            # reads tests/golden/arg_proteins.fasta.xz

The values reproduce SURVEY.md §8c; the oracle and the CUDA engine are tested against
the JSON this script writes.  Reference lines restated: src/protein.rs:9-13,29-37,49-54,
107-138; src/main.rs:100-137; src/graph/mod.rs:44-51,242,545,580-587,695.
"""
import hashlib
import json
import lzma
import os

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ALPHABET = "CSTAGPDEQNHRKMILVWYF*"
DIAG = [9, 4, 5, 4, 6, 7, 6, 5, 5, 6, 8, 5, 5, 5, 4, 4, 4, 11, 7, 6, 0]


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load():
    text = lzma.open(os.path.join(HERE, "arg_proteins.fasta.xz")).read().decode()
    ids, seqs = [], []
    for line in text.splitlines():
        if line.startswith(">"):
            ids.append(line[1:].split()[0])
            seqs.append("")
        elif line:
            seqs[-1] += line
    return ids, seqs


def main():
    ids, seqs = load()
    lut = np.full(256, 20, dtype=np.int64)
    for i, c in enumerate(ALPHABET):
        lut[ord(c)] = i
    names = [i.split("|")[3] for i in ids]
    table = {}
    cls = np.array([table.setdefault(nm, len(table)) for nm in names])
    out = {"n_proteins": len(ids), "n_residues": sum(map(len, seqs)), "n_classes": len(table),
           "protein0_id": ids[0], "protein0_len": len(seqs[0])}
    for k in (5, 7):
        rows, cols, allk = [], [], []
        for p, s in enumerate(seqs):
            codes = lut[np.frombuffer(s.encode(), dtype=np.uint8)]
            if len(codes) < k:
                continue
            km = np.zeros(len(codes) - k + 1, dtype=np.int64)
            for j in range(k):
                km = km * 21 + codes[j:len(codes) - k + 1 + j]
            allk.append(km)
            u = np.unique(km)
            rows.append(np.full(u.size, p))
            cols.append(u)
        flat = np.concatenate(allk).astype(np.uint32)
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        distinct, counts = np.unique(cols, return_counts=True)
        rep = distinct[counts > 1]
        keep = np.isin(cols, rep)
        rid = np.searchsorted(rep, cols[keep])
        A = sp.csr_matrix((np.ones(rid.size, dtype=np.int64), (rows[keep], rid)),
                          shape=(len(ids), rep.size))
        f = np.asarray(A.sum(axis=0)).ravel()
        ss = np.zeros(rep.size, dtype=np.int64)
        t = rep.copy()
        for _ in range(k):
            ss += np.array(DIAG)[t % 21]
            t //= 21
        S = sp.triu(A @ A.T, k=1).tocoo()
        B = sp.triu((A.multiply(ss[None, :]).tocsr()) @ A.T, k=1).tocsr()
        a, b, c = S.row, S.col, S.data
        order = np.lexsort((b, a))
        a, b, c = a[order], b[order], c[order]
        cross = cls[a] != cls[b]
        g = {
            "first_kmers_protein0": [int(x) for x in allk[0][:6]], "last_kmer_protein0": int(allk[0][-1]),
            "n_positions": int(flat.size), "sum_positions": int(flat.astype(np.uint64).sum()),
            "xor_positions": int(np.bitwise_xor.reduce(flat)),
            "n_incidences": int(cols.size), "n_distinct": int(distinct.size),
            "n_singleton": int((counts == 1).sum()), "n_repeated": int(rep.size),
            "sha_distinct": sha16(distinct.astype("<u4")), "sha_repeated": sha16(rep.astype("<u4")),
            "repeated_min": int(rep.min()), "repeated_max": int(rep.max()),
            "top3": [[int(rep[i]), int(f[i])] for i in np.argsort(-f, kind="stable")[:3]],
            "nnz": int(A.nnz), "n_multi_edges": int((f * (f - 1) // 2).sum()),
            "n_pairs_all": int(a.size),
        }
        for name, m, thr in (("cross_gt10", cross & (c > 10), 10), ("cross_gt0", cross, 0),
                             ("all_gt10", c > 10, 10)):
            e = np.stack([a[m], b[m], c[m]], axis=1).astype("<u4")
            bl = np.asarray(B[a[m], b[m]]).ravel()
            g[name] = {"n": int(e.shape[0]), "sum_count": int(c[m].sum()), "sha": sha16(e),
                       "first3": e[:3].tolist(), "max_count": int(c[m].max()),
                       "blosum_sum": int(bl.sum()), "blosum_first3": [int(x) for x in bl[:3]]}
        g["n_multi_edges_cross"] = int(c[cross].sum())
        g["n_pairs_cross"] = int(cross.sum())
        out[f"k{k}"] = g
    with open(os.path.join(HERE, "arg_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out)[:400])


if __name__ == "__main__":
    main()
