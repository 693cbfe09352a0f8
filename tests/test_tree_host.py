"""Host tree clustering (csrc/tree.cpp, the reference's src/tree.rs) against the literal Python
model (oracle/tree_model.py).  Host-only: runs without a GPU, the id lists come from the oracle."""
import numpy as np
import pytest

import uniprot_kmer_based_clustering_b200 as kc
from conftest import random_protein_set
from oracle import tree_model as tm
from oracle.oracle import Oracle
from uniprot_kmer_based_clustering_b200.tree import Tree


def _index(ps, k):
    o = Oracle(k, 2)
    o.set_proteins(ps.residues, ps.offsets, ps.class_id)
    o.extract_kmers()
    return o.build_index()


def _check(ix, n):
    rows = [ix.ids[int(ix.row_offsets[p]):int(ix.row_offsets[p + 1])] for p in range(n)]
    model = tm.build_tree(rows)
    tree = Tree.from_id_rows(ix.row_offsets, ix.ids, ix.stats["n_repeated"])
    assert tree.nested() == tm.nested(model.root)
    assert tree.n_merges == model.log.count("Merging")
    assert tree.n_no_common == model.log.count("No kmers in common")
    cl = tree.clusters()
    exp = tm.clusters(model)
    assert tree.n_clusters == len(exp)
    for ci, members in enumerate(exp):
        assert np.all(cl[members] == ci)
    return tree, model


@pytest.mark.parametrize("k,seed", [(5, 0), (5, 1), (7, 2), (7, 3)])
def test_tree_matches_literal_model_on_random_sets(k, seed):
    ps = random_protein_set(seed, 160, min_len=0, max_len=120, n_classes=3, family=6, mutate=0.05)
    tree, model = _check(_index(ps, k), ps.n)
    assert tree.n_merges > 10


@pytest.mark.parametrize("k", [5, 7])
def test_tree_matches_literal_model_on_arg_subset(k, arg_set):
    n = 400
    sub = kc.ProteinSet(arg_set.residues[:int(arg_set.offsets[n])], arg_set.offsets[:n + 1], arg_set.class_id[:n])
    tree, model = _check(_index(sub, k), n)
    sizes = np.bincount(tree.clusters())
    assert sizes.sum() == n and tree.n_clusters >= 2


def test_tree_degenerate_inputs():
    # one protein: a single leaf
    t = Tree.from_id_rows(np.array([0, 3], np.uint64), np.array([1, 4, 7], np.uint32), 10)
    assert t.nested() == 0 and t.clusters().tolist() == [0] and t.n_clusters == 1
    # proteins without any repeated k-mer never merge: "No kmers in common" each time
    ro = np.array([0, 0, 0, 0, 0], np.uint64)
    t = Tree.from_id_rows(ro, np.zeros(0, np.uint32), 5)
    assert t.nested() == [0, 1, 2, 3] and t.n_no_common == 2 and t.n_merges == 0
    # identical proteins: every similarity is equal, so max == min and nothing merges
    ro = np.array([0, 3, 6, 9, 12], np.uint64)
    t = Tree.from_id_rows(ro, np.tile(np.array([0, 2, 5], np.uint32), 4), 6)
    assert t.nested() == [0, 1, 2, 3] and t.n_merges == 0
    # unsorted rows or out-of-range ids are rejected
    with pytest.raises(kc.KcError):
        Tree.from_id_rows(np.array([0, 2], np.uint64), np.array([3, 1], np.uint32), 10)
    with pytest.raises(kc.KcError):
        Tree.from_id_rows(np.array([0, 1], np.uint64), np.array([10], np.uint32), 10)


@pytest.mark.parametrize("k", [5, 7])
def test_tree_on_full_arg_set_matches_golden(k, arg_oracle):
    """tests/golden/tree_golden.json comes from the literal Python model on all 10 619 proteins
    (tests/golden/make_tree_golden.py, minutes); the C++ tree reproduces it in ~2 s"""
    import hashlib
    import json
    import os
    from conftest import GOLDEN_DIR
    g = json.load(open(os.path.join(GOLDEN_DIR, "tree_golden.json")))[f"k{k}"]
    ix = arg_oracle[k][2]
    tree = Tree.from_id_rows(ix.row_offsets, ix.ids, ix.stats["n_repeated"])
    toks = tree.serialize()
    assert toks.size == g["n_tokens"]
    assert hashlib.sha256(toks.astype("<i8").tobytes()).hexdigest()[:16] == g["sha_tokens"]
    cl = tree.clusters()
    assert tree.n_clusters == g["n_clusters"] and tree.n_merges == g["n_merges"]
    assert tree.n_no_common == g["n_no_common"]
    assert np.sort(np.bincount(cl))[::-1][:8].tolist() == g["largest_clusters"]
