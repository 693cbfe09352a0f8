"""GPU parity suite: the CUDA engine, called through the C ABI, against the CPU oracle and the
committed golden vectors.  Integer work: every comparison is bit-exact."""
import numpy as np
import pytest

import uniprot_kmer_based_clustering_b200 as kc
from conftest import edges_abc, random_protein_set, sha16
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["stream", "stream512", "bucket", "table"])
def index_flavour(request, monkeypatch):
    """every test runs against every index build (kc_config.index_build): the streaming partitioned one
    (csrc/stream_index.cuh; "stream512" = with 512-record shared-memory buckets, so that ordinary buckets
    take the global-memory path of oversized ones), round 1's partitioned one (csrc/bucket.cuh) and the
    universe-table one (csrc/index.cuh)"""
    from uniprot_kmer_based_clustering_b200 import engine as eng
    monkeypatch.setitem(eng.DEFAULTS, "index_build", "stream" if request.param == "stream512" else request.param)
    monkeypatch.setitem(eng.DEFAULTS, "bucket_cap", 512 if request.param == "stream512" else 0)
    return request.param


def run_oracle(ps, k, thr, cross, blosum=True, threads=8):
    o = Oracle(k, threads)
    o.set_proteins(ps.residues, ps.offsets, ps.class_id)
    km = o.extract_kmers()
    ix = o.build_index()
    pr = o.score_pairs(thr, cross, blosum, mode=1)
    return km, ix, pr


def check_index(e, ix):
    st = e.index_stats
    for name in ("n_positions", "n_incidences", "n_distinct", "n_singleton", "n_repeated", "nnz"):
        assert st[name] == ix.stats[name], name
    assert np.array_equal(e.get_distinct_kmers(), ix.distinct)
    v, f = e.get_vocab()
    assert np.array_equal(v, ix.vocab)
    assert np.array_equal(f, ix.freq)
    ro, ids = e.get_protein_ids()
    assert np.array_equal(ro, ix.row_offsets)
    assert np.array_equal(ids, ix.ids)


def check_pairs(stats, edges, pr):
    for name in ("n_multi_edges", "n_multi_edges_kept", "n_pairs_kept", "n_edges_out", "sum_count_out"):
        assert stats[name] == pr.stats[name], name
    assert np.array_equal(edges, pr.edges)


# ---------------------------------------------------------------- ARG set vs golden + oracle
@pytest.mark.parametrize("k", [5, 7])
def test_arg_extract_kmers(k, arg_set, arg_oracle, golden):
    g = golden[f"k{k}"]
    with kc.Engine(k) as e:
        e.set_protein_set(arg_set)
        km = e.extract_kmers()
    assert km.size == g["n_positions"]
    assert int(km.astype(np.uint64).sum()) == g["sum_positions"]
    assert int(np.bitwise_xor.reduce(km)) == g["xor_positions"]
    assert km[:6].tolist() == g["first_kmers_protein0"]
    assert np.array_equal(km, arg_oracle[k][1])


@pytest.mark.parametrize("k", [5, 7])
def test_arg_index(k, arg_set, arg_oracle, golden):
    g = golden[f"k{k}"]
    with kc.Engine(k) as e:
        e.set_protein_set(arg_set)
        e.build_index()
        assert sha16(e.get_distinct_kmers().astype("<u4")) == g["sha_distinct"]
        v, f = e.get_vocab()
        assert sha16(v.astype("<u4")) == g["sha_repeated"]
        top = np.argsort(-f.astype(np.int64), kind="stable")[:3]
        assert [[int(v[i]), int(f[i])] for i in top] == g["top3"]
        check_index(e, arg_oracle[k][2])
        # Mphf::hash stand-in: ids of repeated k-mers are their ranks; others are rejected
        probe = np.concatenate([v[:100], v[-100:], np.array([0, 1, 21 ** k - 1, 21 ** k, 2 ** 32 - 1], np.uint32)])
        ids = e.lookup_kmers(probe)
        assert np.array_equal(ids[:100], np.arange(100))
        assert np.array_equal(ids[100:200], np.arange(v.size - 100, v.size))
        exp_tail = [int(np.searchsorted(v, x)) if (x < 21 ** k and x in v) else 0xFFFFFFFF
                    for x in probe[200:].tolist()]
        assert ids[200:].tolist() == exp_tail


@pytest.mark.parametrize("k", [5, 7])
@pytest.mark.parametrize("name,cross,thr", [("cross_gt10", True, 10), ("cross_gt0", True, 0),
                                            ("all_gt10", False, 10)])
def test_arg_pairs(k, name, cross, thr, arg_set, arg_oracle, golden):
    g = golden[f"k{k}"]
    with kc.Engine(k, threshold=thr, cross_class_only=cross, want_blosum=True) as e:
        e.set_protein_set(arg_set)
        e.build_index()
        st = e.score_pairs()
        edges = e.get_edges()
    assert st["n_multi_edges"] == g["n_multi_edges"]
    assert edges.size == g[name]["n"]
    assert sha16(edges_abc(edges)) == g[name]["sha"]
    assert int(edges["count"].sum()) == g[name]["sum_count"] == st["sum_count_out"]
    assert int(edges["blosum"].astype(np.int64).sum()) == g[name]["blosum_sum"]
    assert edges["blosum"][:3].tolist() == g[name]["blosum_first3"]
    if cross:
        assert st["n_multi_edges_kept"] == g["n_multi_edges_cross"]
        assert st["n_pairs_kept"] == g["n_pairs_cross"]
    else:
        assert st["n_multi_edges_kept"] == g["n_multi_edges"]
        assert st["n_pairs_kept"] == g["n_pairs_all"]
    pr = arg_oracle[k][0].score_pairs(thr, cross, True, mode=1)
    check_pairs(st, edges, pr)


def test_arg_reference_api_flow(arg_set, golden, capsys):
    """The reference's call sequence (src/main.rs:216-232) through the module mirror."""
    import io
    from uniprot_kmer_based_clustering_b200.graph import Graph
    from uniprot_kmer_based_clustering_b200.protein import Mphf, ProteinList
    g = golden["k5"]
    log = io.StringIO()
    with kc.Engine(5) as e:
        pl = ProteinList(e, arg_set)
        pl.build_index()
        p0 = pl[0]
        assert p0.get_amr_class() == "beta_lactam"
        assert p0.get_five_mers()[:6].tolist() == g["first_kmers_protein0"]
        assert p0.get_id_and_seq()[0] == golden["protein0_id"]
        phf = Mphf(e)
        v, _ = e.get_vocab()
        assert phf.hash(int(v[1234])) == 1234
        assert np.array_equal(v[p0.get_five_hash()], np.intersect1d(np.unique(p0.get_five_mers()), v))
        graph = Graph.new(pl.kmer_freq, 8, pl, log=log)
        graph.remove_uninteresting_edges(8)
        graph.combine_edges(8)
        edges = graph.align_and_output_pairs(8)
        text = log.getvalue()
        assert "Number of 5mers found in at least two proteins: 231253" in text
        assert "Number of total edges: 258621291" in text
        assert "Number of edges now: 5300233" in text
        assert "Number of edges now: 4350628" in text
        assert text.count("Cross-checking:") == 465 == len(edges)
        ke = graph.edges[0]
        assert ke.get_vertices_key() == [26, 2838]
        kms = ke.get_kmer_values()
        assert kms.size == 167 and np.all(np.diff(kms.astype(np.int64)) > 0)
        a, b = np.unique(pl[26].get_five_mers()), np.unique(pl[2838].get_five_mers())
        assert np.array_equal(kms, np.intersect1d(np.intersect1d(a, b), v))
        assert np.array_equal(ke.get_kmers(), np.searchsorted(v, kms))
        assert ke.get_proteins_ids_and_sequences()[1][0] == arg_set.ids[2838]
        # `pub edges` after combine_edges (src/graph/mod.rs:32): every pair that shares a k-mer
        allp = graph.all_pair_edges()
        assert allp.size == 4350628 and int(allp["count"].sum()) == 5300233
        assert np.array_equal(edges_abc(allp[allp["count"] > 10]), edges_abc(edges))


# ---------------------------------------------------------------- randomised + edge cases
@pytest.mark.parametrize("k", [5, 7])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("cross", [True, False])
def test_random_small_sets(k, seed, cross):
    ps = random_protein_set(seed, 150 + 37 * seed, min_len=0, max_len=90, n_classes=1 + seed, family=5)
    thr = [10, 0, 3, 1][seed]
    km, ix, pr = run_oracle(ps, k, thr, cross)
    with kc.Engine(k, threshold=thr, cross_class_only=cross, want_blosum=True) as e:
        e.set_protein_set(ps)
        assert np.array_equal(e.extract_kmers(), km)
        e.build_index()
        check_index(e, ix)
        st = e.score_pairs()
        check_pairs(st, e.get_edges(), pr)


@pytest.mark.parametrize("k", [5, 7])
def test_degenerate_inputs(k):
    z = np.zeros(0, np.uint8)
    with kc.Engine(k, threshold=0, cross_class_only=False, want_blosum=True) as e:
        # empty set
        e.set_proteins(z, np.zeros(1, np.uint64), np.zeros(0, np.uint32))
        assert e.extract_kmers().size == 0
        st = e.build_index()
        assert st["n_repeated"] == 0 and st["n_positions"] == 0
        assert e.score_pairs()["n_edges_out"] == 0 and e.get_edges().size == 0
        # only proteins shorter than k, including empty ones
        res = np.frombuffer(b"ACDACD" + b"MK", np.uint8)
        e.set_proteins(res, np.array([0, 3, 3, 6, 8], np.uint64), np.array([0, 1, 0, 1], np.uint32))
        assert e.extract_kmers().size == 0
        assert e.build_index()["n_distinct"] == 0
        assert e.score_pairs()["n_edges_out"] == 0
        # one protein: every k-mer is a singleton
        one = np.frombuffer(b"MKHKNQATHKEFSQLEKKFDARLGLYAIDTG", np.uint8)
        e.set_proteins(one, np.array([0, one.size], np.uint64), np.array([0], np.uint32))
        st = e.build_index()
        assert st["n_repeated"] == 0 and st["n_singleton"] == st["n_distinct"] > 0
        assert e.score_pairs()["n_edges_out"] == 0
        # identical proteins: every pair shares every distinct k-mer
        n = 9
        e.set_proteins(np.tile(one, n), (np.arange(n + 1) * one.size).astype(np.uint64), np.zeros(n, np.uint32))
        st = e.build_index()
        nd = np.unique(np.lib.stride_tricks.sliding_window_view(one, k), axis=0).shape[0]
        assert st["n_repeated"] == nd and st["n_singleton"] == 0 and st["nnz"] == n * nd
        ps = e.score_pairs()
        ed = e.get_edges()
        assert ps["n_edges_out"] == n * (n - 1) // 2 and np.all(ed["count"] == nd)
        assert ps["n_multi_edges"] == nd * n * (n - 1) // 2
    # same set, one class, cross-class only: nothing survives
    with kc.Engine(k, threshold=0, cross_class_only=True) as e:
        e.set_proteins(np.tile(one, n), (np.arange(n + 1) * one.size).astype(np.uint64), np.zeros(n, np.uint32))
        e.build_index()
        ps = e.score_pairs()
        assert ps["n_edges_out"] == 0 and ps["n_multi_edges_kept"] == 0 and ps["n_multi_edges"] > 0


@pytest.mark.parametrize("k", [5, 7])
def test_unknown_residues_and_case(k):
    """bytes outside the 20 letters (X, Z, *, lower case, digits) all map to code 20"""
    a = b"MKXKNQZTHK*FSQLEKKFDARLGLYAIDTGmkhknqathkefsqlekk1234567890"
    b_ = b"MKBKNQJTHKOFSQLEKKFDARLGLYAIDTGUUUUUUUUUUUUUUUUUUU-.-.-.-.-."
    ps = kc.ProteinSet(np.frombuffer(a + b_, np.uint8), np.array([0, len(a), len(a) + len(b_)], np.uint64),
                       np.array([0, 1], np.uint32), ["p0|a|b|c0|g", "p1|a|b|c1|g"], ["c0", "c1"])
    km, ix, pr = run_oracle(ps, k, 0, True)
    with kc.Engine(k, threshold=0, want_blosum=True) as e:
        e.set_protein_set(ps)
        assert np.array_equal(e.extract_kmers(), km)
        e.build_index()
        check_index(e, ix)
        v, _ = e.get_vocab()
        assert (21 ** k - 1) in v  # the all-'*' k-mer is an ordinary repeated k-mer
        st = e.score_pairs()
        check_pairs(st, e.get_edges(), pr)
        assert st["n_edges_out"] == 1


@pytest.mark.parametrize("k", [5, 7])
def test_long_proteins_cover_block_and_global_paths(k):
    """positions > 1024 (one CTA, shared memory) and > 32768 (global scratch) per protein"""
    rng = np.random.default_rng(7)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    lens = [40000, 1500, 300, 36000, 5000, 1030, 1024 + k - 1, 1025 + k - 1, 20]
    base = letters[rng.integers(0, 6, size=max(lens))]  # small alphabet: many repeats
    seqs = []
    for i, L in enumerate(lens):
        s = base[:L].copy()
        m = rng.random(L) < 0.02 * i
        s[m] = letters[rng.integers(0, 20, size=int(m.sum()))]
        seqs.append(s)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ps = kc.ProteinSet(np.concatenate(seqs), off, (np.arange(len(lens)) % 2).astype(np.uint32))
    km, ix, pr = run_oracle(ps, k, 10, False)
    with kc.Engine(k, threshold=10, cross_class_only=False, want_blosum=True) as e:
        e.set_protein_set(ps)
        assert np.array_equal(e.extract_kmers(), km)
        e.build_index()
        check_index(e, ix)
        check_pairs(e.score_pairs(), e.get_edges(), pr)


@pytest.mark.parametrize("k,cross", [(5, False), (7, False), (5, True)])
def test_synthetic_20k_against_oracle(k, cross):
    ps = kc.ProteinSet.synthetic(20000, "A", 0xB2000003, threads=8)
    km, ix, pr = run_oracle(ps, k, 10, cross)
    with kc.Engine(k, threshold=10, cross_class_only=cross, want_blosum=True) as e:
        e.set_protein_set(ps)
        e.build_index()
        check_index(e, ix)
        st = e.score_pairs()
        check_pairs(st, e.get_edges(), pr)
        assert st["n_edges_out"] > 1000


def test_skewed_lengths_against_oracle():
    ps = kc.ProteinSet.synthetic(6000, "B", 0xB2000005, threads=8)
    km, ix, pr = run_oracle(ps, 5, 10, False)
    with kc.Engine(5, threshold=10, cross_class_only=False, want_blosum=False) as e:
        e.set_protein_set(ps)
        assert np.array_equal(e.extract_kmers(), km)
        e.build_index()
        check_index(e, ix)
        st = e.score_pairs()
        ed = e.get_edges()
        assert np.array_equal(edges_abc(ed), edges_abc(pr.edges))
        assert np.all(ed["blosum"] == 0)


@pytest.mark.parametrize("slices", ["1", "5", "64"])
def test_l2_blocking_slices_do_not_change_results(slices, monkeypatch, arg_set, arg_oracle):
    """the index is built slice by slice over the k-mer universe (L2 blocking); any slicing must
    give the same index and edges"""
    from uniprot_kmer_based_clustering_b200 import engine as eng
    monkeypatch.setitem(eng.DEFAULTS, "index_slices", int(slices))
    monkeypatch.setitem(eng.DEFAULTS, "index_build", "table")
    ps = random_protein_set(4, 300, min_len=0, max_len=400, n_classes=3, family=6)
    for k in (5, 7):
        km, ix, pr = run_oracle(ps, k, 2, True)
        with kc.Engine(k, threshold=2, cross_class_only=True, want_blosum=True) as e:
            e.set_protein_set(ps)
            e.build_index()
            check_index(e, ix)
            check_pairs(e.score_pairs(), e.get_edges(), pr)
    with kc.Engine(5, threshold=10, want_blosum=True) as e:
        e.set_protein_set(arg_set)
        e.build_index()
        check_index(e, arg_oracle[5][2])
        st = e.score_pairs()
        check_pairs(st, e.get_edges(), arg_oracle[5][0].score_pairs(10, True, True, mode=1))


# ---------------------------------------------------------------- shards, device input, extras
@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_shards_partition_the_pair_triangle(n_shards, arg_set):
    with kc.Engine(7, threshold=10, cross_class_only=False, want_blosum=True) as e:
        e.set_protein_set(arg_set)
        e.build_index()
        full_st = e.score_pairs()
        full = e.get_edges()
        parts, tot = [], {"n_multi_edges_kept": 0, "n_pairs_kept": 0, "n_edges_out": 0, "n_rows": 0}
        for s in range(n_shards):
            st = e.score_pairs(s, n_shards)
            parts.append(e.get_edges())
            for key in tot:
                tot[key] += st[key]
            assert st["n_multi_edges"] == full_st["n_multi_edges"]
        for key in ("n_multi_edges_kept", "n_pairs_kept", "n_edges_out"):
            assert tot[key] == full_st[key], key
        assert tot["n_rows"] == arg_set.n
        merged = np.concatenate(parts)
        merged = merged[np.lexsort((merged["b"], merged["a"]))]
        assert np.array_equal(merged, full)
        # work balance: the shards are cut by estimated work (multi-edges + list lengths), so the multi-edges a
        # shard accumulated (its share of the kept multi-edges) stay within 2x of an equal share
        work = []
        for sh in range(n_shards):
            work.append(e.score_pairs(sh, n_shards)["n_multi_edges_kept"])
        assert max(work) <= 2 * max(1, sum(work)) // n_shards + 1, work


def test_device_resident_input_matches_host_input(arg_set):
    torch = pytest.importorskip("torch")
    res = torch.from_numpy(arg_set.residues).cuda()
    off = torch.from_numpy(arg_set.offsets.astype(np.int64)).cuda()
    cls = torch.from_numpy(arg_set.class_id.astype(np.int32)).cuda()
    with kc.Engine(5, want_blosum=True) as e, kc.Engine(5, want_blosum=True) as h:
        e.set_proteins_ptr(res.data_ptr(), off.data_ptr(), cls.data_ptr(), arg_set.n, on_device=True)
        h.set_protein_set(arg_set)
        assert e.build_index() == h.build_index()
        assert e.score_pairs() == h.score_pairs()
        assert np.array_equal(e.get_edges(), h.get_edges())


@pytest.mark.parametrize("k", [5, 7])
def test_chunked_upload_builds_the_same_index(k, arg_set, arg_oracle):
    """kc_set_proteins from host buffers uploads a large residue stream in chunks on a copy stream and the
    streaming build runs one level-1 pass per chunk (level 2 reads every partition through a segment table).
    no_upload_overlap = 2 forces the chunked path at this size; the result must not depend on it, and a rebuild
    of the resident stream (one pass) must give the same index again."""
    o, _, ix = arg_oracle[k]
    pr = o.score_pairs(10, False, True, mode=1)
    with kc.Engine(k, cross_class_only=False, want_blosum=True, no_upload_overlap=2) as e, \
            kc.Engine(k, cross_class_only=False, want_blosum=True, no_upload_overlap=1) as h:
        for _ in range(2):  # the second round uploads into buffers the first one still owns
            e.set_protein_set(arg_set)
            h.set_protein_set(arg_set)
            ist = e.build_index()
            assert ist == h.build_index()
            check_index(e, ix)
            pst = e.score_pairs()
            assert pst == h.score_pairs()
            check_pairs(pst, e.get_edges(), pr)
            assert np.array_equal(e.get_edges(), h.get_edges())
        assert e.build_index() == ist  # resident now: a single pass
        assert e.score_pairs() == pst
        check_pairs(pst, e.get_edges(), pr)


def test_rerun_is_idempotent_and_edge_buffer_grows(arg_set):
    with kc.Engine(5, threshold=10, cross_class_only=False, max_edges=1000) as e:
        e.set_protein_set(arg_set)
        e.build_index()
        st1 = e.score_pairs()
        e1 = e.get_edges()
        assert st1["n_retries"] == 1 and st1["n_edges_out"] == 5969297
        st2 = e.score_pairs()
        assert st2["n_retries"] == 0
        assert np.array_equal(e1, e.get_edges())
        assert np.all((e1["a"][1:] > e1["a"][:-1]) | ((e1["a"][1:] == e1["a"][:-1]) & (e1["b"][1:] > e1["b"][:-1])))
        e.build_index()
        assert e.score_pairs()["n_edges_out"] == st1["n_edges_out"]


def test_permuting_the_input_permutes_the_edges():
    ps = kc.ProteinSet.synthetic(3000, "A", 99, threads=4)
    perm = np.random.default_rng(5).permutation(ps.n)
    lens = np.diff(ps.offsets.astype(np.int64))
    res2 = np.concatenate([ps.residues[int(ps.offsets[p]):int(ps.offsets[p + 1])] for p in perm])
    off2 = np.concatenate([[0], np.cumsum(lens[perm])]).astype(np.uint64)
    with kc.Engine(5, cross_class_only=True, want_blosum=True) as e:
        e.set_protein_set(ps)
        e.build_index()
        s1 = e.score_pairs()
        e1 = e.get_edges()
        e.set_proteins(res2, off2, ps.class_id[perm])
        e.build_index()
        s2 = e.score_pairs()
        e2 = e.get_edges()
    for key in ("n_multi_edges", "n_multi_edges_kept", "n_pairs_kept", "n_edges_out", "sum_count_out"):
        assert s1[key] == s2[key]
    a, b = perm[e2["a"]], perm[e2["b"]]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    order = np.lexsort((hi, lo))
    assert np.array_equal(lo[order], e1["a"]) and np.array_equal(hi[order], e1["b"])
    assert np.array_equal(e2["count"][order], e1["count"])
    assert np.array_equal(e2["blosum"][order], e1["blosum"])


def test_bitset_pair_counts_match_sparse_counts(arg_set, arg_oracle):
    rows = np.array([26, 39, 67, 2838, 0, 3, 61, 75, 10618, 5000, 5001], dtype=np.uint32)
    ix = arg_oracle[5][2]
    sets = [set(ix.ids[int(ix.row_offsets[r]):int(ix.row_offsets[r + 1])].tolist()) for r in rows]
    exp = np.array([[len(a & b) for b in sets] for a in sets], dtype=np.uint32)
    with kc.Engine(5) as e:
        e.set_protein_set(arg_set)
        e.build_index()
        got = e.bitset_pair_counts(rows)
    assert np.array_equal(got, exp)
    assert got[0, 3] == 167 and got[1, 3] == 217 and got[2, 3] == 258


def test_call_order_errors():
    with kc.Engine(5) as e:
        with pytest.raises(kc.KcError):
            e.build_index()
        with pytest.raises(kc.KcError):
            e.score_pairs()
        with pytest.raises(kc.KcError):
            e.set_proteins(np.zeros(4, np.uint8), np.array([0, 3, 2], np.uint64), np.zeros(2, np.uint32))


def test_cli_prints_the_reference_counters(tmp_path, arg_fasta_bytes, golden):
    """`kmer_cluster <fasta> <threads>` = `cargo run --release -- <fasta> <threads>` (src/main.rs:50-60)"""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "uniprot_kmer_based_clustering_b200", "bin", "kmer_cluster")
    fa = tmp_path / "arg.fasta"
    fa.write_bytes(arg_fasta_bytes)
    r = subprocess.run([exe, str(fa), "4", "--blosum"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    err = r.stderr
    assert err.startswith("We start main\n")
    for line in ("Number of 5mers found in at least two proteins: 231253", "Number of total edges: 258621291",
                 "Remove edges without diverging AMR labels", "Number of edges now: 5300233",
                 "Combine edges with the same two vertices", "Number of edges now: 4350628"):
        assert line in err, line
    assert err.count("Cross-checking:") == 465
    assert "kmers in common:567" in err
    rows = [ln.split("\t") for ln in r.stdout.strip().splitlines()[1:]]
    assert len(rows) == 465 and rows[0][:2] == ["26", "2838"] and rows[0][4] == "167"
    assert sum(int(x[5]) for x in rows) == golden["k5"]["cross_gt10"]["blosum_sum"]
    # wrong usage panics like the reference (exit status 101)
    assert subprocess.run([exe, str(fa)], capture_output=True).returncode == 101
    assert subprocess.run([exe, "/nonexistent.fasta", "2"], capture_output=True).returncode == 101
    r7 = subprocess.run([exe, str(fa), "2", "--k", "7"], capture_output=True, text=True, timeout=120)
    assert "Number of 7mers found in at least two proteins: 288551" in r7.stderr
    assert r7.stderr.count("Cross-checking:") == 463


def test_device_edge_view_matches_host_copy(arg_set):
    """kc_get_edges_device: the sorted edge list in HBM (what the NCCL gather sends)"""
    torch = pytest.importorskip("torch")
    from uniprot_kmer_based_clustering_b200.sharded import _DeviceWords, gather_edges_device
    with kc.Engine(7, threshold=10, cross_class_only=False, want_blosum=True) as e:
        e.set_protein_set(arg_set)
        e.build_index()
        e.score_pairs()
        host = e.get_edges()
        ptr, n = e.edges_device()
        assert n == host.size and ptr != 0
        dev = torch.as_tensor(_DeviceWords(ptr, n * 4), device="cuda")
        assert np.array_equal(dev.cpu().numpy().view(kc.EDGE_DTYPE), host)
        assert np.array_equal(gather_edges_device(e, None, 0, 1), host)


@pytest.mark.parametrize("k", [5, 7])
def test_subsampling_mode_matches_oracle(k):
    """Protein::new_with_rand_fivemers (src/protein.rs:77-104): a tenth of the start positions; the
    engine and the oracle share the counter-based sampler definition (kc_sample_position)"""
    ps = random_protein_set(9, 400, min_len=0, max_len=500, n_classes=3, family=8, mutate=0.03)
    lens = [40000, 1500, 36000, 1030]  # block / global-scratch extract paths under sampling
    rng = np.random.default_rng(3)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    base = letters[rng.integers(0, 8, size=max(lens))]
    extra = np.concatenate([base[:L] for L in lens])
    res = np.concatenate([ps.residues, extra])
    off = np.concatenate([ps.offsets, ps.offsets[-1] + np.cumsum(lens).astype(np.uint64)])
    cls = np.concatenate([ps.class_id, np.array([0, 1, 2, 0], np.uint32)])
    seed = 0xB2005EED
    o = Oracle(k, 4, sample_every=10, sample_seed=seed)
    o.set_proteins(res, off, cls)
    km = o.extract_kmers()
    ix = o.build_index()
    pr = o.score_pairs(3, False, True, mode=1)
    with kc.Engine(k, threshold=3, cross_class_only=False, want_blosum=True, sample_every=10, sample_seed=seed) as e:
        e.set_proteins(res, off, cls)
        got = e.extract_kmers()
        assert got.size == km.size == sum(max(0, int(L) - k + 1) // 10 for L in np.diff(off.astype(np.int64)))
        assert np.array_equal(got, km)
        e.build_index()
        check_index(e, ix)
        check_pairs(e.score_pairs(), e.get_edges(), pr)
    # sampled k-mers are a subset of the full k-mers
    with kc.Engine(k) as e:
        e.set_proteins(res, off, cls)
        full = e.extract_kmers()
    assert np.isin(km, full).all()


def test_tree_from_engine_index(arg_set, arg_oracle):
    from uniprot_kmer_based_clustering_b200.tree import Tree
    n = 600
    sub = kc.ProteinSet(arg_set.residues[:int(arg_set.offsets[n])], arg_set.offsets[:n + 1], arg_set.class_id[:n])
    o = Oracle(7, 4)
    o.set_proteins(sub.residues, sub.offsets, sub.class_id)
    o.extract_kmers()
    ix = o.build_index()
    exp = Tree.from_id_rows(ix.row_offsets, ix.ids, ix.stats["n_repeated"])
    with kc.Engine(7) as e:
        e.set_protein_set(sub)
        e.build_index()
        got = Tree.from_engine(e)
    assert np.array_equal(got.serialize(), exp.serialize())
    assert np.array_equal(got.clusters(), exp.clusters())


def test_cli_tree_and_sampling(tmp_path, arg_fasta_bytes):
    import json
    import os
    import subprocess
    from conftest import GOLDEN_DIR, ROOT
    exe = os.path.join(ROOT, "uniprot_kmer_based_clustering_b200", "bin", "kmer_cluster")
    fa = tmp_path / "arg.fasta"
    fa.write_bytes(arg_fasta_bytes)
    g = json.load(open(os.path.join(GOLDEN_DIR, "tree_golden.json")))["k7"]
    r = subprocess.run([exe, str(fa), "4", "--k", "7", "--tree"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-400:]
    assert f"Tree: {g['n_clusters']} top-level clusters, {g['n_merges']} merges" in r.stderr
    rows = [ln for ln in r.stdout.splitlines() if ln.startswith("#") and not ln.startswith("#protein")]
    assert len(rows) == 10619
    r2 = subprocess.run([exe, str(fa), "2", "--sample-every", "10", "--seed", "7"], capture_output=True, text=True,
                        timeout=120)
    assert r2.returncode == 0 and "Number of 5mers found in at least two proteins:" in r2.stderr


@pytest.mark.parametrize("blosum", [True, False])
def test_stream_pair_kernel_over_materialised_lists(blosum, monkeypatch, arg_set, arg_oracle):
    """kc_config.pair_lists materialises the multi-edge lists (table build) and scores them with the stream
    kernel (opt-in: the fill costs more than the stream kernel saves); both paths must give identical results"""
    from uniprot_kmer_based_clustering_b200 import engine as eng
    monkeypatch.setitem(eng.DEFAULTS, "pair_lists", True)
    for k, cross, thr in ((5, True, 10), (7, False, 10)):
        pr = arg_oracle[k][0].score_pairs(thr, cross, blosum, mode=1)
        with kc.Engine(k, threshold=thr, cross_class_only=cross, want_blosum=blosum) as e:
            e.set_protein_set(arg_set)
            e.build_index()
            check_pairs(e.score_pairs(), e.get_edges(), pr)
    ps = kc.ProteinSet.synthetic(20000, "A", 0xB2000004, threads=8)
    km, ix, pr = run_oracle(ps, 7, 10, False, blosum=blosum)
    with kc.Engine(7, threshold=10, cross_class_only=False, want_blosum=blosum) as e:
        e.set_protein_set(ps)
        e.build_index()
        check_pairs(e.score_pairs(), e.get_edges(), pr)


@pytest.mark.parametrize("k,cross", [(5, True), (7, False), (7, True)])
def test_pair_index_is_a_perfect_hash_of_the_canonical_one(k, cross, arg_set, arg_oracle):
    """kc_get_pair_index (the ids the pair stage really uses) against the oracle's canonical index:
    same vocabulary, same kmer_freq, same BLOSUM self-scores, same per-protein k-mer sets"""
    ix = arg_oracle[k][2]
    diag = np.array([9, 4, 5, 4, 6, 7, 6, 5, 5, 6, 8, 5, 5, 5, 4, 4, 4, 11, 7, 6, 0])
    with kc.Engine(k, cross_class_only=cross, want_blosum=True) as e:
        e.set_protein_set(arg_set)
        e.build_index()
        v, f, ss, ro, ids = e.get_pair_index()
    assert np.unique(v).size == v.size == ix.vocab.size
    order = np.argsort(v, kind="stable")
    assert np.array_equal(v[order], ix.vocab)
    assert np.array_equal(f[order], ix.freq)
    digits = (v[:, None].astype(np.int64) // (21 ** np.arange(k))[None, :]) % 21
    assert np.array_equal(ss, diag[digits].sum(axis=1))
    assert np.array_equal(ro, ix.row_offsets)
    canon = np.empty(v.size, dtype=np.int64)
    canon[order] = np.arange(v.size)
    got = canon[ids]
    for p in (0, 1, 26, 2838, arg_set.n - 1):
        a, b = int(ro[p]), int(ro[p + 1])
        assert np.array_equal(np.sort(got[a:b]), ix.ids[a:b])
    # every row at once: sort within rows via a (row, id) key
    row_of = np.repeat(np.arange(arg_set.n), np.diff(ro).astype(np.int64))
    key = row_of * np.int64(v.size + 1) + got
    assert np.array_equal(np.sort(key) - row_of * np.int64(v.size + 1), ix.ids)


@pytest.mark.parametrize("n_hot,expect", [(2500, 8192), (9000, 0)])
def test_kmers_with_more_holders_than_a_shared_memory_bucket(n_hot, expect, index_flavour):
    """one k-mer held by more proteins than a shared-memory bucket takes.  The streaming build sends such a
    bucket through its global-memory path (no fallback, no retry: 9 000 holders, src/graph/mod.rs:44-48 sizes
    f(f-1)/2 edges for them).  Round 1's partitioned build: 2 500 holders overflow the 4 096-record buckets and
    fit the 8 192-record ones, 9 000 holders fit neither and it builds the universe-table index instead"""
    rng = np.random.default_rng(11)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    hot = letters[:7]
    tail = 0 if n_hot == 9000 else 120
    seqs = [np.concatenate([hot, letters[rng.integers(0, 20, size=tail)]]) for _ in range(n_hot)]
    n = len(seqs)
    res = np.concatenate(seqs).astype(np.uint8)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([x.size for x in seqs])
    cls = (np.arange(n) % 3).astype(np.uint32)
    ps = kc.ProteinSet(res, off, cls, [], ["a", "b", "c"])
    km, ix, pr = run_oracle(ps, 7, 3, True)
    with kc.Engine(7, threshold=3, cross_class_only=True, want_blosum=True) as e:
        e.set_protein_set(ps)
        e.build_index()
        if index_flavour == "bucket":
            assert e.index_flavour() == expect
            e.build_index()  # the slot size that worked is remembered for this protein set
            assert e.index_flavour() == expect
        elif index_flavour.startswith("stream"):
            assert e.index_flavour() == 1
        else:
            assert e.index_flavour() == 0
        check_index(e, ix)
        check_pairs(e.score_pairs(), e.get_edges(), pr)


@pytest.mark.parametrize("n_shards", [2, 3, 5])
@pytest.mark.parametrize("k,cross", [(7, False), (7, True), (5, False)])
def test_sharded_index_build_adds_up_to_the_whole(n_shards, k, cross, index_flavour):
    """kc_build_index_shard + kc_score_pairs_shard for every shard in turn (one GPU standing in for
    n_shards ranks): the index stats and pair counters add up to the oracle's whole-set numbers and
    the union of the edge lists is the oracle's edge list.  With the universe-table build every
    shard holds the whole index (stats are whole-set numbers on every rank)."""
    ps = kc.ProteinSet.synthetic(6000, "A", 0xB2000004, threads=8)
    km, ix, pr = run_oracle(ps, k, 10, cross)
    tot_i = {name: 0 for name in ("n_positions", "n_incidences", "n_distinct", "n_singleton", "n_repeated", "nnz")}
    tot_p = {name: 0 for name in ("n_multi_edges", "n_multi_edges_kept", "n_pairs_kept", "n_edges_out",
                                  "sum_count_out", "n_rows")}
    parts = []
    sharded = None
    with kc.Engine(k, threshold=10, cross_class_only=cross, want_blosum=True) as e:
        e.set_protein_set(ps)
        for s in range(n_shards):
            ist = e.build_index(s, n_shards)
            info = e.index_shard_info()
            sharded = info["n_shards"] > 1
            assert sharded == (index_flavour != "table")  # (the streaming build shards through round 1's build)
            pst = e.score_pairs(s, n_shards)
            parts.append(e.get_edges())
            for name in tot_i:
                tot_i[name] += ist[name]
            for name in tot_p:
                tot_p[name] += pst[name]
            if sharded:
                assert info["shard"] == s and info["n_blocks"] == 2 * n_shards
                assert info["block_bounds"][0] == 0 and info["block_bounds"][-1] == ps.n
                if not cross:  # pair order = input order: the host mirror of the block cut applies as is
                    from uniprot_kmer_based_clustering_b200 import sharded as sh
                    lens = np.diff(ps.offsets.astype(np.int64))
                    exp_bounds, _ = sh.zigzag_blocks(np.maximum(lens - k + 1, 0), n_shards)
                    assert np.array_equal(info["block_bounds"].astype(np.int64), exp_bounds)
                with pytest.raises(kc.KcError):
                    e.get_vocab()
                with pytest.raises(kc.KcError):
                    e.score_pairs((s + 1) % n_shards, n_shards)
    scale = 1 if sharded else n_shards
    for name in tot_i:
        assert tot_i[name] == ix.stats[name] * scale, name
    assert tot_p["n_multi_edges"] == pr.stats["n_multi_edges"] * scale
    assert tot_p["n_rows"] == ps.n
    for name in ("n_multi_edges_kept", "n_pairs_kept", "n_edges_out", "sum_count_out"):
        assert tot_p[name] == pr.stats[name], name
    from uniprot_kmer_based_clustering_b200.sharded import merge_edge_lists
    assert np.array_equal(merge_edge_lists(parts), pr.edges)


# ---------------------------------------------------------------- full-size benchmark configurations
@pytest.mark.parametrize("name", ["synth_20k_k5", "synth_100k_k5", "synth_250k_skew_k7", "synth_1m_k7"])
def test_full_size_synthetic_goldens(name, index_flavour):
    """BASELINE.json's synthetic configurations at FULL size against the oracle's committed golden values
    (tests/golden/synth_golden.json, written by tests/golden/make_synth_golden.py): the index counters, the
    counters the reference prints (src/graph/mod.rs:50-51,545,695) and the SHA-256 of the whole sorted
    (a, b, count, blosum) edge list.  bench.py asserts the same SHA at every GPU count."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN_DIR
    if index_flavour in ("stream512", "bucket") or (index_flavour == "table" and name in ("synth_1m_k7", "synth_250k_skew_k7")):
        pytest.skip("full size: the default build (and the table build on the k = 5 sets)")
    with open(os.path.join(GOLDEN_DIR, "synth_golden.json")) as fh:
        g = json.load(fh)[name]
    ps = kc.ProteinSet.synthetic(g["n"], g["law"], int(g["seed"], 16), threads=os.cpu_count() or 8)
    assert hashlib.sha256(ps.residues.tobytes()).hexdigest() == g["residues_sha256"]
    with kc.Engine(g["k"], threshold=g["threshold"], cross_class_only=g["cross_class_only"], want_blosum=True) as e:
        e.set_protein_set(ps)
        ist = e.build_index()
        pst = e.score_pairs()
        edges = e.get_edges()
    for key, val in g["index"].items():
        assert ist[key] == val, key
    for key, val in g["pairs"].items():
        assert pst[key] == val, key
    assert hashlib.sha256(np.ascontiguousarray(edges).tobytes()).hexdigest() == g["edges_sha256"]
    assert int(edges["blosum"].astype(np.int64).sum()) == g["sum_blosum"]


def test_cli_two_gpus_matches_one(tmp_path, arg_fasta_bytes, golden, index_flavour):
    """`kmer_cluster <fasta> <threads> --gpus 2`: one engine and one NCCL rank per GPU inside the C library (host
    threads, kc_comm_* / kc_*_dist / kc_gather_edges_shared); same counters and the same TSV as one GPU.  The ARG
    set's hot k-mers overflow the sharded bucket build, so this also covers its streaming fallback."""
    import os
    import subprocess
    if index_flavour != "stream":
        pytest.skip("one flavour is enough: the CLI picks the build itself")
    if kc.lib().kc_device_count() < 2:
        pytest.skip("needs two GPUs")
    from uniprot_kmer_based_clustering_b200._lib import _PKG
    exe = os.path.join(_PKG, "bin", "kmer_cluster")
    fa = tmp_path / "arg.fasta"
    fa.write_bytes(arg_fasta_bytes)
    outs = []
    for gpus in ("1", "2"):
        r = subprocess.run([exe, str(fa), "4", "--gpus", gpus], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r)
    g = golden["k5"]
    for r in outs:
        assert "Number of 5mers found in at least two proteins: 231253" in r.stderr
        assert "Number of total edges: 258621291" in r.stderr
        assert "Number of edges now: 5300233" in r.stderr
        assert "Number of edges now: 4350628" in r.stderr
        assert r.stderr.count("Cross-checking:") == 465
    tsv = ["\n".join(ln for ln in r.stdout.splitlines() if not ln.startswith("NCCL version")) for r in outs]
    assert tsv[0] == tsv[1]  # (NCCL announces its version on stdout)
