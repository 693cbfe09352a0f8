"""CPU suite: the oracle against the golden vectors, the host logic, the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import uniprot_kmer_based_clustering_b200 as kc
from conftest import ROOT, edges_abc, has_gpu, random_protein_set, sha16
from oracle.fasta_ref import parse_fasta_bytes
from oracle.oracle import Oracle, self_score
from oracle.ref_model import run_reference_model


# ---------------------------------------------------------------- oracle vs golden (§8c)
@pytest.mark.parametrize("k", [5, 7])
def test_oracle_extract_and_index_match_golden(k, golden, arg_oracle, arg_set):
    g = golden[f"k{k}"]
    _, km, ix = arg_oracle[k]
    assert golden["n_proteins"] == arg_set.n == 10619
    assert golden["protein0_id"] == arg_set.ids[0]
    assert km.size == g["n_positions"]
    assert int(km.astype(np.uint64).sum()) == g["sum_positions"]
    assert int(np.bitwise_xor.reduce(km)) == g["xor_positions"]
    assert km[:6].tolist() == g["first_kmers_protein0"]
    assert int(km[int(arg_set.offsets[1]) - k]) == g["last_kmer_protein0"]
    for name in ("n_positions", "n_incidences", "n_distinct", "n_singleton", "n_repeated", "nnz"):
        assert ix.stats[name] == g[name], name
    assert sha16(ix.distinct.astype("<u4")) == g["sha_distinct"]
    assert sha16(ix.vocab.astype("<u4")) == g["sha_repeated"]
    assert (int(ix.vocab.min()), int(ix.vocab.max())) == (g["repeated_min"], g["repeated_max"])
    top = np.argsort(-ix.freq.astype(np.int64), kind="stable")[:3]
    assert [[int(ix.vocab[i]), int(ix.freq[i])] for i in top] == g["top3"]


@pytest.mark.parametrize("k", [5, 7])
@pytest.mark.parametrize("name,cross,thr", [("cross_gt10", True, 10), ("cross_gt0", True, 0),
                                            ("all_gt10", False, 10)])
def test_oracle_pairs_match_golden(k, name, cross, thr, golden, arg_oracle):
    g = golden[f"k{k}"]
    o, _, _ = arg_oracle[k]
    r = o.score_pairs(thr, cross, True, mode=1)
    e = r.edges
    assert r.stats["n_multi_edges"] == g["n_multi_edges"]
    assert e.size == g[name]["n"]
    assert sha16(edges_abc(e)) == g[name]["sha"]
    assert int(e["count"].sum()) == g[name]["sum_count"]
    assert int(e["count"].max()) == g[name]["max_count"]
    assert int(e["blosum"].astype(np.int64).sum()) == g[name]["blosum_sum"]
    assert e["blosum"][:3].tolist() == g[name]["blosum_first3"]
    if cross:
        assert r.stats["n_multi_edges_kept"] == g["n_multi_edges_cross"]
        assert r.stats["n_pairs_kept"] == g["n_pairs_cross"]
    else:
        assert r.stats["n_pairs_kept"] == g["n_pairs_all"]


@pytest.mark.parametrize("k", [5, 7])
def test_oracle_literal_mode_equals_fast_mode(k, arg_oracle):
    o, _, _ = arg_oracle[k]
    a = o.score_pairs(10, True, True, mode=0)
    b = o.score_pairs(10, True, True, mode=1)
    assert a.stats == b.stats
    assert np.array_equal(a.edges, b.edges)


@pytest.mark.parametrize("k,seed", [(5, 1), (5, 2), (7, 3)])
def test_oracle_matches_literal_reference_model(k, seed):
    """oracle/ref_model.py follows the reference's data structures (triangular edge layout,
    class filter, combine) with random MPHF ids and a shuffled arrival order."""
    ps = random_protein_set(seed, 40, min_len=12, max_len=50, n_classes=3, family=4, mutate=0.04,
                            alphabet="ACDEFGHIKLMNPQRSTVWYX")
    recs = [(ps.ids[i], ps.seq(i)) for i in range(ps.n)]
    for thr in (10, 1):
        m = run_reference_model(recs, k=k, threshold=thr, seed=seed, shuffle_arrival=True)
        o = Oracle(k, 2)
        o.set_proteins(ps.residues, ps.offsets, ps.class_id)
        o.extract_kmers()
        ix = o.build_index()
        r = o.score_pairs(thr, True, True, mode=0)
        assert m["repeated"] == ix.vocab.tolist()
        assert [m["freq_by_kmer"][int(v)] for v in ix.vocab] == ix.freq.tolist()
        for p in range(ps.n):
            row = ix.ids[int(ix.row_offsets[p]):int(ix.row_offsets[p + 1])]
            assert m["hash_sets"][p] == ix.vocab[row].tolist()
        assert (m["n_total_edges"], m["n_after_class"], m["n_after_combine"]) == (
            r.stats["n_multi_edges"], r.stats["n_multi_edges_kept"], r.stats["n_pairs_kept"])
        assert [(a, b, c) for a, b, c, _ in m["pairs"]] == [
            (int(e["a"]), int(e["b"]), int(e["count"])) for e in r.edges]


def test_blosum_diagonal_matches_reference_table():
    """src/blosum.rs lower triangle, index j(j+1)/2+i: the diagonal drives the self-scores."""
    diag = [9, 4, 5, 4, 6, 7, 6, 5, 5, 6, 8, 5, 5, 5, 4, 4, 4, 11, 7, 6]
    for c in range(20):
        assert self_score(c, 1 + 4) == diag[c] + 4 * diag[0]  # kmer = 0,0,0,0,c -> C C C C x
    assert self_score(20, 5) == 4 * 9  # '*' scores 0
    ref = "/root/reference/src/blosum.rs"
    if os.path.exists(ref):  # only in the build container
        rows = [ln for ln in open(ref).read().splitlines() if re.match(r"\s*/\*[A-Z]\*/", ln)]
        tri = [int(x) for ln in rows for x in re.findall(r"-?\d+", ln.split("*/")[1])]
        assert len(tri) == 210
        assert [tri[j * (j + 1) // 2 + j] for j in range(20)] == diag


# ---------------------------------------------------------------- host logic
def test_product_fasta_parser_matches_independent_reader(arg_fasta_bytes, arg_set):
    ref = parse_fasta_bytes(arg_fasta_bytes)
    assert ref["ids"] == arg_set.ids
    assert np.array_equal(ref["residues"], arg_set.residues)
    assert np.array_equal(ref["offsets"], arg_set.offsets)
    assert np.array_equal(ref["class_id"], arg_set.class_id)
    assert ref["class_names"] == arg_set.class_names
    assert len(arg_set.class_names) == 15 and arg_set.n_missing_class == 0


def test_fasta_parser_edge_cases():
    data = (b">a|b|c|cls1|g desc text\nACDE\nFGH\r\n>short|x\nAC\n>empty|1|2|cls1|\n"
            b">trail|1|2|cls2|gene \nMKV\n\n>last|1|2|cls1|z\nWW")
    for threads in (1, 3):
        ps = kc.ProteinSet.from_fasta_bytes(data, threads)
        ref = parse_fasta_bytes(data)
        assert ps.ids == ["a|b|c|cls1|g", "short|x", "empty|1|2|cls1|", "trail|1|2|cls2|gene", "last|1|2|cls1|z"]
        assert ps.ids == ref["ids"]
        assert [ps.seq(i) for i in range(ps.n)] == ["ACDEFGH", "AC", "", "MKV", "WW"]
        assert np.array_equal(ps.offsets, ref["offsets"])
        assert ps.class_names == ["cls1", "", "cls2"] == ref["class_names"]
        assert ps.class_id.tolist() == [0, 1, 0, 2, 0] == ref["class_id"].tolist()
        assert ps.n_missing_class == 1
    assert kc.ProteinSet.from_fasta_bytes(b"", 2).n == 0
    with pytest.raises(kc.KcError):
        kc.ProteinSet.from_fasta("/nonexistent/file.fasta")


def _py_splitmix(state):
    state = (state + 0x9E3779B97F4A7C15) & (2**64 - 1)
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    return state, z ^ (z >> 31)


def _py_stream(seed, index, tag):
    s = seed ^ (((index + 1) * 0x9E3779B97F4A7C15) & (2**64 - 1)) ^ ((tag * 0xC2B2AE3D27D4EB4F) & (2**64 - 1))
    s, _ = _py_splitmix(s)
    return s


def _py_generate(n, law, seed):
    """Pure-Python restatement of generator G1 (include/kc_synth.h)."""
    letters = "LAGVISTFREKDPQNYMHWC"
    counts = [377380, 336508, 269059, 258845, 257326, 206108, 194902, 176903, 163437, 158642,
              158511, 153089, 143715, 124590, 118853, 101563, 92047, 64642, 54450, 26176]
    cum = np.cumsum(counts)
    total = int(cum[-1])
    assert total == 3436746

    def pick(u):
        return letters[int(np.searchsorted(cum, u, side="right"))]

    seqs, cls = [], []
    for i in range(n):
        fam, j = divmod(i, 16)
        s = _py_stream(seed, fam, 1)
        if law == 0:
            L = 50
            for _ in range(4):
                s, r = _py_splitmix(s)
                L += ((r >> 32) * 151) >> 32
        else:
            s, r = _py_splitmix(s)
            t = ((r >> 32) * 65536) >> 32
            L = 50 + ((((t ** 4) >> 32) * 1950) >> 32)
        s = _py_stream(seed, fam, 2)
        base = []
        for _ in range(L):
            s, r = _py_splitmix(s)
            base.append("X" if (r & 8191) == 0 else pick(((r >> 32) * total) >> 32))
        s = _py_stream(seed, i, 3)
        out = []
        for p in range(L):
            s, r = _py_splitmix(s)
            out.append(pick((((r >> 16) & 0xFFFFFFFF) * total) >> 32) if (r & 0xFFFF) < j * 1311 else base[p])
        seqs.append("".join(out))
        cls.append((fam + j) % 15 if fam % 8 == 7 else fam % 15)
    return seqs, cls


@pytest.mark.parametrize("law,seed", [("A", 0xB2000003), ("B", 0xB2000005)])
def test_synthetic_generator_matches_python_restatement(law, seed):
    n = 16 * 8 + 5
    ps = kc.ProteinSet.synthetic(n, law, seed, threads=3, with_ids=True)
    seqs, cls = _py_generate(n, 0 if law == "A" else 1, seed)
    assert [ps.seq(i) for i in range(n)] == seqs
    assert ps.class_id.tolist() == cls
    lens = np.diff(ps.offsets.astype(np.int64))
    assert lens.min() >= 50 and lens.max() <= (650 if law == "A" else 2000)
    assert ps.ids[20] == f"S20|FEATURES|SYNTH|class{cls[20]}|fam1"
    # order independence: a prefix generated alone is identical
    ps2 = kc.ProteinSet.synthetic(40, law, seed, threads=1)
    assert np.array_equal(ps2.residues, ps.residues[:int(ps.offsets[40])])
    # family members stay close to member 0, later members drift further
    a, b = np.frombuffer(seqs[0].encode(), np.uint8), np.frombuffer(seqs[15].encode(), np.uint8)
    assert 0 < (a != b).mean() < 0.5


def test_synthetic_set_statistics():
    ps = kc.ProteinSet.synthetic(4000, "A", 0xB2000003, threads=4)
    lens = np.diff(ps.offsets.astype(np.int64))
    assert 330 < lens.mean() < 370
    assert len(set(ps.class_id.tolist())) == 15


# ---------------------------------------------------------------- C-ABI surface
def test_library_exports_every_declared_symbol():
    L = kc.lib()
    declared = set()
    for hdr in ("kc_b200.h", "kc_host.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(kc_[a-z0-9_]+)\s*\(", text))
    from uniprot_kmer_based_clustering_b200._lib import EXPORTED
    assert declared == set(EXPORTED)
    for name in declared:
        assert hasattr(L, name), name
    assert L.kc_abi_version() == 2
    # the benchmark generator is its own host-only library
    from uniprot_kmer_based_clustering_b200._lib import SYNTH_EXPORTED, synth_lib
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "kc_synth.h")).read(), flags=re.S)
    assert set(re.findall(r"\b(kc_[a-z0-9_]+)\s*\(", text)) == set(SYNTH_EXPORTED)
    for name in SYNTH_EXPORTED:
        assert hasattr(synth_lib(), name), name


def test_reference_api_mirror_names():
    from uniprot_kmer_based_clustering_b200 import graph, protein
    for name in ("get_amr_class", "get_five_mers", "get_five_hash", "get_id_and_seq"):
        assert hasattr(protein.Protein, name)
    for name in ("new", "remove_uninteresting_edges", "combine_edges", "align_and_output_pairs", "edges"):
        assert hasattr(graph.Graph, name)
    for name in ("get_kmers", "get_vertices_key", "get_proteins_ids_and_sequences"):
        assert hasattr(graph.KmerEdge, name)
    assert protein.five_mer_back_to_amino_acid(2644056) == "MKHKN"
    assert protein.five_mer_back_to_amino_acid(1166028867, 7) == "MKHKNQA"


def test_engine_fails_loudly_without_gpu_and_on_bad_k():
    cfg_bad = kc._lib.Config(6, 0, 10, 1, 0, 0, 0, 0)
    h = ctypes.c_void_p()
    assert kc.lib().kc_create(ctypes.byref(cfg_bad), ctypes.byref(h)) == kc._lib.KC_EINVAL
    if not has_gpu():
        with pytest.raises(kc.KcError) as ei:
            kc.Engine(5)
        assert ei.value.code == kc._lib.KC_ENODEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "uniprot_kmer_based_clustering_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in text.lower().replace("no cpu fallback", ""), (dirpath, fn)


# ---------------------------------------------------------------- subsampling mode (host side)
def test_position_sampler_is_a_permutation_and_matches_the_oracle():
    """kc_sample_position: distinct positions, exact count, reproducible (src/protein.rs:77-104 takes a
    tenth of the start positions without replacement)"""
    L = kc.lib()
    for n in (1, 2, 7, 10, 33, 100, 1000, 4097):
        for protein in (0, 5, 12345):
            pos = [L.kc_sample_position(0xB2005EED, protein, n, x) for x in range(n)]
            assert sorted(pos) == list(range(n))
    a = [L.kc_sample_position(1, 7, 500, x) for x in range(50)]
    b = [L.kc_sample_position(2, 7, 500, x) for x in range(50)]
    c = [L.kc_sample_position(1, 8, 500, x) for x in range(50)]
    assert a != b and a != c
    # the oracle's restatement of the sampler agrees: sampled k-mers = k-mers at the sampled positions
    ps = random_protein_set(5, 30, min_len=0, max_len=300, n_classes=2, family=3)
    full = Oracle(5, 1)
    full.set_proteins(ps.residues, ps.offsets, ps.class_id)
    fk = full.extract_kmers()
    samp = Oracle(5, 1, sample_every=10, sample_seed=0xB2005EED)
    samp.set_proteins(ps.residues, ps.offsets, ps.class_id)
    sk = samp.extract_kmers()
    lens = np.diff(ps.offsets.astype(np.int64))
    npos = np.maximum(lens - 4, 0)
    fo = np.concatenate([[0], np.cumsum(npos)])
    so = np.concatenate([[0], np.cumsum(npos // 10)])
    assert sk.size == so[-1]
    for p in range(ps.n):
        exp = [fk[fo[p] + L.kc_sample_position(0xB2005EED, p, int(npos[p]), x)] for x in range(int(npos[p]) // 10)]
        assert sk[so[p]:so[p + 1]].tolist() == exp


def test_diamond_handoff_files(tmp_path):
    """kc_write_handoff: the per-pair FASTA files and the blastp TSV header of align_and_output_pairs
    (src/graph/mod.rs:253-261,273-280,304-317); host only"""
    from uniprot_kmer_based_clustering_b200.engine import EDGE_DTYPE, write_handoff
    ps = kc.ProteinSet.from_fasta_bytes(b">A1|F|U|beta_lactam|bla extra\nMKHKNQA\n>B2|F|U|polymyxin|arnA\nMKH\nKNQATT\n"
                                        b">C3|F|U|beta_lactam|x\nAAAA\n")
    edges = np.array([(0, 1, 11, 0), (1, 2, 12, 0)], dtype=EDGE_DTYPE)
    assert write_handoff(ps, edges, str(tmp_path)) == 4
    assert (tmp_path / "fasta_files" / "0_A1.fasta").read_text() == ">A1|F|U|beta_lactam|bla\nMKHKNQA"
    assert (tmp_path / "fasta_files" / "0_B2.fasta").read_text() == ">B2|F|U|polymyxin|arnA\nMKHKNQATT"
    assert (tmp_path / "fasta_files" / "1_C3.fasta").read_text() == ">C3|F|U|beta_lactam|x\nAAAA"
    assert (tmp_path / "db_files").is_dir()
    hdr = (tmp_path / "blastp_output.tsv").read_text()
    assert hdr.startswith("query id\tquery length\tsubject id\t") and hdr.endswith("evalue\tbit score\n")
    assert hdr.count("\t") == 11


def test_rust_ffi_crate_binds_declared_symbols():
    """ffi/src/kc_sys.rs (shipped as source: no Rust toolchain in the image) must only bind entry points the
    headers declare, with the ABI version the library reports"""
    rs = open(os.path.join(ROOT, "ffi", "src", "kc_sys.rs")).read()
    bound = set(re.findall(r"pub fn (kc_[a-z0-9_]+)\s*\(", rs))
    declared = set()
    for hdr in ("kc_b200.h", "kc_host.h"):
        text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", hdr)).read(), flags=re.S)
        declared |= set(re.findall(r"\b(kc_[a-z0-9_]+)\s*\(", text))
    assert bound and bound <= declared, sorted(bound - declared)
    for must in ("kc_create", "kc_set_proteins", "kc_build_index", "kc_score_pairs", "kc_get_edges", "kc_comm_init",
                 "kc_build_index_dist", "kc_gather_edges", "kc_write_handoff"):
        assert must in bound, must
    assert f"KC_ABI_VERSION: c_int = {kc.lib().kc_abi_version()};" in rs
    # the Rust kc_config mirrors the C struct field for field
    c_fields = re.findall(r"^\s+(?:u?int\d+_t)\s+(\w+);", re.search(r"typedef struct kc_config \{(.*?)\} kc_config;", open(
        os.path.join(ROOT, "include", "kc_b200.h")).read(), flags=re.S).group(1), flags=re.M)
    r_fields = re.findall(r"pub (\w+): [iu]\d+,", re.search(r"pub struct kc_config \{(.*?)\n\}", rs, flags=re.S).group(1))
    assert c_fields == r_fields, (c_fields, r_fields)
