import json
import lzma
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    import uniprot_kmer_based_clustering_b200 as kc
    return kc.lib().kc_device_count() > 0


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "arg_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def arg_fasta_bytes():
    return lzma.open(os.path.join(GOLDEN_DIR, "arg_proteins.fasta.xz")).read()


@pytest.fixture(scope="session")
def arg_set(arg_fasta_bytes):
    """The ARG protein set staged by the PRODUCT parser (host code, no GPU needed)."""
    import uniprot_kmer_based_clustering_b200 as kc
    return kc.ProteinSet.from_fasta_bytes(arg_fasta_bytes, threads=4)


@pytest.fixture(scope="session")
def arg_oracle(arg_set):
    """Oracle results on the ARG set, computed once per session: {k: (kmers, index)}."""
    from oracle.oracle import Oracle
    out = {}
    for k in (5, 7):
        o = Oracle(k, threads=min(8, os.cpu_count() or 1))
        o.set_proteins(arg_set.residues, arg_set.offsets, arg_set.class_id)
        km = o.extract_kmers()
        ix = o.build_index()
        out[k] = (o, km, ix)
    return out


def sha16(a) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def edges_abc(e) -> np.ndarray:
    return np.stack([e["a"], e["b"], e["count"]], axis=1).astype("<u4")


def random_protein_set(seed: int, n: int, min_len: int = 0, max_len: int = 80, n_classes: int = 3,
                       family: int = 4, mutate: float = 0.08, alphabet: str = "ACDEFGHIKLMNPQRSTVWYXZ*a"):
    """Small family-structured random set with awkward residues; returns a ProteinSet."""
    import uniprot_kmer_based_clustering_b200 as kc
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    seqs, cls = [], []
    base = None
    for i in range(n):
        if i % family == 0 or base is None:
            L = int(rng.integers(min_len, max_len + 1))
            base = letters[rng.integers(0, 20, size=L)]
        s = base.copy()
        m = rng.random(s.size) < mutate * (i % family)
        s[m] = letters[rng.integers(0, letters.size, size=int(m.sum()))]
        seqs.append(s)
        cls.append(int(rng.integers(0, n_classes)))
    off = np.zeros(n + 1, dtype=np.uint64)
    if n:
        off[1:] = np.cumsum([s.size for s in seqs])
    res = np.concatenate(seqs).astype(np.uint8) if n and off[n] else np.zeros(0, dtype=np.uint8)
    ids = [f"R{i}|F|U|c{cls[i]}|g" for i in range(n)]
    return kc.ProteinSet(res, off, np.array(cls, dtype=np.uint32), ids, [f"c{c}" for c in range(n_classes)])
