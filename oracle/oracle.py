"""ctypes front end of the CPU oracle (oracle/kc_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkc_oracle.so")

EDGE_DTYPE = np.dtype([("a", "<u4"), ("b", "<u4"), ("count", "<u4"), ("blosum", "<i4")])


class _IndexStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_positions", "n_incidences", "n_distinct", "n_singleton", "n_repeated", "nnz")]


class _PairStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_multi_edges", "n_multi_edges_kept", "n_pairs_kept", "n_edges_out", "sum_count_out")]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "kc_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libkc_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.ko_create.restype = vp
        L.ko_create.argtypes = [i32, i32]
        L.ko_destroy.argtypes = [vp]
        L.ko_set_sampling.argtypes = [vp, u32, u64]
        L.ko_last_error.restype = C.c_char_p
        L.ko_last_error.argtypes = [vp]
        L.ko_set_proteins.argtypes = [vp, vp, vp, vp, u64]
        L.ko_extract.argtypes = [vp]
        L.ko_n_positions.restype = u64
        L.ko_n_positions.argtypes = [vp]
        L.ko_get_kmers.argtypes = [vp, vp]
        L.ko_build_index.argtypes = [vp, C.POINTER(_IndexStats)]
        L.ko_n_distinct.restype = u64
        L.ko_n_distinct.argtypes = [vp]
        L.ko_get_distinct.argtypes = [vp, vp, vp]
        L.ko_get_vocab.argtypes = [vp, vp, vp]
        L.ko_get_protein_ids.argtypes = [vp, vp, vp]
        L.ko_score_pairs.argtypes = [vp, u32, i32, i32, i32, u64, u64, C.POINTER(_PairStats)]
        L.ko_n_edges.restype = u64
        L.ko_n_edges.argtypes = [vp]
        L.ko_get_edges.argtypes = [vp, vp]
        L.ko_get_times.argtypes = [vp, vp]
        L.ko_self_score.restype = i32
        L.ko_self_score.argtypes = [u32, i32]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class IndexResult:
    stats: dict
    distinct: np.ndarray
    distinct_counts: np.ndarray
    vocab: np.ndarray
    freq: np.ndarray
    row_offsets: np.ndarray
    ids: np.ndarray


@dataclass
class PairResult:
    stats: dict
    edges: np.ndarray  # EDGE_DTYPE, sorted by (a, b)


class Oracle:
    """CPU restatement of the hot path; stage names follow the reference's modules."""

    def __init__(self, k: int = 5, threads: int = 1, sample_every: int = 0, sample_seed: int = 0):
        self._L = lib()
        self._h = self._L.ko_create(k, threads)
        if not self._h:
            raise ValueError("k must be 5 or 7")
        if sample_every > 1:
            self._L.ko_set_sampling(self._h, sample_every, sample_seed)
        self.k = k
        self.n = 0

    def close(self):
        if self._h:
            self._L.ko_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_proteins(self, residues: np.ndarray, offsets: np.ndarray, class_id: np.ndarray):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        class_id = np.ascontiguousarray(class_id, dtype=np.uint32)
        assert offsets.size == class_id.size + 1
        self.n = int(class_id.size)
        self._L.ko_set_proteins(self._h, _ptr(residues), _ptr(offsets), _ptr(class_id), self.n)

    def extract_kmers(self) -> np.ndarray:
        self._L.ko_extract(self._h)
        out = np.empty(self._L.ko_n_positions(self._h), dtype=np.uint32)
        self._L.ko_get_kmers(self._h, _ptr(out))
        return out

    def build_index(self) -> IndexResult:
        st = _IndexStats()
        self._L.ko_build_index(self._h, C.byref(st))
        stats = {n: int(getattr(st, n)) for n, _ in _IndexStats._fields_}
        d = np.empty(stats["n_distinct"], dtype=np.uint32)
        dc = np.empty(stats["n_distinct"], dtype=np.uint32)
        self._L.ko_get_distinct(self._h, _ptr(d), _ptr(dc))
        v = np.empty(stats["n_repeated"], dtype=np.uint32)
        f = np.empty(stats["n_repeated"], dtype=np.uint32)
        self._L.ko_get_vocab(self._h, _ptr(v), _ptr(f))
        ro = np.empty(self.n + 1, dtype=np.uint64)
        ids = np.empty(stats["nnz"], dtype=np.uint32)
        self._L.ko_get_protein_ids(self._h, _ptr(ro), _ptr(ids))
        return IndexResult(stats, d, dc, v, f, ro, ids)

    def score_pairs(self, threshold: int = 10, cross_class_only: bool = True,
                    want_blosum: bool = True, mode: int = 1,
                    row_lo: int = 0, row_hi: int = 0) -> PairResult:
        st = _PairStats()
        rc = self._L.ko_score_pairs(self._h, threshold, int(cross_class_only), int(want_blosum),
                                    mode, row_lo, row_hi, C.byref(st))
        if rc != 0:
            raise RuntimeError(self._L.ko_last_error(self._h).decode())
        stats = {n: int(getattr(st, n)) for n, _ in _PairStats._fields_}
        e = np.empty(self._L.ko_n_edges(self._h), dtype=EDGE_DTYPE)
        self._L.ko_get_edges(self._h, _ptr(e))
        return PairResult(stats, e)

    def times(self) -> dict:
        t = np.zeros(3, dtype=np.float64)
        self._L.ko_get_times(self._h, _ptr(t))
        return {"extract_s": float(t[0]), "index_s": float(t[1]), "pairs_s": float(t[2])}


def self_score(kmer: int, k: int) -> int:
    return int(lib().ko_self_score(kmer, k))
