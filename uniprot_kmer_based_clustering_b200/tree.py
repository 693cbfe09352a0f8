"""Mirror of the reference's `tree` module (src/tree.rs) over the host tree builder (csrc/tree.cpp).

    tree = Tree.from_id_rows(row_offsets, ids, n_ids)     # Tree::new + add_protein in input order
    tree.clusters()                                       # top-level clusters (the root's children)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KcError


class Tree:
    def __init__(self, handle, n_proteins: int):
        self._L = _lib.lib()
        self._h = handle
        self.n_proteins = n_proteins

    @staticmethod
    def from_id_rows(row_offsets: np.ndarray, ids: np.ndarray, n_ids: int) -> "Tree":
        """row_offsets[n+1], ids ascending within a row (Engine.get_protein_ids)"""
        L = _lib.lib()
        ro = np.ascontiguousarray(row_offsets, dtype=np.uint64)
        ii = np.ascontiguousarray(ids, dtype=np.uint32)
        h = C.c_void_p()
        rc = L.kc_tree_build(ro.ctypes.data_as(C.c_void_p), ii.ctypes.data_as(C.c_void_p), ro.size - 1, n_ids,
                             C.byref(h))
        if rc != 0:
            raise KcError(rc, "kc_tree_build: id rows must be ascending and below n_ids")
        return Tree(h, int(ro.size - 1))

    @staticmethod
    def from_engine(engine) -> "Tree":
        """Tree over the engine's index (src/tree.rs:524-536 with `kmer_size` = the engine's k)"""
        ro, ids = engine.get_protein_ids()
        return Tree.from_id_rows(ro, ids, engine.index_stats["n_repeated"])

    def close(self):
        if getattr(self, "_h", None):
            self._L.kc_tree_free(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def n_merges(self) -> int:
        return int(self._L.kc_tree_n_merges(self._h))

    @property
    def n_no_common(self) -> int:
        return int(self._L.kc_tree_n_no_common(self._h))

    def serialize(self) -> np.ndarray:
        n = self._L.kc_tree_serialize(self._h, None, 0)
        out = np.zeros(n, dtype=np.int64)
        self._L.kc_tree_serialize(self._h, out.ctypes.data_as(C.c_void_p), n)
        return out

    def nested(self):
        """nested-list form (children in order), leaves are protein indices"""
        toks = self.serialize().tolist()
        pos = 0

        def rec():
            nonlocal pos
            t = toks[pos]
            pos += 1
            if t >= 0:
                return t
            return [rec() for _ in range(-t)]

        return rec() if toks else None

    def clusters(self) -> np.ndarray:
        """cluster_of[p] = index of the top-level cluster holding protein p"""
        out = np.zeros(max(self.n_proteins, 1), dtype=np.uint32)
        n = C.c_uint32()
        rc = self._L.kc_tree_clusters(self._h, out.ctypes.data_as(C.c_void_p), C.byref(n))
        if rc != 0:
            raise KcError(rc, "kc_tree_clusters")
        self.n_clusters = int(n.value)
        return out[:self.n_proteins]
