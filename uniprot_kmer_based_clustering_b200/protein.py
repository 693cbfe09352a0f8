"""Mirror of the reference's `protein` module (src/protein.rs) over the GPU engine.

The reference builds one `Protein` per FASTA record and mutates it in place; here the engine
holds the whole set in HBM and `Protein` is a read-only view of one row.  Method names and
meanings follow src/protein.rs:107-179 (k generalised from "five" to the engine's k).
"""
from __future__ import annotations

import numpy as np

from .engine import Engine, ProteinSet

AMINO_ACID_LIST = "CSTAGPDEQNHRKMILVWYF*"  # src/protein.rs:9-13


def five_mer_back_to_amino_acid(kmer: int, k: int = 5) -> str:
    """src/protein.rs:38-48"""
    out = []
    for i in range(k):
        p = 21 ** (k - 1 - i)
        out.append(AMINO_ACID_LIST[kmer // p])
        kmer %= p
    return "".join(out)


class Mphf:
    """Stands in for boomphf::Mphf<u32> (src/main.rs:139-140): `hash(kmer)` is the canonical id
    of a repeated k-mer (its rank among the repeated k-mers in ascending order)."""

    def __init__(self, engine: Engine):
        self._e = engine

    def hash(self, kmer) -> np.ndarray | int:
        ids = self._e.lookup_kmers(np.atleast_1d(np.asarray(kmer, dtype=np.uint32)))
        if np.ndim(kmer) == 0:
            if ids[0] == 0xFFFFFFFF:
                raise KeyError(f"k-mer {int(kmer)} is not a repeated k-mer")
            return int(ids[0])
        return ids


class ProteinList:
    """`Arc<Vec<Arc<Protein>>>` after the rewrite stage (src/main.rs:204-212)."""

    def __init__(self, engine: Engine, ps: ProteinSet):
        self.engine, self.set = engine, ps
        engine.set_protein_set(ps)
        self._kmers = None
        self._kpos = None
        self._rows = None
        self.kmer_freq = None

    def __len__(self):
        return self.set.n

    def __getitem__(self, i: int) -> "Protein":
        if not 0 <= i < self.set.n:
            raise IndexError(i)
        return Protein(self, i)

    def build_index(self) -> dict:
        """census + split + index + rewrite, src/main.rs:84-199"""
        st = self.engine.build_index()
        _, self.kmer_freq = self.engine.get_vocab()
        self._rows = None
        return st

    def _positions(self):
        if self._kmers is None:
            self._kmers = self.engine.extract_kmers()
            k = self.engine.k
            lens = np.diff(self.set.offsets.astype(np.int64))
            npos = np.maximum(lens - k + 1, 0)
            self._kpos = np.concatenate([[0], np.cumsum(npos)])
        return self._kmers, self._kpos

    def _id_rows(self):
        if self._rows is None:
            self._rows = self.engine.get_protein_ids()
        return self._rows


class Protein:
    def __init__(self, plist: ProteinList, index: int):
        self._l, self._i = plist, index

    def get_amr_class(self) -> str:
        """src/protein.rs:135-138 (panics in the reference when the id has < 4 fields)"""
        f = self._l.set.ids[self._i].split("|")
        if f and f[-1] == "":
            f.pop()
        return f[3]

    def get_five_mers(self) -> np.ndarray:
        """src/protein.rs:141-143: one k-mer per start position, duplicates kept"""
        km, kp = self._l._positions()
        return km[kp[self._i]:kp[self._i + 1]].copy()

    def get_five_hash(self) -> np.ndarray:
        """src/protein.rs:146-148: ids of the protein's repeated k-mers (ascending)"""
        ro, ids = self._l._id_rows()
        return ids[int(ro[self._i]):int(ro[self._i + 1])].copy()

    def get_id_and_seq(self):
        """src/protein.rs:177-179"""
        return self._l.set.ids[self._i], self._l.set.seq(self._i)
