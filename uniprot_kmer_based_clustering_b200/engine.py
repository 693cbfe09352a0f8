"""Python host side of the C ABI: `Engine` (one per GPU), FASTA staging, synthetic sets.

Everything here is plumbing over libkc_b200.so: numpy arrays in, numpy arrays out.  No
compute happens in Python and nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import Config, IndexStats, KcError, PairStats, Timings, stats_dict

EDGE_DTYPE = np.dtype([("a", "<u4"), ("b", "<u4"), ("count", "<u4"), ("blosum", "<i4")])

# frozen seeds of the synthetic configurations (BASELINE.json configs 3-5)
SYNTH_SEEDS = {"synth_100k_k5": 0xB2000003, "synth_1m_k7": 0xB2000004, "synth_4m_skew": 0xB2000005}


# Defaults of the engine's tuning knobs for engines created without explicit arguments (kc_config fields;
# the C library reads nothing from the environment).  The GPU test-suite swaps these to run every test
# against every index build.
DEFAULTS = {"index_build": "auto", "bucket_cap": 0, "index_slices": 0, "census_merge": 0, "pair_lists": False,
            "no_upload_overlap": False}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class ProteinSet:
    """The staged input: what `Vec<Protein>` holds after src/main.rs:62-72."""
    residues: np.ndarray            # u8, all sequences back to back
    offsets: np.ndarray             # u64[n+1]
    class_id: np.ndarray            # u32[n]
    ids: list = field(default_factory=list)
    class_names: list = field(default_factory=list)
    n_missing_class: int = 0

    @property
    def n(self) -> int:
        return int(self.class_id.size)

    def seq(self, p: int) -> str:
        return self.residues[int(self.offsets[p]):int(self.offsets[p + 1])].tobytes().decode()

    @staticmethod
    def from_fasta(path: str, threads: int = 1) -> "ProteinSet":
        L = _lib.lib()
        h = C.c_void_p()
        rc = L.kc_fasta_parse_file(path.encode(), threads, C.byref(h))
        if rc != 0:
            raise KcError(rc, f"cannot read FASTA {path!r}")
        return ProteinSet._take(L, h)

    @staticmethod
    def from_fasta_bytes(data: bytes, threads: int = 1) -> "ProteinSet":
        L = _lib.lib()
        h = C.c_void_p()
        rc = L.kc_fasta_parse_buffer(data, len(data), threads, C.byref(h))
        if rc != 0:
            raise KcError(rc, "cannot parse FASTA buffer")
        return ProteinSet._take(L, h)

    @staticmethod
    def _take(L, h) -> "ProteinSet":
        try:
            n = L.kc_fasta_n_proteins(h)
            nres = L.kc_fasta_n_residues(h)
            def view(ptr, dtype, count):
                if not ptr or count == 0:
                    return np.zeros(count, dtype=dtype)
                return np.frombuffer(C.string_at(ptr, count * np.dtype(dtype).itemsize), dtype=dtype).copy()

            res = view(L.kc_fasta_residues(h), np.uint8, nres)
            off = view(L.kc_fasta_offsets(h), np.uint64, n + 1)
            cls = view(L.kc_fasta_class_ids(h), np.uint32, n)
            ids = [L.kc_fasta_id(h, i).decode() for i in range(n)]
            names = [L.kc_fasta_class_name(h, c).decode() for c in range(L.kc_fasta_n_classes(h))]
            missing = L.kc_fasta_n_missing_class(h)
        finally:
            L.kc_fasta_free(h)
        return ProteinSet(res, off, cls, ids, names, int(missing))

    @staticmethod
    def synthetic(n: int, length_law: str = "A", seed: int = 0xB2000003, threads: int = 8,
                  with_ids: bool = False) -> "ProteinSet":
        """Generator G1 (include/kc_synth.h, libkc_synth.so: host only, the engine library is not loaded)."""
        L = _lib.synth_lib()
        law = {"A": 0, "B": 1}[length_law]
        off = np.zeros(n + 1, dtype=np.uint64)
        cls = np.zeros(max(n, 1), dtype=np.uint32)[:n]
        rc = L.kc_synth_layout(n, law, seed, _ptr(off), _ptr(cls))
        if rc != 0:
            raise KcError(rc, "kc_synth_layout")
        res = np.empty(int(off[n]), dtype=np.uint8)
        rc = L.kc_synth_residues(n, law, seed, threads, _ptr(off), _ptr(res))
        if rc != 0:
            raise KcError(rc, "kc_synth_residues")
        ids = ([f"S{i}|FEATURES|SYNTH|class{int(cls[i])}|fam{i // 16}" for i in range(n)]
               if with_ids else [])
        return ProteinSet(res, off, cls, ids, [f"class{c}" for c in range(15)])

    def to_fasta_bytes(self) -> bytes:
        out = []
        for p in range(self.n):
            out.append(b">" + self.ids[p].encode() + b"\n" +
                       self.residues[int(self.offsets[p]):int(self.offsets[p + 1])].tobytes() + b"\n")
        return b"".join(out)


class Engine:
    """One engine per GPU.  Method names follow include/kc_b200.h."""

    def __init__(self, k: int = 5, device: int = 0, threshold: int = 10, cross_class_only: bool = True,
                 want_blosum: bool = False, max_edges: int = 0, sample_every: int = 0, sample_seed: int = 0,
                 index_build: str | None = None, bucket_cap: int | None = None, index_slices: int | None = None,
                 census_merge: int | None = None, pair_lists: bool | None = None,
                 no_upload_overlap: bool | None = None):
        """index_build: "auto" | "stream" | "bucket" | "table" (kc_config.index_build); the rest are the
        tuning knobs of include/kc_b200.h (0 = the engine's own choice; None = DEFAULTS)"""
        self._L = _lib.lib()
        self._h = C.c_void_p()
        index_build = DEFAULTS["index_build"] if index_build is None else index_build
        bucket_cap = DEFAULTS["bucket_cap"] if bucket_cap is None else bucket_cap
        index_slices = DEFAULTS["index_slices"] if index_slices is None else index_slices
        census_merge = DEFAULTS["census_merge"] if census_merge is None else census_merge
        pair_lists = DEFAULTS["pair_lists"] if pair_lists is None else pair_lists
        no_upload_overlap = DEFAULTS["no_upload_overlap"] if no_upload_overlap is None else no_upload_overlap
        cfg = Config(k, device, threshold, int(cross_class_only), int(want_blosum), sample_every, max_edges,
                     sample_seed, _lib.INDEX_BUILDS[index_build], bucket_cap, index_slices, census_merge,
                     int(pair_lists), int(no_upload_overlap))
        rc = self._L.kc_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            self._h = None
            raise KcError(rc, "kc_create: no usable CUDA device (there is no CPU fallback)"
                          if rc == _lib.KC_ENODEVICE else "kc_create")
        self.k, self.device = k, device
        self.threshold, self.cross_class_only, self.want_blosum = threshold, bool(cross_class_only), bool(want_blosum)
        self.n = 0
        self.index_stats: dict = {}
        self.pair_stats: dict = {}
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.kc_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise KcError(rc, self._L.kc_last_error(self._h).decode())

    # ---- staging --------------------------------------------------------------------
    def set_proteins(self, residues: np.ndarray, offsets: np.ndarray, class_id: np.ndarray):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        class_id = np.ascontiguousarray(class_id, dtype=np.uint32)
        if offsets.size != class_id.size + 1:
            raise ValueError("offsets must have n+1 entries")
        self.n = int(class_id.size)
        # the engine borrows the three arrays until the next build_index / extract_kmers returns (asynchronous
        # upload, host copies made inside the build): keep them (or the converted copies made above) alive
        self._keep = [residues, offsets, class_id]
        self._check(self._L.kc_set_proteins(self._h, _ptr(residues), _ptr(offsets), _ptr(class_id), self.n))

    def set_proteins_ptr(self, residues_ptr: int, offsets_ptr: int, class_ptr: int, n: int, on_device: bool):
        """Raw-pointer variant (pinned host buffers or device buffers owned by the caller)."""
        self.n = int(n)
        fn = self._L.kc_set_proteins_device if on_device else self._L.kc_set_proteins
        self._check(fn(self._h, C.c_void_p(residues_ptr), C.c_void_p(offsets_ptr), C.c_void_p(class_ptr), n))

    def set_proteins_device_residues(self, d_residues_ptr: int, offsets_ptr: int, class_ptr: int, n: int):
        """residue stream in HBM (device pointer), offsets / classes on the host (pointers)"""
        self.n = int(n)
        self._check(self._L.kc_set_proteins_device_residues(self._h, C.c_void_p(d_residues_ptr), C.c_void_p(offsets_ptr),
                                                            C.c_void_p(class_ptr), n))

    def set_protein_set(self, ps: ProteinSet):
        self.set_proteins(ps.residues, ps.offsets, ps.class_id)

    def set_stream(self, cuda_stream: int):
        self._check(self._L.kc_set_stream(self._h, C.c_void_p(cuda_stream)))

    # ---- K1 -------------------------------------------------------------------------
    def extract_kmers(self, readback: bool = True) -> np.ndarray | int:
        npos = C.c_uint64()
        self._check(self._L.kc_extract_kmers(self._h, None, 0, C.byref(npos)))
        if not readback:
            return int(npos.value)
        out = np.empty(npos.value, dtype=np.uint32)
        self._check(self._L.kc_extract_kmers(self._h, _ptr(out), out.size, C.byref(npos)))
        return out

    # ---- K2-K5 ----------------------------------------------------------------------
    def build_index(self, shard: int = 0, n_shards: int = 1) -> dict:
        """n_shards > 1: the index of one row block of the pair triangle (kc_build_index_shard);
        the stats then add up to the whole-set numbers over the shards"""
        st = IndexStats()
        self._check(self._L.kc_build_index_shard(self._h, shard, n_shards, C.byref(st)))
        self.index_stats = stats_dict(st)
        return self.index_stats

    def index_shard_info(self) -> dict:
        """what the current index covers; n_shards == 1: a whole index (whole-set stats)"""
        info = np.zeros(4, dtype=np.uint32)
        self._check(self._L.kc_index_shard_info(self._h, _ptr(info)))
        bounds = np.zeros(int(info[2]) + 1, dtype=np.uint32)
        self._check(self._L.kc_index_shard_blocks(self._h, _ptr(bounds), bounds.size))
        return {"shard": int(info[0]), "n_shards": int(info[1]), "n_blocks": int(info[2]),
                "n_own_rows": int(info[3]), "block_bounds": bounds}

    def index_flavour(self) -> int:
        """0: universe-table build; 1: streaming partitioned build; 4096 / 8192: round 1's partitioned
        build with that bucket slot size"""
        return int(self._L.kc_index_flavour(self._h))

    def get_distinct_kmers(self) -> np.ndarray:
        out = np.empty(self.index_stats["n_distinct"], dtype=np.uint32)
        self._check(self._L.kc_get_distinct_kmers(self._h, _ptr(out), out.size))
        return out

    def get_vocab(self):
        v = np.empty(self.index_stats["n_repeated"], dtype=np.uint32)
        f = np.empty(self.index_stats["n_repeated"], dtype=np.uint32)
        self._check(self._L.kc_get_vocab(self._h, _ptr(v), _ptr(f), v.size))
        return v, f

    def get_protein_ids(self):
        ro = np.empty(self.n + 1, dtype=np.uint64)
        ids = np.empty(self.index_stats["nnz"], dtype=np.uint32)
        self._check(self._L.kc_get_protein_ids(self._h, _ptr(ro), _ptr(ids), ids.size))
        return ro, ids

    def get_pair_index(self):
        """(kmers[id], freq[id], self_score[id], row_offsets, ids): the index as the pair stage reads it"""
        V, nnz = self.index_stats["n_repeated"], self.index_stats["nnz"]
        v = np.empty(V, dtype=np.uint32)
        f = np.empty(V, dtype=np.uint32)
        ss = np.empty(V, dtype=np.uint8)
        ro = np.empty(self.n + 1, dtype=np.uint64)
        ids = np.empty(nnz, dtype=np.uint32)
        self._check(self._L.kc_get_pair_index(self._h, _ptr(v), _ptr(f), _ptr(ss), V, _ptr(ro), _ptr(ids), nnz))
        return v, f, ss, ro, ids

    def lookup_kmers(self, kmers: np.ndarray) -> np.ndarray:
        kmers = np.ascontiguousarray(kmers, dtype=np.uint32)
        out = np.empty(kmers.size, dtype=np.uint32)
        self._check(self._L.kc_lookup_kmers(self._h, _ptr(kmers), kmers.size, _ptr(out)))
        return out

    # ---- K7-K9 ----------------------------------------------------------------------
    def score_pairs(self, shard: int = 0, n_shards: int = 1) -> dict:
        st = PairStats()
        self._check(self._L.kc_score_pairs_shard(self._h, shard, n_shards, C.byref(st)))
        self.pair_stats = stats_dict(st)
        return self.pair_stats

    def get_edges(self) -> np.ndarray:
        out = np.empty(self.pair_stats["n_edges_out"], dtype=EDGE_DTYPE)
        self._check(self._L.kc_get_edges(self._h, _ptr(out), out.size))
        return out

    def edges_device(self):
        """(device pointer, n_edges) of the sorted edge list in HBM (16 bytes per edge)"""
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._L.kc_get_edges_device(self._h, C.byref(p), C.byref(n)))
        return (p.value or 0), int(n.value)

    def get_edges_into(self, out_ptr: int, capacity: int):
        self._check(self._L.kc_get_edges(self._h, C.c_void_p(out_ptr), capacity))

    def get_edge_kmers(self, edge_index: int, count: int) -> np.ndarray:
        out = np.empty(count, dtype=np.uint32)
        self._check(self._L.kc_get_edge_kmers(self._h, edge_index, _ptr(out), out.size))
        return out

    def bitset_pair_counts(self, rows: np.ndarray) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        out = np.zeros((rows.size, rows.size), dtype=np.uint32)
        self._check(self._L.kc_bitset_pair_counts(self._h, _ptr(rows), rows.size, _ptr(out)))
        return out

    def popc_microbench(self, iters: int = 4096) -> float:
        """AND + POPC + ADD issue rate, 10^9 operations per second (kc_popc_microbench)"""
        v = C.c_double()
        self._check(self._L.kc_popc_microbench(self._h, iters, C.byref(v)))
        return float(v.value)

    # ---- multi-GPU (NCCL below the C ABI, csrc/dist.cuh) --------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128 bytes made by rank 0 (ncclGetUniqueId); the host distributes them to every rank"""
        buf = C.create_string_buffer(128)
        rc = _lib.lib().kc_comm_unique_id(buf)
        if rc != 0:
            raise KcError(rc, "kc_comm_unique_id: NCCL is not available")
        return buf.raw

    def comm_init(self, comm_id: bytes, rank: int, world: int):
        self._check(self._L.kc_comm_init(self._h, C.c_char_p(comm_id), rank, world))
        self.rank, self.world = rank, world

    def set_proteins_dist_ptr(self, residues_ptr: int, offsets_ptr: int, class_ptr: int, n: int):
        """host pointers, the same arrays on every rank: 1 / world is uploaded, the rest all-gathered"""
        self.n = int(n)
        self._check(self._L.kc_set_proteins_dist(self._h, C.c_void_p(residues_ptr), C.c_void_p(offsets_ptr),
                                                 C.c_void_p(class_ptr), n))

    def build_index_dist(self) -> dict:
        st = IndexStats()
        self._check(self._L.kc_build_index_dist(self._h, C.byref(st)))
        self.index_stats = stats_dict(st)
        return self.index_stats

    def score_pairs_dist(self) -> dict:
        st = PairStats()
        self._check(self._L.kc_score_pairs_dist(self._h, C.byref(st)))
        self.pair_stats = stats_dict(st)  # whole-job counters
        return self.pair_stats

    def gather_edges_into(self, out_ptr: int, capacity: int, shared: bool = False) -> int:
        """the edge lists of all ranks as one sorted list: into rank 0's buffer, or (shared) into one host
        buffer every rank maps; returns the total number of edges"""
        n = C.c_uint64()
        fn = self._L.kc_gather_edges_shared if shared else self._L.kc_gather_edges
        self._check(fn(self._h, C.c_void_p(out_ptr) if out_ptr else None, capacity, C.byref(n)))
        return int(n.value)

    # ---- timing ---------------------------------------------------------------------
    def timings(self) -> dict:
        t = Timings()
        self._check(self._L.kc_get_timings(self._h, C.byref(t)))
        d = stats_dict(t)
        d.pop("reserved", None)
        return d

    def reset_timings(self):
        self._check(self._L.kc_reset_timings(self._h))


def write_handoff(ps: ProteinSet, edges: np.ndarray, directory: str) -> int:
    """The files align_and_output_pairs leaves for DIAMOND (kc_write_handoff, include/kc_host.h); returns the
    number of FASTA files written.  Host only."""
    L = _lib.lib()
    data = ps.to_fasta_bytes()
    h = C.c_void_p()
    rc = L.kc_fasta_parse_buffer(data, len(data), 1, C.byref(h))
    if rc != 0:
        raise KcError(rc, "cannot re-stage the protein set")
    try:
        edges = np.ascontiguousarray(edges, dtype=EDGE_DTYPE)
        n_files = C.c_uint64()
        rc = L.kc_write_handoff(h, _ptr(edges), edges.size, directory.encode(), C.byref(n_files))
        if rc != 0:
            raise KcError(rc, f"kc_write_handoff({directory!r})")
        return int(n_files.value)
    finally:
        L.kc_fasta_free(h)


def cluster(ps: ProteinSet, k: int = 5, threshold: int = 10, cross_class_only: bool = True,
            want_blosum: bool = False, device: int = 0):
    """The whole hot path on one GPU: returns (index_stats, pair_stats, edges)."""
    with Engine(k, device, threshold, cross_class_only, want_blosum) as e:
        e.set_protein_set(ps)
        ist = e.build_index()
        pst = e.score_pairs()
        return ist, pst, e.get_edges()
