"""ctypes binding of libkc_b200.so (include/kc_b200.h, include/kc_host.h).

The shared library is built in-tree by `make -C uniprot_kmer_based_clustering_b200/csrc`
(or `__graft_entry__.build()`).  There is no fallback: if the library is missing, loading
fails with instructions, and every compute call fails if no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libkc_b200.so")
CSRC = os.path.join(_PKG, "csrc")

KC_OK, KC_EINVAL, KC_ENODEVICE, KC_ECUDA, KC_ENOMEM, KC_ETOOLARGE = range(6)
ERROR_NAMES = {1: "KC_EINVAL", 2: "KC_ENODEVICE", 3: "KC_ECUDA", 4: "KC_ENOMEM", 5: "KC_ETOOLARGE"}


class KcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("k", C.c_int32), ("device", C.c_int32), ("threshold", C.c_uint32),
                ("cross_class_only", C.c_int32), ("want_blosum", C.c_int32), ("sample_every", C.c_uint32),
                ("max_edges", C.c_uint64), ("sample_seed", C.c_uint64),
                ("index_build", C.c_uint32), ("bucket_cap", C.c_uint32), ("index_slices", C.c_uint32),
                ("census_merge", C.c_uint32), ("pair_lists", C.c_uint32), ("no_upload_overlap", C.c_uint32)]


INDEX_BUILDS = {"auto": 0, "stream": 1, "bucket": 2, "table": 3}


class IndexStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_positions", "n_incidences", "n_distinct", "n_singleton", "n_repeated", "nnz")]


class PairStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_multi_edges", "n_multi_edges_kept", "n_pairs_kept", "n_edges_out", "sum_count_out",
                 "n_rows", "n_retries", "n_rows_rescored")]


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("extract_ms", C.c_float), ("index_ms", C.c_float),
                ("pairs_ms", C.c_float), ("edges_ms", C.c_float), ("d2h_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32),
                ("pair_kernel_ms", C.c_float), ("census_kernel_ms", C.c_float),
                ("index_records", C.c_uint64), ("index_mid_buckets", C.c_uint32), ("index_huge_buckets", C.c_uint32)]


def stats_dict(s: C.Structure) -> dict:
    return {n: getattr(s, n) for n, _ in s._fields_}


# every symbol the two headers declare; tests check the library exports all of them
EXPORTED = [
    "kc_sample_position", "kc_abi_version", "kc_device_count", "kc_create", "kc_destroy", "kc_last_error", "kc_set_stream",
    "kc_set_proteins", "kc_set_proteins_device", "kc_set_proteins_device_residues", "kc_extract_kmers", "kc_build_index", "kc_build_index_shard", "kc_index_shard_info", "kc_index_shard_blocks", "kc_index_flavour",
    "kc_get_distinct_kmers", "kc_get_vocab", "kc_get_protein_ids", "kc_lookup_kmers", "kc_get_pair_index", "kc_score_pairs",
    "kc_score_pairs_shard", "kc_get_edges", "kc_get_edges_device", "kc_get_edge_kmers", "kc_get_timings", "kc_reset_timings",
    "kc_bitset_pair_counts", "kc_popc_microbench",
    "kc_comm_unique_id", "kc_comm_init", "kc_comm_info", "kc_set_proteins_dist", "kc_build_index_dist",
    "kc_score_pairs_dist", "kc_gather_edges", "kc_gather_edges_shared",
    "kc_fasta_parse_file", "kc_fasta_parse_buffer", "kc_fasta_free", "kc_fasta_n_proteins",
    "kc_fasta_n_residues", "kc_fasta_residues", "kc_fasta_offsets", "kc_fasta_class_ids",
    "kc_fasta_n_classes", "kc_fasta_n_missing_class", "kc_fasta_class_name", "kc_fasta_id",
    "kc_write_handoff", "kc_blastp_header",
    "kc_tree_build", "kc_tree_free", "kc_tree_n_merges", "kc_tree_n_no_common", "kc_tree_serialize",
    "kc_tree_clusters",
]


def build(verbose: bool = False) -> str:
    """Compile the CUDA engine for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "all"], stdout=out)
    return LIB_PATH


SYNTH_LIB_PATH = os.path.join(_PKG, "lib", "libkc_synth.so")
SYNTH_EXPORTED = ["kc_synth_layout", "kc_synth_residues"]
_synth = None


def synth_lib():
    """libkc_synth.so (include/kc_synth.h): the benchmark generator, host only.  Loading it does not
    load the engine."""
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise ImportError(f"{SYNTH_LIB_PATH} is missing: build it with `make -C {CSRC}`")
        L = C.CDLL(SYNTH_LIB_PATH)
        L.kc_synth_layout.restype = C.c_int
        L.kc_synth_layout.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.kc_synth_residues.restype = C.c_int
        L.kc_synth_residues.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        _synth = L
    return _synth


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (needs nvcc, sm_100a). "
            "There is no CPU fallback for the k-mer clustering engine.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32, cp = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_char_p
    P = C.POINTER
    sig = {
        "kc_sample_position": (u32, [u64, u32, u32, u32]),
        "kc_abi_version": (i32, []),
        "kc_device_count": (i32, []),
        "kc_create": (i32, [P(Config), P(vp)]),
        "kc_destroy": (None, [vp]),
        "kc_last_error": (cp, [vp]),
        "kc_set_stream": (i32, [vp, vp]),
        "kc_set_proteins": (i32, [vp, vp, vp, vp, u64]),
        "kc_set_proteins_device": (i32, [vp, vp, vp, vp, u64]),
        "kc_set_proteins_device_residues": (i32, [vp, vp, vp, vp, u64]),
        "kc_extract_kmers": (i32, [vp, vp, u64, P(u64)]),
        "kc_build_index": (i32, [vp, P(IndexStats)]),
        "kc_get_distinct_kmers": (i32, [vp, vp, u64]),
        "kc_get_vocab": (i32, [vp, vp, vp, u64]),
        "kc_get_protein_ids": (i32, [vp, vp, vp, u64]),
        "kc_lookup_kmers": (i32, [vp, vp, u64, vp]),
        "kc_score_pairs": (i32, [vp, P(PairStats)]),
        "kc_score_pairs_shard": (i32, [vp, u32, u32, P(PairStats)]),
        "kc_get_edges": (i32, [vp, vp, u64]),
        "kc_get_edges_device": (i32, [vp, P(vp), P(u64)]),
        "kc_get_edge_kmers": (i32, [vp, u64, vp, u64]),
        "kc_get_pair_index": (i32, [vp, vp, vp, vp, u64, vp, vp, u64]),
        "kc_build_index_shard": (i32, [vp, u32, u32, P(IndexStats)]),
        "kc_index_shard_info": (i32, [vp, vp]),
        "kc_index_shard_blocks": (i32, [vp, vp, u32]),
        "kc_index_flavour": (i32, [vp]),
        "kc_get_timings": (i32, [vp, P(Timings)]),
        "kc_reset_timings": (i32, [vp]),
        "kc_bitset_pair_counts": (i32, [vp, vp, u32, vp]),
        "kc_popc_microbench": (i32, [vp, u32, P(C.c_double)]),
        "kc_comm_unique_id": (i32, [vp]),
        "kc_comm_init": (i32, [vp, vp, i32, i32]),
        "kc_comm_info": (i32, [vp, P(i32), P(i32)]),
        "kc_set_proteins_dist": (i32, [vp, vp, vp, vp, u64]),
        "kc_build_index_dist": (i32, [vp, P(IndexStats)]),
        "kc_score_pairs_dist": (i32, [vp, P(PairStats)]),
        "kc_gather_edges": (i32, [vp, vp, u64, P(u64)]),
        "kc_gather_edges_shared": (i32, [vp, vp, u64, P(u64)]),
        "kc_fasta_parse_file": (i32, [cp, i32, P(vp)]),
        "kc_fasta_parse_buffer": (i32, [cp, u64, i32, P(vp)]),
        "kc_fasta_free": (None, [vp]),
        "kc_fasta_n_proteins": (u64, [vp]),
        "kc_fasta_n_residues": (u64, [vp]),
        "kc_fasta_residues": (vp, [vp]),
        "kc_fasta_offsets": (vp, [vp]),
        "kc_fasta_class_ids": (vp, [vp]),
        "kc_fasta_n_classes": (u32, [vp]),
        "kc_fasta_n_missing_class": (u64, [vp]),
        "kc_fasta_class_name": (cp, [vp, u32]),
        "kc_fasta_id": (cp, [vp, u64]),
        "kc_write_handoff": (i32, [vp, vp, u64, cp, P(u64)]),
        "kc_blastp_header": (cp, []),
        "kc_tree_build": (i32, [vp, vp, u64, u32, P(vp)]),
        "kc_tree_free": (None, [vp]),
        "kc_tree_n_merges": (u64, [vp]),
        "kc_tree_n_no_common": (u64, [vp]),
        "kc_tree_serialize": (u64, [vp, vp, u64]),
        "kc_tree_clusters": (i32, [vp, vp, P(u32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
