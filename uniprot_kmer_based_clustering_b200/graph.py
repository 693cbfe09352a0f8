"""Mirror of the reference's `graph` module (src/graph/mod.rs, edge.rs) over the GPU engine.

The reference materialises the multigraph and refines it in three passes; the engine scores
all pairs in one pass and never builds the multigraph.  The methods below keep the reference's
call sequence and print its counters (same wording, on stderr) so a maintainer can swap the
module in:

    graph = Graph.new(kmer_freq, threads, protein_list)      # src/main.rs:216-218
    graph.remove_uninteresting_edges(threads)                # src/main.rs:224
    graph.combine_edges(threads)                             # src/main.rs:226
    graph.align_and_output_pairs(threads)                    # src/main.rs:232 (no DIAMOND)
"""
from __future__ import annotations

import sys

import numpy as np

from .protein import ProteinList


class KmerEdge:
    """KmerEdgeGroup view (src/graph/edge.rs:48-52,119-154)."""

    def __init__(self, graph: "Graph", index: int):
        self._g, self._i = graph, index

    def get_vertices_key(self):
        e = self._g._edges[self._i]
        return [int(e["a"]), int(e["b"])]

    def get_kmers(self) -> np.ndarray:
        """ids of the shared repeated k-mers (canonical ids, ascending)"""
        e = self._g._edges[self._i]
        eng = self._g.protein_list.engine
        values = eng.get_edge_kmers(self._i, int(e["count"]))
        return eng.lookup_kmers(values)

    def get_kmer_values(self) -> np.ndarray:
        e = self._g._edges[self._i]
        return self._g.protein_list.engine.get_edge_kmers(self._i, int(e["count"]))

    def get_proteins_ids_and_sequences(self):
        a, b = self.get_vertices_key()
        pl = self._g.protein_list
        return [pl[a].get_id_and_seq(), pl[b].get_id_and_seq()]

    def __len__(self):
        return int(self._g._edges[self._i]["count"])


class Graph:
    def __init__(self, protein_list: ProteinList, log=sys.stderr):
        self.protein_list = protein_list
        self._log = log
        self._stats = None
        self._edges = None

    @classmethod
    def new(cls, kmer_freq, thread_count: int, protein_list: ProteinList, log=sys.stderr) -> "Graph":
        """Graph::new (src/graph/mod.rs:39-193): prints the two construction counters."""
        if kmer_freq is None or len(kmer_freq) != protein_list.engine.index_stats.get("n_repeated", -1):
            raise ValueError("kmer_freq must be the engine's kmer_freq (build_index first)")
        g = cls(protein_list, log)
        g._score()
        k = protein_list.engine.k
        print(f"Number of {k}mers found in at least two proteins: {len(kmer_freq)}", file=log)
        print(f"Number of total edges: {g._stats['n_multi_edges']}", file=log)
        return g

    def _score(self):
        eng = self.protein_list.engine
        self._stats = eng.score_pairs()
        self._edges = eng.get_edges()

    def remove_uninteresting_edges(self, thread_count: int = 1):
        """src/graph/mod.rs:549-697; the class filter itself is fixed when the engine is
        created (cross_class_only)."""
        print("Remove edges without diverging AMR labels", file=self._log)
        print(f"Number of edges now: {self._stats['n_multi_edges_kept']}", file=self._log)

    def combine_edges(self, thread_count: int = 1):
        """src/graph/mod.rs:322-546"""
        print("Combine edges with the same two vertices", file=self._log)
        print(f"Number of edges now: {self._stats['n_pairs_kept']}", file=self._log)

    @property
    def edges(self):
        """pairs over the threshold (the only edges align_and_output_pairs looks at)"""
        return [KmerEdge(self, i) for i in range(len(self._edges))]

    def all_pair_edges(self) -> np.ndarray:
        """`pub edges` as the reference holds it after combine_edges (src/graph/mod.rs:32): EVERY pair that
        shares at least one k-mer (4 350 628 on the ARG set), not only the pairs over the threshold, as an
        (a, b, count, blosum) array sorted by (a, b).  Scored by a second engine with threshold 0 on the same
        device (the threshold is fixed when an engine is created); the pairs over the threshold are the
        subset `count > threshold`."""
        from .engine import Engine
        eng = self.protein_list.engine
        cross = bool(eng.cross_class_only)
        with Engine(eng.k, device=eng.device, threshold=0, cross_class_only=cross, want_blosum=eng.want_blosum) as e0:
            e0.set_protein_set(self.protein_list.set)
            e0.build_index()
            e0.score_pairs()
            return e0.get_edges()

    def align_and_output_pairs(self, thread_count: int = 1, handoff_dir: str | None = None):
        """src/graph/mod.rs:195-319 without the DIAMOND subprocesses: logs every surviving pair like :250-251
        and returns the edge array; with `handoff_dir`, also writes what the reference leaves for DIAMOND
        (fasta_files/{i}_{accession}.fasta per endpoint, db_files/, blastp_output.tsv with the header line,
        :202-220,253-261,273-280,304-317: kc_write_handoff)."""
        ids = self.protein_list.set.ids
        for e in self._edges:
            print(f"Cross-checking:\n\treference protein:{ids[e['a']]}\n\tquery protein:{ids[e['b']]}"
                  f"\n\tkmers in common:{e['count']}", file=self._log)
        if handoff_dir is not None:
            from .engine import write_handoff
            write_handoff(self.protein_list.set, self._edges, handoff_dir)
        return self._edges

    @property
    def stats(self) -> dict:
        return dict(self._stats)
