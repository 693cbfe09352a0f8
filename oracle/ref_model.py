"""Literal, slow Python model of the reference's data structures.  TEST INFRASTRUCTURE ONLY.

Where oracle/kc_oracle.cpp restates WHAT the reference computes, this file follows HOW it
computes it, structure by structure, so that the two restatements are independent:

* `Protein` .............. src/protein.rs:59-64,107-174 (k generalised: 5 or 7)
* `census` ............... src/main.rs:23-48,84-137 (sorted insert list, split unique/repeat)
* MPHF ................... boomphf ids are an arbitrary bijection; here a seeded random
                           permutation plays that role, which exercises SURVEY C7 (no
                           set-level result may depend on the ids)
* `Graph.new` ............ src/graph/mod.rs:39-193 with the triangular edge layout of
                           ProteinVertex::update_graph_edges, src/graph/vertex.rs:59-140, and
                           KmerEdgeSingle::add_vertex, src/graph/edge.rs:33-45
* `remove_uninteresting_edges` src/graph/mod.rs:549-697, keep_specified_edges vertex.rs:40-56
* `combine_edges` ........ src/graph/mod.rs:322-546, KmerEdgeGroup::new edge.rs:56-85
* `pairs_over_threshold` . src/graph/mod.rs:229-251 (`len() <= 10 -> continue`)

Only meant for inputs of a few hundred proteins.
"""
from __future__ import annotations

import random

ALPHABET = "CSTAGPDEQNHRKMILVWYF*"


def amino_acid_to_bits(ch: str) -> int:
    i = ALPHABET.find(ch)
    return i if i >= 0 else 20


def create_kmer(codes) -> int:
    v = 0
    n = len(codes)
    for i, c in enumerate(codes):
        v += c * 21 ** (n - 1 - i)
    return v


class Protein:
    def __init__(self, pid: str, seq: str, k: int = 5):
        self.id, self.seq, self.k = pid, seq, k
        self.kmers = [create_kmer([amino_acid_to_bits(c) for c in seq[s:s + k]])
                      for s in range(max(0, len(seq) - k + 1))]
        self.hash_kmers: list[int] = []

    def get_amr_class(self) -> str:
        f = self.id.split("|")
        if f and f[-1] == "":
            f.pop()
        return f[3]

    def remove_unique(self, phf, unique_table):
        index = 0
        while index < len(self.kmers):
            while unique_table[phf[self.kmers[index]]]:
                del self.kmers[index]
                if index == len(self.kmers):
                    break
            index += 1

    def modify_hash(self, length, phf):
        seen = [False] * length
        for km in self.kmers:
            h = phf[km]
            if not seen[h]:
                seen[h] = True
                self.hash_kmers.append(h)


def _sorted_insert(lst, start, end, item):
    if end - start <= 1:
        if item < lst[start][0]:
            lst.insert(start, [item, 1])
        elif item > lst[start][0]:
            lst.insert(start + 1, [item, 1])
        else:
            lst[start][1] += 1
    else:
        mid = (end + start) // 2
        if item > lst[mid][0]:
            _sorted_insert(lst, mid, end, item)
        elif item < lst[mid][0]:
            _sorted_insert(lst, start, mid, item)
        else:
            lst[mid][1] += 1


def census(proteins):
    freq_list = []
    for p in proteins:
        km = sorted(set(p.kmers))
        if not freq_list:
            freq_list = [[x, 1] for x in km]
        else:
            for item in km:
                _sorted_insert(freq_list, 0, len(freq_list), item)
    return freq_list


class Graph:
    def __init__(self, kmer_freq, proteins, order=None):
        self.proteins = proteins
        nk = len(kmer_freq)
        per = [f * (f - 1) // 2 for f in kmer_freq]
        prefix, s = [], 0
        for x in per:
            s += x
            prefix.append(s)
        self.n_kmers = nk
        self.n_total_edges = prefix[-1] if prefix else 0
        # phase A: edges[e].kmer = h with prefix[h-1] <= e < prefix[h]
        self.edges = []
        h = 0
        for e in range(self.n_total_edges):
            while prefix[h] <= e:
                h += 1
            self.edges.append({"kmers": [h], "v": [0, 0], "visited": 0})
        self.keys = [[e] for e in range(self.n_total_edges)]   # Arc<AtomicUsize>: boxed ints
        self.vertex_edges = [[] for _ in proteins]
        visited = [0] * nk
        order = list(range(len(proteins))) if order is None else order
        # phase B: arrival order = `order` (index order when threads == 1)
        for key in order:
            for km in sorted(set(proteins[key].hash_kmers)):
                left = 0 if km == 0 else prefix[km - 1]
                v = visited[km]
                visited[km] += 1
                f = kmer_freq[km]
                for batch in range(v + 1):
                    offset = sum(f - 1 - x for x in range(batch))
                    if batch == v:
                        targets = [left + offset + j for j in range(f - 1 - batch)]
                    else:
                        targets = [left + offset + (v - 1 - batch)]
                    for ei in targets:
                        edge = self.edges[ei]
                        assert km in edge["kmers"], "Math error yet again"
                        assert edge["visited"] < 2, "I did my math wrong again"
                        edge["v"][edge["visited"]] = key
                        edge["visited"] += 1
                        self.vertex_edges[key].append(self.keys[ei])

    def _keep(self, kept_keys):
        for vi in range(len(self.vertex_edges)):
            bits = [False] * len(self.edges)
            for kk in self.vertex_edges[vi]:
                bits[kk[0]] = True
            self.vertex_edges[vi] = [kk for kk in kept_keys if bits[kk[0]]]

    def _renumber(self, kept_edges, kept_keys):
        self._keep(kept_keys)
        self.edges, self.keys = kept_edges, kept_keys
        for i, kk in enumerate(self.keys):
            kk[0] = i

    def remove_uninteresting_edges(self):
        ke, kk = [], []
        for i, e in enumerate(self.edges):
            if self.proteins[e["v"][0]].get_amr_class() != self.proteins[e["v"][1]].get_amr_class():
                ke.append(e)
                kk.append(self.keys[i])
        self._renumber(ke, kk)
        return len(self.edges)

    def combine_edges(self):
        ke, kk = [], []
        for edge_key in range(len(self.edges)):
            e = self.edges[edge_key]
            first, second = self.vertex_edges[e["v"][0]], self.vertex_edges[e["v"][1]]
            if len(first) == 1 or len(second) == 1:
                ke.append(e)
                kk.append(self.keys[edge_key])
                continue
            bits = [False] * len(self.edges)
            for x in first:
                bits[x[0]] = True
            skip, merge = False, []
            for x in second:
                if bits[x[0]]:
                    if x[0] < edge_key:
                        skip = True
                        break
                    merge.append(x)
            if skip:
                continue
            merge.sort(key=lambda x: x[0])
            if len(merge) > 1:
                kmers = []
                for x in merge:
                    kmers += self.edges[x[0]]["kmers"]
                e = {"kmers": kmers, "v": list(self.edges[merge[0][0]]["v"]), "visited": 2}
                self.edges[edge_key] = e
            ke.append(e)
            kk.append(self.keys[edge_key])
        self._renumber(ke, kk)
        return len(self.edges)

    def pairs_over_threshold(self, threshold=10):
        return [(e["v"][0], e["v"][1], list(e["kmers"])) for e in self.edges
                if len(e["kmers"]) > threshold]


def run_reference_model(records, k=5, threshold=10, seed=0, shuffle_arrival=False):
    """records: list of (id, seq).  Returns the reference's counters and the surviving
    pairs as sorted (a, b, count, sorted shared k-mer VALUES) with a < b in input order."""
    rng = random.Random(seed)
    proteins = [Protein(i, s, k) for i, s in records]
    freq_list = census(proteins)
    all_k = [x for x, _ in freq_list]
    repeat = [x for x, c in freq_list if c != 1]
    perm_all = list(range(len(all_k)))
    rng.shuffle(perm_all)
    perm_rep = list(range(len(repeat)))
    rng.shuffle(perm_rep)
    phf_all = {x: perm_all[i] for i, x in enumerate(all_k)}
    phf_rep = {x: perm_rep[i] for i, x in enumerate(repeat)}
    id_to_kmer = {v: x for x, v in phf_rep.items()}
    unique_table = [False] * len(all_k)
    for x, c in freq_list:
        if c == 1:
            unique_table[phf_all[x]] = True
    kmer_freq = [0] * len(repeat)
    for p in proteins:
        p.remove_unique(phf_all, unique_table)
        p.modify_hash(len(repeat), phf_rep)
        for km in sorted(set(p.kmers)):
            kmer_freq[phf_rep[km]] += 1
    out = {"n_repeated": len(repeat), "n_distinct": len(all_k),
           "repeated": repeat, "freq_by_kmer": {id_to_kmer[i]: f for i, f in enumerate(kmer_freq)},
           "hash_sets": [sorted(id_to_kmer[h] for h in p.hash_kmers) for p in proteins]}
    if not repeat:
        out.update(n_total_edges=0, n_after_class=0, n_after_combine=0, pairs=[])
        return out
    order = list(range(len(proteins)))
    if shuffle_arrival:
        rng.shuffle(order)
    g = Graph(kmer_freq, proteins, order)
    out["n_total_edges"] = g.n_total_edges
    out["n_after_class"] = g.remove_uninteresting_edges()
    out["n_after_combine"] = g.combine_edges()
    pairs = []
    for a, b, kms in g.pairs_over_threshold(threshold):
        pairs.append((min(a, b), max(a, b), len(kms), sorted(id_to_kmer[h] for h in kms)))
    out["pairs"] = sorted(pairs)
    return out
