"""World-size-2 gloo test (CPU) of the multi-GPU host logic: shard bounds, per-rank scoring,
variable-length edge gather to rank 0, stats reduction.  The oracle stands in for the GPU
engine (same interface for a row range); the plumbing under test is the product's sharded.py."""
import os
import socket

import numpy as np
import pytest

from conftest import random_protein_set


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _row_work(ix, n):
    """numpy mirror of suffix_ranges_kernel + WorkIn: multi-edges accumulated per row."""
    rowlen = np.diff(ix.row_offsets.astype(np.int64))
    rows = np.repeat(np.arange(n), rowlen)
    order = np.lexsort((rows, ix.ids))
    sid = ix.ids[order]
    start = np.searchsorted(sid, sid, side="left")
    pos = np.arange(sid.size) - start
    f = ix.freq[sid].astype(np.int64)
    work = np.zeros(n, dtype=np.int64)
    np.add.at(work, rows[order], f - pos - 1)
    weight = work + 2 * rowlen + np.where(work > 0, 64, 0)
    return work, np.concatenate([[0], np.cumsum(weight)])


def _worker(rank, world, port, seed, q):
    import torch.distributed as dist
    from oracle.oracle import Oracle
    from uniprot_kmer_based_clustering_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ps = random_protein_set(seed, 400, min_len=20, max_len=120, n_classes=4, family=8, mutate=0.03,
                                alphabet="ACDEFGHIKLMNPQRSTVWY")
        o = Oracle(5, 1)
        o.set_proteins(ps.residues, ps.offsets, ps.class_id)
        o.extract_kmers()
        ix = o.build_index()
        work, prefix = _row_work(ix, ps.n)
        bounds = sharded.shard_bounds(prefix, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        mine = o.score_pairs(3, False, True, mode=1, row_lo=lo, row_hi=hi) if hi > lo else None
        stats = dict(mine.stats) if mine else {"n_multi_edges": 0, "n_multi_edges_kept": 0, "n_pairs_kept": 0,
                                               "n_edges_out": 0, "sum_count_out": 0}
        stats["n_rows"] = hi - lo
        edges = mine.edges if mine else np.zeros(0, dtype=sharded_dtype())
        total = sharded.reduce_pair_stats(stats, dist, world)
        merged = sharded.gather_edges(edges, dist, rank, world)
        if rank == 0:
            full = o.score_pairs(3, False, True, mode=1)
            ok = (np.array_equal(merged, full.edges)
                  and total["n_pairs_kept"] == full.stats["n_pairs_kept"]
                  and total["n_multi_edges_kept"] == full.stats["n_multi_edges_kept"] == int(work.sum())
                  and total["n_edges_out"] == full.edges.size and total["n_rows"] == ps.n
                  and full.edges.size > 50)
            balance = [int(work[int(bounds[s]):int(bounds[s + 1])].sum()) for s in range(world)]
            q.put((ok, balance, int(full.edges.size)))
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def sharded_dtype():
    from uniprot_kmer_based_clustering_b200 import EDGE_DTYPE
    return EDGE_DTYPE


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scoring_and_gather_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 11, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, balance, n_edges = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok, (balance, n_edges)
    assert max(balance) <= 2.0 * (sum(balance) / world) + 1000, balance


def test_shard_bounds_cover_all_rows():
    from uniprot_kmer_based_clustering_b200 import sharded
    rng = np.random.default_rng(0)
    w = rng.integers(0, 1000, size=1000)
    prefix = np.concatenate([[0], np.cumsum(w)])
    for s in (1, 2, 3, 8, 64):
        b = sharded.shard_bounds(prefix, s)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0)
        loads = [int(w[b[i]:b[i + 1]].sum()) for i in range(s)]
        assert max(loads) <= sum(loads) / s + 1000
    assert sharded.shard_bounds(np.zeros(1, dtype=np.int64), 4).tolist() == [0, 0, 0, 0, 0]
    assert sharded.merge_edge_lists([None, np.zeros(0, dtype=sharded_dtype())]).size == 0


def _worker_owner_computes(rank, world, port, seed, q):
    """The sharded-index scheme on the host side: zig-zag row blocks, per-rank scoring of the owned
    blocks, index / pair counters owned by the rank of a k-mer's first holder, summed over gloo."""
    import torch.distributed as dist
    from oracle.oracle import Oracle
    from uniprot_kmer_based_clustering_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ps = random_protein_set(seed, 700, min_len=20, max_len=120, n_classes=4, family=8, mutate=0.03,
                                alphabet="ACDEFGHIKLMNPQRSTVWY")
        k = 5
        o = Oracle(k, 1)
        o.set_proteins(ps.residues, ps.offsets, ps.class_id)
        o.extract_kmers()
        ix = o.build_index()
        lens = np.diff(ps.offsets.astype(np.int64))
        bounds, owner = sharded.zigzag_blocks(np.maximum(lens - k + 1, 0), world)
        row_owner = np.zeros(ps.n, dtype=np.int64)
        for b in range(owner.size):
            row_owner[bounds[b]:bounds[b + 1]] = owner[b]
        # index counters: a k-mer belongs to the rank of its first holder
        rowlen = np.diff(ix.row_offsets.astype(np.int64))
        rows = np.repeat(np.arange(ps.n), rowlen)
        first = np.full(ix.vocab.size, ps.n, dtype=np.int64)
        np.minimum.at(first, ix.ids, rows)
        mine_k = row_owner[first] == rank
        f = ix.freq.astype(np.int64)
        ist = {"n_positions": int(np.maximum(lens - k + 1, 0)[row_owner == rank].sum()), "n_incidences": 0,
               "n_distinct": 0, "n_singleton": 0, "n_repeated": int(mine_k.sum()), "nnz": int(f[mine_k].sum())}
        pst = {"n_multi_edges": int((f[mine_k] * (f[mine_k] - 1) // 2).sum()), "n_multi_edges_kept": 0,
               "n_pairs_kept": 0, "n_edges_out": 0, "sum_count_out": 0, "n_rows": int((row_owner == rank).sum())}
        parts = []
        for b in range(owner.size):
            if owner[b] != rank or bounds[b + 1] <= bounds[b]:
                continue
            r = o.score_pairs(3, False, True, mode=1, row_lo=int(bounds[b]), row_hi=int(bounds[b + 1]))
            for name in ("n_multi_edges_kept", "n_pairs_kept", "n_edges_out", "sum_count_out"):
                pst[name] += r.stats[name]
            parts.append(r.edges)
        tot_i = sharded.reduce_index_stats(ist, dist, world, sharded_index=True)
        tot_p = sharded.reduce_pair_stats(pst, dist, world, sharded_index=True)
        merged = sharded.gather_edges(sharded.merge_edge_lists(parts), dist, rank, world)
        if rank == 0:
            full = o.score_pairs(3, False, True, mode=1)
            ok = (np.array_equal(merged, full.edges)
                  and tot_i["n_repeated"] == ix.stats["n_repeated"] and tot_i["nnz"] == ix.stats["nnz"]
                  and tot_i["n_positions"] == ix.stats["n_positions"]
                  and tot_p["n_multi_edges"] == full.stats["n_multi_edges"]
                  and tot_p["n_pairs_kept"] == full.stats["n_pairs_kept"] and tot_p["n_rows"] == ps.n
                  and bounds[0] == 0 and bounds[-1] == ps.n and np.all(bounds[1:-1] % 64 == 0))
            q.put((ok, bounds.tolist(), int(full.edges.size)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_owner_computes_sharding_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_owner_computes, args=(r, world, port, 13, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, bounds, n_edges = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok, (bounds, n_edges)


def test_zigzag_blocks_pair_early_with_late():
    from uniprot_kmer_based_clustering_b200 import sharded
    pos = np.full(64 * 40, 100)
    bounds, owner = sharded.zigzag_blocks(pos, 4)
    assert owner.tolist() == [0, 1, 2, 3, 3, 2, 1, 0]
    assert bounds[0] == 0 and bounds[-1] == pos.size and np.all(np.diff(bounds) == 64 * 5)
    b1, o1 = sharded.zigzag_blocks(pos, 1)
    assert b1.tolist() == [0, pos.size] and o1.tolist() == [0]
