"""Multi-GPU host logic: one process per GPU, `torch.distributed` (NCCL over NVLink on the
GPU box, gloo in the CPU tests) for the plumbing.

The pair triangle S = A*A^T is cut into `world` contiguous row blocks of equal estimated work
(`kc_score_pairs_shard`); every rank scores its own block with no data-path collective, then
the sorted per-rank edge lists are gathered to rank 0 and merged.  The index
is either built by every rank (replicated, no communication) or built by rank 0 and broadcast
(`broadcast_index`, device buffers exposed by `kc_index_export`).
"""
from __future__ import annotations

import numpy as np

EDGE_WORDS = 4  # a, b, count, blosum


def shard_bounds(work_prefix: np.ndarray, n_shards: int) -> np.ndarray:
    """Row range of every shard from the exclusive work prefix (length n+1, last = total).
    Mirrors shard_bounds_kernel (csrc/engine.cu): shard s starts at the first row whose prefix
    reaches total*s/n_shards."""
    n = work_prefix.size - 1
    total = int(work_prefix[n])
    out = np.zeros(n_shards + 1, dtype=np.int64)
    for s in range(1, n_shards):
        target = total // n_shards * s + (total % n_shards) * s // n_shards
        out[s] = int(np.searchsorted(work_prefix[:n], target, side="left"))
    out[n_shards] = n
    return out


def merge_edge_lists(parts) -> np.ndarray:
    """Concatenate per-rank edge arrays and restore the canonical (a, b) order."""
    parts = [p for p in parts if p is not None and p.size]
    if not parts:
        from .engine import EDGE_DTYPE
        return np.zeros(0, dtype=EDGE_DTYPE)
    allp = np.concatenate(parts)
    return allp[np.lexsort((allp["b"], allp["a"]))]


def gather_edges(edges: np.ndarray, dist, rank: int, world: int, device=None, dst: int = 0):
    """Variable-length gather of edge lists to `dst`: sizes first (all_gather), then one padded
    gather.  `device` = torch device of the staging tensors (cuda for NCCL, cpu for gloo).
    Returns the merged, sorted edge array on `dst`, None elsewhere."""
    import torch
    from .engine import EDGE_DTYPE
    if world == 1:
        return edges
    device = device or torch.device("cpu")
    cnt = torch.tensor([edges.size], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    buf = torch.zeros(mx * EDGE_WORDS, dtype=torch.int32, device=device)
    if edges.size:
        flat = torch.from_numpy(np.ascontiguousarray(edges).view(np.int32).reshape(-1))
        buf[:flat.numel()] = flat.to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    parts = [o[:c * EDGE_WORDS].cpu().numpy().view(EDGE_DTYPE) for o, c in zip(out, counts)]
    return merge_edge_lists(parts)


def reduce_pair_stats(stats: dict, dist, world: int, device=None) -> dict:
    """Whole-job counters from per-shard counters (n_multi_edges is a whole-set constant)."""
    import torch
    if world == 1:
        return dict(stats)
    device = device or torch.device("cpu")
    keys = ["n_multi_edges_kept", "n_pairs_kept", "n_edges_out", "sum_count_out", "n_rows"]
    t = torch.tensor([stats[k] for k in keys], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    out = dict(stats)
    for k, v in zip(keys, t.tolist()):
        out[k] = int(v)
    return out


def score_sharded(engine, dist, rank: int, world: int, device=None):
    """Rank-local scoring + gather: returns (whole-job stats, merged edges on rank 0)."""
    st = engine.score_pairs(rank, world)
    edges = engine.get_edges()
    return reduce_pair_stats(st, dist, world, device), gather_edges(edges, dist, rank, world, device)
