"""Multi-GPU host-side helpers in Python: the mirror of the engine's row-block cut (`zigzag_blocks`,
`shard_bounds`), edge-list merging and `torch.distributed` reductions of the per-rank counters.

The product path for several GPUs is BELOW the C ABI (csrc/dist.cuh: `kc_comm_init`, `kc_set_proteins_dist`,
`kc_build_index_dist`, `kc_score_pairs_dist`, `kc_gather_edges[_shared]`; `Engine.comm_init` etc.), NCCL included;
`bench.py` and the CLI use that.  What is left here serves the CPU tests (gloo, world size 2-3, a CPU checker standing
in for the engine: `tests/test_sharded_gloo.py`), callers that already run `torch.distributed` and want the counters
reduced there (`reduce_index_stats`, `reduce_pair_stats`), and `gather_edges*`, the torch-level edge gathers of
round 1 (kept for comparison; the library's gather replaced them in `bench.py`).
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


BIN_ROWS = 64  # row blocks are cut at 64-row borders (kBinRows, csrc/common.cuh)


def zigzag_blocks(positions_per_row: np.ndarray, n_shards: int):
    """Row blocks of a sharded index build (mirror of build_index_bucketed, csrc/engine.cu): the pair
    order is cut into 2 * n_shards blocks of equal k-mer positions at 64-row borders; block b belongs
    to rank b if b < n_shards, else to rank 2 * n_shards - 1 - b.  Returns (bounds[n_blocks + 1],
    owner[n_blocks])."""
    n = int(positions_per_row.size)
    if n_shards <= 1:
        return np.array([0, n], dtype=np.int64), np.array([0], dtype=np.int64)
    n_blocks = 2 * n_shards
    prefix = np.concatenate([[0], np.cumsum(positions_per_row.astype(np.int64))])
    total = int(prefix[n])
    bounds = np.full(n_blocks + 1, n, dtype=np.int64)
    bounds[0] = 0
    for b in range(1, n_blocks):
        target = total // n_blocks * b + (total % n_blocks) * b // n_blocks
        r = int(np.searchsorted(prefix, target, side="left"))
        bounds[b] = max(bounds[b - 1], min(n, (r + BIN_ROWS - 1) // BIN_ROWS * BIN_ROWS))
    owner = np.array([b if b < n_shards else n_blocks - 1 - b for b in range(n_blocks)], dtype=np.int64)
    return bounds, owner


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


class _DeviceWords:
    """CUDA array interface over a raw device pointer (int32 words), so torch can wrap engine memory"""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def gather_edges_device(engine, dist, rank: int, world: int, pinned_out=None, dst: int = 0,
                        rows_in_input_order: bool = True):
    """NCCL path: the per-rank sorted edge lists go GPU-to-GPU (send/recv of the engine's device
    buffers over NVLink) into one device buffer on `dst`, then one D2H copy.  With the input-order
    pair order (all-classes mode) every rank's list is one run per row block it owns, so laying the
    runs out in block order gives the list sorted by (a, b) with no merge: one block per rank for a
    whole index (contiguous shards), two per rank (zig-zag) for a sharded index build.  Otherwise
    the merged list is sorted on the host."""
    import torch
    from .engine import EDGE_DTYPE
    ptr, n_e = engine.edges_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    mine = (torch.as_tensor(_DeviceWords(ptr, n_e * EDGE_WORDS), device=dev) if n_e
            else torch.zeros(0, dtype=torch.int32, device=dev))
    if world == 1:
        runs = [[n_e, 0]]
    else:
        info = engine.index_shard_info()
        n_low = n_e
        if info["n_shards"] > 1 and n_e:  # edges of the rank's early block come first (sorted by a)
            hi_start = int(info["block_bounds"][info["n_blocks"] - 1 - rank])
            a_col = mine.view(-1, EDGE_WORDS)[:, 0].contiguous()
            n_low = int(torch.searchsorted(a_col, torch.tensor([hi_start], dtype=torch.int32, device=dev)).item())
        cnt = torch.tensor([n_low, n_e - n_low], dtype=torch.int64, device=dev)
        allc = torch.zeros(2 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, cnt)
        runs = allc.view(world, 2).tolist()
    total = sum(int(a + b) for a, b in runs)
    if rank != dst:
        n_low = int(runs[rank][0])
        if n_low:
            dist.send(mine[:n_low * EDGE_WORDS], dst=dst)
        if n_e - n_low:
            dist.send(mine[n_low * EDGE_WORDS:], dst=dst)
        return None
    # block order: low runs of ranks 0 .. world-1, then high runs of ranks world-1 .. 0
    order = [(r, 0) for r in range(world)] + [(r, 1) for r in reversed(range(world))]
    place = {}
    off = 0
    for r, part in order:
        place[(r, part)] = off
        off += int(runs[r][part])
    buf = torch.empty(total * EDGE_WORDS, dtype=torch.int32, device=dev)
    reqs = []
    for src in range(world):
        for part in (0, 1):  # the sender posts its low run first
            c = int(runs[src][part])
            if not c:
                continue
            o = place[(src, part)]
            seg = buf[o * EDGE_WORDS:(o + c) * EDGE_WORDS]
            if src == dst:
                lo = 0 if part == 0 else int(runs[src][0])
                seg.copy_(mine[lo * EDGE_WORDS:(lo + c) * EDGE_WORDS])
            else:
                reqs.append(dist.irecv(seg, src=src))
    for q in reqs:
        q.wait()
    if pinned_out is not None and pinned_out.numel() >= buf.numel():
        host = pinned_out[:buf.numel()]
        host.copy_(buf, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    else:
        host = buf.cpu()
    edges = host.numpy().view(EDGE_DTYPE)
    if not rows_in_input_order and world > 1:
        edges = edges[np.lexsort((edges["b"], edges["a"]))]
    return edges


INDEX_STAT_KEYS = ["n_positions", "n_incidences", "n_distinct", "n_singleton", "n_repeated", "nnz"]


def reduce_index_stats(stats: dict, dist, world: int, device=None, sharded_index: bool = True) -> dict:
    """Whole-set index counters.  A sharded build (kc_build_index_shard) reports totals over the
    k-mers each rank owns, which add up; a whole index per rank already holds the whole-set numbers."""
    import torch
    if world == 1 or not sharded_index:
        return dict(stats)
    device = device or torch.device("cpu")
    t = torch.tensor([stats[k] for k in INDEX_STAT_KEYS], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    out = dict(stats)
    for k, v in zip(INDEX_STAT_KEYS, t.tolist()):
        out[k] = int(v)
    return out


def stage_residues_allgather(engine, dist, rank: int, world: int, h_res, h_off, h_cls, n: int, d_full):
    """Multi-GPU staging: every rank needs the whole residue stream, but only its 1/world slice crosses
    its PCIe link; the slices are all-gathered in place over NVLink (NCCL) into `d_full` (a uint8 CUDA
    tensor of world * ceil(R / world) bytes, kept by the caller), then handed to the engine together
    with the host offsets / classes (pinned torch tensors).
    EXPERIMENTAL: verified on 2 GPUs only (slower there than the engine's chunked whole-stream upload,
    which overlaps the first index kernel); bench.py does not use it."""
    R = h_res.numel()
    per = d_full.numel() // world
    lo = rank * per
    hi = min(R, lo + per)
    if hi > lo:
        d_full[lo:hi].copy_(h_res[lo:hi], non_blocking=True)
    dist.all_gather_into_tensor(d_full, d_full[lo:lo + per])
    engine.set_proteins_device_residues(d_full.data_ptr(), h_off.data_ptr(), h_cls.data_ptr(), n)


def reduce_pair_stats(stats: dict, dist, world: int, device=None, sharded_index: bool = False) -> dict:
    """Whole-job counters from per-shard counters (n_multi_edges is a whole-set constant when every
    rank holds the whole index, a per-rank share when the index is sharded)."""
    import torch
    if world == 1:
        return dict(stats)
    device = device or torch.device("cpu")
    keys = ["n_multi_edges_kept", "n_pairs_kept", "n_edges_out", "sum_count_out", "n_rows"]
    if sharded_index:
        keys = keys + ["n_multi_edges"]
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
