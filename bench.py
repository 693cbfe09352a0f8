#!/usr/bin/env python
"""bench.py — headline benchmark of the k-mer clustering hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Metric (BASELINE.json): protein pairs scored per second (plus k-mers indexed per second),
on the synthetic 1 M-protein, mean-350-aa, k=7, all-pairs configuration (config 4), which fits
one GPU.  One "step" = one pass of the hot path over the whole protein set:
    build_index (K1-K5: extract, per-protein dedup, census, perfect index, postings)
  + score_pairs (K7-K9: all-pairs shared-k-mer counts, threshold 10, BLOSUM, sorted edges).
`value` is measured with the residue stream already resident in HBM, CUDA events on the
launching stream; `e2e` is the same step through the C ABI from pinned HOST buffers with the
H2D staging and the D2H edge readback inside the timed region (wall clock around a sync; the
upload is chunked and overlaps the first index kernel, so `breakdown_ms.stage_h2d` is only the
host side of the staging call and the rest of the copy shows up in `build_index`).
At N > 1 the pair triangle is cut into 2 N row blocks (rank g owns blocks g and 2 N - 1 - g) and every rank
computes everything for its own rows (kc_build_index_shard / kc_score_pairs_shard: owner computes, no collective
between the kernels).  The e2e leg goes through the library's multi-GPU entry points (NCCL below the C ABI):
every rank uploads 1 / N of the residue stream and the slices are all-gathered over NVLink, the counters are
all-reduced, and every rank copies its own sorted edge runs over its own PCIe link into ONE host buffer shared
by all ranks (POSIX shared memory).  Total work is fixed, so "scaling" is "strong".  At every N the gathered
edge list is checked against the oracle's committed SHA-256 (tests/golden/synth_golden.json).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (n_proteins, length law, k, seed, cross_class_only)
    "synth_1m_k7": (1_000_000, "A", 7, 0xB2000004, False),    # BASELINE.json configs[3]
    "synth_100k_k5": (100_000, "A", 5, 0xB2000003, False),    # configs[2]
    "synth_20k_k5": (20_000, "A", 5, 0xB2000003, False),      # quick check
    "synth_1m_skew_k7": (1_000_000, "B", 7, 0xB2000005, False),  # one GPU's quarter of configs[4] (skewed lengths 50-2000)
    "synth_250k_skew_k7": (250_000, "B", 7, 0xB2000005, False),  # the golden prefix of configs[4]
    "synth_4m_skew_k7": (4_000_000, "B", 7, 0xB2000005, False),  # BASELINE.json configs[4], k = 7 (8 GPUs)
    "synth_4m_skew_k5": (4_000_000, "B", 5, 0xB2000005, False),  # configs[4], k = 5: ~1.6e12 multi-edges
    "synth_250k_skew_k5": (250_000, "B", 5, 0xB2000005, False),  # the k = 5 regime at a size that fits a short run
}
THRESHOLD = 10
# cpu_baseline leg of the GPU arm: a bounded sample (the default run has to finish within minutes).
# The reference arm (--impl reference) runs the FULL workload per step: same config as the GPU arm.
CPU_SAMPLE = {"synth_1m_k7": 250_000, "synth_100k_k5": 50_000, "synth_20k_k5": 20_000, "synth_1m_skew_k7": 200_000,
              "synth_250k_skew_k7": 100_000, "synth_4m_skew_k7": 200_000, "synth_4m_skew_k5": 20_000,
              "synth_250k_skew_k5": 20_000}


def golden_for(workload: str, n: int):
    """the oracle's committed full-size values for this workload (tests/golden/synth_golden.json), or None"""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "synth_golden.json")) as fh:
            g = json.load(fh).get(workload)
        return g if g and g["n"] == n else None
    except Exception:
        return None


def config_of(workload: str, ps, k: int, cross: bool, world: int, nnz: int | None) -> dict:
    """the `config` object of the JSON line: identical for the GPU arm and the reference arm"""
    n = ps.n
    return {"workload": workload, "n_proteins": n, "mean_len": round(ps.residues.size / n, 1),
            "k": k, "threshold": THRESHOLD, "cross_class_only": cross, "blosum": True,
            "generator": "G1 (include/kc_synth.h)", "seed": hex(WORKLOADS[workload][3]),
            "l2_policy": f"inputs larger than L2 ({ps.residues.size / 1e6:.0f} MB residues); no flush needed",
            "parallelism": ("1 GPU" if world == 1 else
                            f"row-block sharded index + pair triangle on {world} GPUs (owner computes; NCCL: residue "
                            "all-gather, counter all-reduce, edge gather)")}


def measured_traffic(workload: str, kernel: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r2.json")) as fh:
            return int(json.load(fh)[workload][kernel])
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_set(workload: str, n_override: int | None):
    import uniprot_kmer_based_clustering_b200 as kc
    n, law, k, seed, cross = WORKLOADS[workload]
    if n_override:
        n = n_override
    threads = min(32, os.cpu_count() or 8)
    return kc.ProteinSet.synthetic(n, law, seed, threads=threads), k, cross


def oracle_run(ps, k, cross, n_sample, threads):
    """CPU restatement (oracle/) on the first n_sample proteins; returns metric dict."""
    from oracle.oracle import Oracle
    n = min(n_sample, ps.n)
    o = Oracle(k, threads)
    o.set_proteins(ps.residues[:int(ps.offsets[n])], ps.offsets[:n + 1], ps.class_id[:n])
    t0 = time.perf_counter()
    o.extract_kmers()
    ix = o.build_index()
    t1 = time.perf_counter()
    pr = o.score_pairs(THRESHOLD, cross, True, mode=1)
    t2 = time.perf_counter()
    pairs = n * (n - 1) // 2
    return {"n": n, "pairs": pairs, "total_s": t2 - t0, "index_s": t1 - t0, "pairs_s": t2 - t1,
            "positions": ix.stats["n_positions"], "multi_edges": pr.stats["n_multi_edges"],
            "edges": pr.stats["n_edges_out"]}


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (the oracle port; the Rust crate needs a nightly
    toolchain + crates.io and cannot be built here) on all host threads, the FULL workload per step: the same
    config as the GPU arm.  Warm-up steps run a small prefix (they only fault the code and the pages in)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ps, k, cross = make_set(args.workload, args.n_proteins)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    for _ in range(args.warmup):
        oracle_run(ps, k, cross, max(2000, ps.n // 50), threads)
    runs = [oracle_run(ps, k, cross, ps.n, threads) for _ in range(args.steps)]
    t = sum(r["total_s"] for r in runs) / len(runs)
    r = runs[-1]
    value = r["pairs"] / t
    sample = (f"all {r['n']} proteins of {args.workload} (k={k}), whole hot path per step; oracle port "
              f"(oracle/kc_oracle.cpp), {threads} threads")
    line = {
        "impl": "reference", "metric": "protein_pairs_scored_per_s", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_of(args.workload, ps, k, cross, world, None),
        "kmers_indexed_per_s": r["positions"] / (sum(x["index_s"] for x in runs) / len(runs)),
        "multi_edges_per_s": r["multi_edges"] / (sum(x["pairs_s"] for x in runs) / len(runs)),
        "counts": {"n_positions": r["positions"], "n_multi_edges": r["multi_edges"], "n_edges_out": r["edges"]},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="synth_1m_k7", choices=sorted(WORKLOADS))
    ap.add_argument("--n-proteins", type=int, default=None, help="override the workload's protein count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--index", default="auto", choices=["auto", "stream", "bucket", "table"],
                    help="kc_config.index_build (auto: streaming build on one GPU, bucket build first on shards)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import uniprot_kmer_based_clustering_b200 as kc
    from uniprot_kmer_based_clustering_b200 import sharded

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(args.warmup, 3)

    ps, k, cross = make_set(args.workload, args.n_proteins)
    n = ps.n
    pairs_total = n * (n - 1) // 2
    # pinned host staging buffers (e2e) and device-resident copies (value)
    h_res = torch.from_numpy(ps.residues).pin_memory()
    h_off = torch.from_numpy(ps.offsets.view(np.int64)).pin_memory()
    h_cls = torch.from_numpy(ps.class_id.view(np.int32)).pin_memory()
    d_res, d_off, d_cls = h_res.cuda(), h_off.cuda(), h_cls.cuda()
    stream = torch.cuda.current_stream()

    eng = kc.Engine(k, device=local_rank, threshold=THRESHOLD, cross_class_only=cross, want_blosum=True,
                    index_build=args.index)
    eng.set_stream(stream.cuda_stream)
    if world > 1:  # the library's own NCCL rank (csrc/dist.cuh); torch.distributed only hands the id around
        box = [kc.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], rank, world)

    def step_resident():
        ist = eng.build_index(rank, world)
        pst = eng.score_pairs(rank, world)
        return ist, pst

    e2e_parts = {"stage_h2d": 0.0, "build_index": 0.0, "score_pairs": 0.0, "edges_d2h": 0.0}

    def step_e2e():
        t0 = time.perf_counter()
        # no sync here: the copies are asynchronous (N = 1: chunked on the engine's copy stream, the first
        # index kernel starts on the first chunk; N > 1: 1 / N uploaded per rank + NCCL all-gather)
        if world == 1:
            eng.set_proteins_ptr(h_res.data_ptr(), h_off.data_ptr(), h_cls.data_ptr(), n, on_device=False)
        else:
            eng.set_proteins_dist_ptr(h_res.data_ptr(), h_off.data_ptr(), h_cls.data_ptr(), n)
        t1 = time.perf_counter()
        ist = eng.build_index_dist() if world > 1 else eng.build_index()
        t2 = time.perf_counter()
        pst = eng.score_pairs_dist() if world > 1 else eng.score_pairs()
        t3 = time.perf_counter()
        if world == 1:
            n_e = pst["n_edges_out"]
            if edge_cap < n_e:
                raise RuntimeError("edge staging buffer too small")
            eng.get_edges_into(edge_ptr, edge_cap)
        else:  # every rank copies its own runs into the buffer all ranks map
            n_e = eng.gather_edges_into(edge_ptr, edge_cap, shared=True)
        t4 = time.perf_counter()
        for key, dt in zip(e2e_parts, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            e2e_parts[key] += dt * 1e3
        return ist, pst, n_e

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, CUDA events on the launching stream -------------------
    eng.set_proteins_ptr(d_res.data_ptr(), d_off.data_ptr(), d_cls.data_ptr(), n, on_device=True)
    for _ in range(W):
        ist, pst = step_resident()
    eng.reset_timings()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {"index_ms": 0.0, "pairs_ms": 0.0, "edges_ms": 0.0, "pair_kernel_ms": 0.0, "census_kernel_ms": 0.0}
    ev0.record(stream)
    for _ in range(args.steps):
        ist, pst = step_resident()
        t = eng.timings()
        for key in stage_ms:
            stage_ms[key] += t[key]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    launches = eng.timings()["kernel_launches"]
    for key in stage_ms:
        stage_ms[key] /= args.steps

    # ---- e2e: host buffers, H2D + D2H (+ NCCL) inside the timed region, wall clock -------------
    n_e_all = torch.tensor([pst["n_edges_out"]], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(n_e_all)
    edge_cap = max(int(n_e_all.item() * 1.25) + 1024, 1 << 16)
    shm = None
    if world == 1:
        h_edges = torch.empty(edge_cap * 4, dtype=torch.int32).pin_memory()
        edge_ptr, edges_np = h_edges.data_ptr(), h_edges.numpy()
    else:
        # ONE host buffer for the gathered list, mapped by every rank (POSIX shared memory) and page-locked
        # in every process: each rank's D2H copy goes over its own PCIe link
        from multiprocessing import shared_memory
        name = f"kc_b200_edges_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
        if rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except FileNotFoundError:
                pass
            shm = shared_memory.SharedMemory(name=name, create=True, size=edge_cap * 16)
        dist.barrier()
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name)
        edges_np = np.frombuffer(shm.buf, dtype=np.int32, count=edge_cap * 4)
        edge_ptr = edges_np.ctypes.data
        rc = torch.cuda.cudart().cudaHostRegister(edge_ptr, edge_cap * 16, 0)
        if int(rc) != 0 and rank == 0:
            print(f"cudaHostRegister of the shared edge buffer failed ({rc}): pageable copies", file=sys.stderr)
    for _ in range(2):
        step_e2e()
    for key in e2e_parts:
        e2e_parts[key] = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ist_e, pst_e, n_e_total = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps

    # max over ranks; sums over ranks for the sharded quantities
    red = torch.tensor([dev_ms, e2e_s, stage_ms["index_ms"], stage_ms["pairs_ms"], stage_ms["pair_kernel_ms"],
                        stage_ms["census_kernel_ms"], stage_ms["edges_ms"]] + [e2e_parts[k2] / args.steps for k2 in e2e_parts],
                       dtype=torch.float64, device="cuda")
    sums = torch.tensor([pst["n_multi_edges_kept"], pst["n_edges_out"], pst["n_pairs_kept"], launches],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    red = red.tolist()
    dev_ms, e2e_s, index_ms, pairs_ms, pair_kernel_ms, census_ms, edges_ms = red[:7]
    e2e_break = dict(zip(e2e_parts, red[7:]))
    m_kept, e_out, p_kept, launches_all = (int(x) for x in sums.tolist())
    # a sharded index build reports per-rank shares that add up; a whole index per rank (the table build)
    # reports the whole-set numbers on every rank
    sharded_index = eng.index_shard_info()["n_shards"] > 1
    ist_rank = dict(ist)
    ist = sharded.reduce_index_stats(ist, dist, world, torch.device("cuda"), sharded_index)
    pst_all = sharded.reduce_pair_stats(pst, dist, world, torch.device("cuda"), sharded_index)

    if rank == 0:
        peak, peak_src = measured_peak()
        nnz = ist["nnz"]
        # 4 B per multi-edge + 4 B per CSR nonzero + 16 B per emitted edge; at N > 1 every
        # rank streams its own rows' suffixes, so the per-rank figure uses the rank-0 share
        nnz_rank = ist_rank["nnz"] if sharded_index else nnz // world
        algo_bytes = 4 * pst["n_multi_edges_kept"] + 4 * nnz_rank + 16 * pst["n_edges_out"]
        achieved = algo_bytes / (stage_ms["pair_kernel_ms"] * 1e-3) / 1e9 if stage_ms["pair_kernel_ms"] > 0 else 0.0
        idx_bytes = 9 * (ist_rank["n_positions"] if sharded_index else ist["n_positions"])
        idx_achieved = idx_bytes / (stage_ms["index_ms"] * 1e-3) / 1e9 if stage_ms["index_ms"] > 0 else 0.0
        # whole job: the residue stream crosses PCIe once (1 / N per rank), every rank uploads the offsets and
        # (cross-class mode) the row layout; every edge crosses PCIe once
        h2d = int(h_res.numel() + world * (h_off.numel() * 8 + (16 * n if cross else 0)))
        d2h = int(n_e_total * 16 + world * 512)
        flavour = eng.index_flavour()
        line = {
            "metric": "protein_pairs_scored_per_s", "value": pairs_total / (dev_ms * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": dev_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": config_of(args.workload, ps, k, cross, world, nnz),
            "kmers_indexed_per_s": ist["n_positions"] / (index_ms * 1e-3),
            "pair_stage_pairs_per_s": pairs_total / (pairs_ms * 1e-3) if pairs_ms > 0 else None,
            "multi_edges_per_s": m_kept / (pair_kernel_ms * 1e-3) if pair_kernel_ms > 0 else None,
            "stage_ms": {"index": index_ms, "partition_kernels": census_ms, "pairs": pairs_ms,
                         "pair_kernels": pair_kernel_ms, "edges_sort_blosum": edges_ms},
            "counts": {"n_positions": ist["n_positions"], "n_repeated": ist["n_repeated"], "nnz": nnz,
                       "n_multi_edges": pst_all["n_multi_edges"], "n_pairs_nonzero": p_kept, "n_edges_out": e_out},
            "roofline": {"bound": "hbm",
                         "kernel": "pairs_tile_kernel + pairs_main_scored_kernel + packed/dense kernels (K7-K9)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (measured_traffic(args.workload, "pairs_main_scored_kernel") +
                                     measured_traffic(args.workload, "pairs_tile_kernel"))
                         if world == 1 and not args.n_proteins and
                         measured_traffic(args.workload, "pairs_main_scored_kernel") is not None and
                         measured_traffic(args.workload, "pairs_tile_kernel") is not None else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes": algo_bytes},
            "roofline_index": {"bound": "hbm",
                               "kernel": "K1-K5 (sx_l1/l2 partition kernels, sx_warp_bucket_kernel, rows_finalize_kernel)"
                               if flavour == 1 else "K1-K5 (extract, census, ids, postings, suffix ranges)" if flavour == 0
                               else "K1-K5 (extract_scatter, bucket_build, rows_finalize kernels: the sharded build)",
                               "achieved": idx_achieved, "peak": peak, "unit": "GB/s", "frac": idx_achieved / peak,
                               "algorithmic_bytes": idx_bytes},
            "e2e": {"value": pairs_total / e2e_s, "unit": "pairs/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "breakdown_ms": e2e_break},
            "gpu_launches": launches_all,
            "clocks": clocks,
        }
        # the gathered list against the oracle: complete, sorted, and (full-size workloads) bit-identical to the
        # committed golden SHA-256 at EVERY GPU count
        assert n_e_total == e_out, (n_e_total, e_out)
        edges = edges_np[:n_e_total * 4].view(kc.EDGE_DTYPE)
        ekey = (edges["a"].astype(np.uint64) << np.uint64(32)) | edges["b"].astype(np.uint64)
        assert bool(np.all(ekey[1:] > ekey[:-1])), "gathered edge list is not sorted by (a, b)"
        gold = golden_for(args.workload, n)
        if gold is not None:
            import hashlib
            sha = hashlib.sha256(np.ascontiguousarray(edges).tobytes()).hexdigest()
            assert sha == gold["edges_sha256"], "edge list differs from the oracle's golden SHA-256"
            for key, val in gold["index"].items():
                assert ist_e[key] == val, (key, ist_e[key], val)
            for key in ("n_multi_edges", "n_multi_edges_kept", "n_pairs_kept", "n_edges_out", "sum_count_out"):
                assert pst_e[key] == gold["pairs"][key], (key, pst_e[key], gold["pairs"][key])
            line["parity"] = {"edges_sha256": sha, "golden": "tests/golden/synth_golden.json", "checked": True}
        else:
            line["parity"] = {"checked": False, "why": "no golden for this workload / size"}
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            r = oracle_run(ps, k, cross, CPU_SAMPLE[args.workload], threads)
            line["cpu_baseline"] = {
                "value": r["pairs"] / r["total_s"], "unit": "pairs/s", "cores": threads, "kind": "port",
                "sample": f"first {r['n']} of {n} proteins, whole hot path once ({r['total_s']:.1f} s); "
                          f"oracle port (CPU restatement), not the Rust binary",
                "kmers_indexed_per_s": r["positions"] / r["index_s"],
                "multi_edges_per_s": r["multi_edges"] / r["pairs_s"] if r["pairs_s"] > 0 else None}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.barrier()
        try:
            torch.cuda.cudart().cudaHostUnregister(edge_ptr)
        except Exception:
            pass
        del edges_np
        if rank == 0:
            try:
                del edges
            except NameError:
                pass
        try:
            shm.close()
        except BufferError:
            pass
        if rank == 0:
            try:
                shm.unlink()
            except FileNotFoundError:
                pass
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
