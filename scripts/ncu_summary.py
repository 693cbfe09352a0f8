#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv            > profiles/x.md
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep [kernel-regex] > profiles/y.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    tot = collections.OrderedDict()
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "usecond": v / 1e3, "nsecond": v / 1e6}.get(r["Metric Unit"], v)
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))
        tot.setdefault(name, [0, 0.0])
        tot[name][0] += 1
        tot[name][1] += v
    s = sum(v for _, v in tot.values())
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {v:.3f} | {100 * v / s:.1f}% |")
    print(f"| **all** | {sum(c for c, _ in tot.values())} | {s:.3f} | 100% |")


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
    ("smsp__inst_executed.sum", "warp insts"),
]


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in WANT if m in idx]
    print("| kernel | " + " | ".join(n for _, n in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for r in rows:
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[idx["Kernel Name"]]))
        if pattern and not re.search(pattern, name):
            continue
        vals = []
        for m, _ in cols:
            v = r[idx[m]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3f}".rstrip("0").rstrip(".") if abs(f) < 1e6 else f"{f:.3e}"
            except ValueError:
                pass
            vals.append(f"{v} {units[idx[m]]}".strip())
        print(f"| `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
