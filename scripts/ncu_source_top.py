#!/usr/bin/env python
"""Top stall sites of one kernel from an .ncu-rep source page (needs -lineinfo + --import-source on).
    python scripts/ncu_source_top.py gpurun_out/prof.ncu-rep 'pairs_packed_kernel<(int)11' [N]
"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:
            cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]:
        continue
    h = {n: i for i, n in enumerate(b["hdr"])}
    si, ii, ti = h["# Samples"], h["Instructions Executed"], h["Avg. Threads Executed"]
    stall_cols = [(n, i) for n, i in h.items() if n.startswith("stall_") or n.lower().startswith("warp stall")]
    tot = sum(int(r[si] or 0) for r in b["rows"])
    print(b["name"][:100], "total samples", tot)
    # aggregate stall reason columns if present
    reasons = [(n, i) for n, i in h.items() if n.startswith("Stall") or n.startswith("stall")]
    ranked = sorted(b["rows"], key=lambda r: -int(r[si] or 0))[:top]
    for r in ranked:
        extra = ""
        best = sorted(((int(r[i] or 0), n) for n, i in h.items()
                       if i > si and n not in ("Instructions Executed", "Thread Instructions Executed",
                                               "Predicated-On Thread Instructions Executed", "Avg. Threads Executed",
                                               "Avg. Predicated-On Threads Executed", "Divergent Branches")
                       and (r[i] or "0").isdigit()), reverse=True)[:2]
        extra = " ".join(f"{n}={v}" for v, n in best if v)
        print(f"{100*int(r[si] or 0)/max(tot,1):5.1f}%  thr={r[ti]:>5s} inst={r[ii]:>9s}  {r[h['Source']].strip()[:70]:70s} {extra[:90]}")
    break
