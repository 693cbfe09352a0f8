#!/usr/bin/env python
"""Stall samples and executed instructions per CUDA source line of one kernel
(needs -lineinfo and ncu --import-source on).
    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep bucket_build [N]
"""
import csv, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
acc = collections.OrderedDict()
cur = None
for r in rows:
    if "# Samples" in r:
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if not r[0].strip().isdigit():   # SASS rows have an empty line number
        continue
    src = r[1]
    num = lambda x: int(x) if x.isdigit() else 0
    s, ie = num(r[hdr["# Samples"]]), num(r[hdr["Instructions Executed"]])
    if True:                        # a CUDA source line: its totals are the sums of its SASS
        key = (r[0], src.strip())
        acc.setdefault(key, [0, 0, collections.Counter()])
        acc[key][0] += s
        acc[key][1] += ie
        for n, i in hdr.items():
            if n.startswith("stall_") and "Not Issued" not in n and (r[i] or "0").isdigit():
                acc[key][2][n] += int(r[i])
tot = sum(v[0] for v in acc.values()) or 1
toti = sum(v[1] for v in acc.values()) or 1
print("samples", tot, "warp insts", toti)
for (ln, src), (s, ie, st) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    why = " ".join(f"{n[6:]}={100 * v // max(s, 1)}%" for n, v in st.most_common(2))
    print(f"{100 * s / tot:5.1f}%  inst {100 * ie / toti:5.1f}%  L{ln:>4s} {src[:84]:84s} {why}")
