#!/usr/bin/env python3
"""Top source lines by stall samples with the dominant stall reasons (ncu source page, SASS rows summed per CUDA line).
   python tools/ncu_stall_lines.py rep.ncu-rep [N]"""
import csv, subprocess, collections, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None
agg = collections.defaultdict(lambda: collections.Counter())
reasons = ["stall_barrier", "stall_branch_resolving", "stall_dispatch", "stall_lg", "stall_long_sb", "stall_math", "stall_membar", "stall_mio",
           "stall_no_inst", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_sleep", "stall_wait"]
for r in rows:
    if "Instructions Executed" in r: hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0].strip().isdigit(): cur = (int(r[0]), r[1].strip()[:90]); continue
    if r[0] == "" and cur:
        for k in reasons:
            if k not in hdr: continue
            v = r[hdr[k]]
            if v.isdigit(): agg[cur][k] += int(v)
tot = sum(sum(c.values()) for c in agg.values())
allr = collections.Counter()
for c in agg.values(): allr.update(c)
print("all samples:", tot, " ".join(f"{k[6:]}={v / tot:.1%}" for k, v in allr.most_common(8)))
for cur, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:topn]:
    s = sum(c.values())
    print(f"{s / tot:5.1%} :{cur[0]:<4d} {' '.join(f'{k[6:]}={v / s:.0%}' for k, v in c.most_common(3)):45s} {cur[1]}")
