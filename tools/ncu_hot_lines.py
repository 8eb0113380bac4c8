#!/usr/bin/env python3
"""Per-source-line instruction counts from an .ncu-rep captured with --import-source on (-lineinfo build).
   python tools/ncu_hot_lines.py rep.ncu-rep [N]"""
import csv
import subprocess
import sys
import collections

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
agg = collections.defaultdict(lambda: [0, 0, 0])  # inst, thread inst, samples
total = [0, 0, 0]
ie = ti = sm = None
cur_line = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        cur_file = r[1].split("/")[-1]
        continue
    if "Instructions Executed" in r:
        ie, ti, sm = r.index("Instructions Executed"), r.index("Thread Instructions Executed"), r.index("# Samples")
        continue
    if ie is None or len(r) <= ti:
        continue
    if r[0].strip().isdigit():
        cur_line = (cur_file, int(r[0]), r[1].strip()[:90])
        continue
    if r[0] == "" and cur_line and r[ie].isdigit():
        a = agg[cur_line]
        a[0] += int(r[ie]); a[1] += int(r[ti]); a[2] += int(r[sm]) if r[sm].isdigit() else 0
        total[0] += int(r[ie]); total[1] += int(r[ti]); total[2] += int(r[sm]) if r[sm].isdigit() else 0
print(f"total warp-inst {total[0]:,}  thread-inst {total[1]:,}  samples {total[2]:,}")
byfile = collections.defaultdict(int)
for (f, l, s), a in agg.items():
    byfile[f] += a[0]
print({k: f"{v / max(total[0], 1):.1%}" for k, v in byfile.items()})
for (f, l, s), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{a[0] / max(total[0], 1):6.2%} inst {a[2] / max(total[2], 1):6.2%} smp  lanes {a[1] / max(a[0], 1):5.1f}  {f}:{l}  {s}")
