#!/usr/bin/env python3
"""Instruction share / lane utilisation per source file and top lines of one file.
   python tools/ncu_by_file.py rep.ncu-rep ndocs [file [N]]"""
import csv, subprocess, collections, sys
rep, ndocs = sys.argv[1], int(sys.argv[2])
which = sys.argv[3] if len(sys.argv) > 3 else None
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
by_samples = bool(int(__import__("os").environ.get("BY_SAMPLES", "0")))
kfilter = sys.argv[5] if len(sys.argv) > 5 else None  # substring of the kernel name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files = {}
for f in ["obm_warp.h", "obm_warp_core.h", "obm_warp.cuh", "obm_core.h", "obm_tile.h", "obm_fast.cuh", "obm_lib.cu", "obm_pipe.h", "obm_pipe.cuh"]:
    for i, l in enumerate(open("operator-builder_b200/csrc/" + f).read().split("\n"), 1):
        files.setdefault((i, l.strip()[:60]), f)
ie = None; cur = None
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0]
lines = collections.defaultdict(lambda: [0, 0, 0])
kcur = True
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        kcur = (kfilter is None) or (kfilter in r[1])
        continue
    if not kcur:
        continue
    if "Instructions Executed" in r:
        ie, ti, sm = r.index("Instructions Executed"), r.index("Thread Instructions Executed"), r.index("# Samples"); continue
    if ie is None or len(r) <= ie: continue
    if r[0].strip().isdigit():
        cur = (files.get((int(r[0]), r[1].strip()[:60]), "other"), int(r[0]), r[1].strip()[:100]); continue
    if r[0] == "" and cur and r[ie].isdigit():
        s = int(r[sm]) if r[sm].isdigit() else 0
        a = agg[cur[0]]; a[0] += int(r[ie]); a[1] += int(r[ti]); a[2] += s; tot[0] += int(r[ie]); tot[1] += s
        l = lines[cur]; l[0] += int(r[ie]); l[1] += int(r[ti]); l[2] += s
for k, a in agg.items():
    print(f"{k:14s} {a[0]/tot[0]:6.1%} inst  lanes {a[1]/max(a[0],1):5.1f}  {a[0]/ndocs:7.0f} warp-inst/doc  {a[2]/max(tot[1],1):6.1%} samples")
print(f"total {tot[0]/ndocs:.0f} warp-inst/doc")
if which:
    for k, a in sorted([(k, a) for k, a in lines.items() if k[0] == which], key=lambda kv: -(kv[1][2] if by_samples else kv[1][0]))[:topn]:
        print(f"{a[0]/ndocs:7.1f}/doc lanes {a[1]/max(a[0],1):5.1f} smp {a[2]/max(tot[1],1):5.1%} :{k[1]} {k[2]}")
