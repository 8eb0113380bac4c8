#!/usr/bin/env python3
"""profiles/traffic.json from an ncu metrics run over tools/profile_run.py:
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/traffic.csv python tools/profile_run.py --docs 262144 --iters 3
   python tools/make_traffic.py gpurun_out/traffic.csv 262144 4096 > profiles/traffic.json
Sums DRAM reads + writes over the kernels of the LAST scan (from its k_wtile_index launch on)."""
import csv, json, sys, collections

path, docs, doc_bytes = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r and "Metric Name" in r)
ix = {n: i for i, n in enumerate(hdr)}
launches = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) < len(hdr) or not r[ix["ID"]].isdigit():
        continue
    L = launches.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1)
    L[r[ix["Metric Name"]]] = val * scale
ids = list(launches)
starts = [i for i in ids if "k_wtile_index" in launches[i]["name"]]
last = [i for i in ids if i >= starts[-1]]
rd = sum(launches[i].get("dram__bytes_read.sum", 0) for i in last)
wr = sum(launches[i].get("dram__bytes_write.sum", 0) for i in last)
us = sum(launches[i].get("gpu__time_duration.sum", 0) for i in last)
n = docs * doc_bytes
per = collections.OrderedDict()
for i in last:
    k = launches[i]["name"][:40]
    a = per.setdefault(k, [0.0, 0.0, 0.0])
    a[0] += launches[i].get("dram__bytes_read.sum", 0) / 1e6; a[1] += launches[i].get("dram__bytes_write.sum", 0) / 1e6
    a[2] += launches[i].get("gpu__time_duration.sum", 0)
print(json.dumps({"dram_bytes_per_input_byte": round((rd + wr) / n, 4), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr), "input_bytes": n,
                  "kernels_in_scan": len(last), "us_all_kernels_under_ncu": round(us, 1),
                  "per_kernel_MB_read_write_us": {k: [round(v[0], 2), round(v[1], 2), round(v[2], 1)] for k, v in per.items()},
                  "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over every kernel of one mode-0 scan, {docs:,} docs x {doc_bytes:,} B "
                            "(tools/make_traffic.py); bench.py scales it by the input bytes of its launch"}, indent=1))
