#!/usr/bin/env python3
"""Scan throughput for documents made of one repeated line pattern (tuning / adversarial cells).
   python tools/bench_pattern.py 'PATTERN' [--doc-bytes 4096] [--mib 256]   -- PATTERN is a Python bytes literal body, \\n appended"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import operator_builder_b200 as ob
ap = argparse.ArgumentParser(); ap.add_argument("pattern"); ap.add_argument("--doc-bytes", type=int, default=4096); ap.add_argument("--mib", type=int, default=256)
ap.add_argument("--iters", type=int, default=4); ap.add_argument("--mode", type=int, default=0)
a = ap.parse_args()
line = a.pattern.encode().decode("unicode_escape").encode("latin1") + b"\n"
doc = line * max(1, a.doc_bytes // len(line))
ndocs = max(1, (a.mib << 20) // len(doc)); n = ndocs * len(doc)
dev = torch.device("cuda:0"); sc = ob.Scanner(0); sc.set_mode(a.mode); st = torch.cuda.current_stream().cuda_stream
d_bytes = torch.from_numpy(np.tile(np.frombuffer(doc, dtype=np.uint8), ndocs)).to(dev)
d_off = (torch.arange(ndocs + 1, dtype=torch.int64) * len(doc)).to(dev)
d_out = torch.empty(n, dtype=torch.int64, device=dev); d_toff = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
d_status = torch.zeros(4, dtype=torch.int32, device=dev); d_counts = torch.zeros(2, dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters):
    e0.record(); sc.lex_batch_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, n, d_out.data_ptr(), n, d_toff.data_ptr(), d_status.data_ptr(), d_counts.data_ptr(), st); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print(f"{len(line):5d}-byte lines, doc {len(doc)} B: {ms:8.3f} ms  {n / ms / 1e6:8.1f} GB/s  tuples/B {int(d_toff[-1]) / n:.3f} status {d_status.tolist()}  {line[:50]!r}")
