#!/usr/bin/env python3
"""Small driver for ncu: generates `--docs` synthetic manifests in HBM and runs the scan `--iters` times.
   ncu --set full -k regex:k_tile_scan -s 1 -c 1 -o gpurun_out/prof python tools/profile_run.py --docs 65536"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import operator_builder_b200 as ob

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=65536)
ap.add_argument("--doc-bytes", type=int, default=4096)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--flavour", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
sc = ob.Scanner(0)
sc.set_mode(a.mode)
st = torch.cuda.current_stream().cuda_stream
n = a.docs * a.doc_bytes
d_bytes = torch.empty(n, dtype=torch.uint8, device=dev)
d_off = torch.empty(a.docs + 1, dtype=torch.int64, device=dev)
sc.generate_corpus_device(d_bytes.data_ptr(), d_off.data_ptr(), a.docs, a.doc_bytes, 0, a.flavour, st)
cap = n // 16
d_out = torch.empty(cap, dtype=torch.int64, device=dev)
d_toff = torch.empty(a.docs + 1, dtype=torch.int64, device=dev)
d_status = torch.zeros(4, dtype=torch.int32, device=dev)
d_counts = torch.zeros(2, dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters):
    e0.record()
    sc.lex_batch_device(d_bytes.data_ptr(), d_off.data_ptr(), a.docs, n, d_out.data_ptr(), cap, d_toff.data_ptr(),
                        d_status.data_ptr(), d_counts.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"iter {i}: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s  tuples={int(d_toff[-1])} status={d_status.tolist()} counts={d_counts.tolist()}")
