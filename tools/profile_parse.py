#!/usr/bin/env python3
"""ncu driver for the next-row kernels: scan `--docs` synthetic manifests, then run the device parser `--iters` times."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import operator_builder_b200 as ob
ap = argparse.ArgumentParser(); ap.add_argument("--docs", type=int, default=262144); ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0"); sc = ob.Scanner(0); st = torch.cuda.current_stream().cuda_stream
n = a.docs * 4096
d_bytes = torch.empty(n, dtype=torch.uint8, device=dev); d_off = torch.empty(a.docs + 1, dtype=torch.int64, device=dev)
sc.generate_corpus_device(d_bytes.data_ptr(), d_off.data_ptr(), a.docs, 4096, 0, 0, st)
cap = n // 16
d_out = torch.empty(cap, dtype=torch.int64, device=dev); d_toff = torch.empty(a.docs + 1, dtype=torch.int64, device=dev)
sc.lex_batch_device(d_bytes.data_ptr(), d_off.data_ptr(), a.docs, n, d_out.data_ptr(), cap, d_toff.data_ptr(), None, None, st)
reg = ob.Registry()
d_res = torch.empty(a.docs * 12 * 32, dtype=torch.uint8, device=dev); d_args = torch.empty(a.docs * 48 * 16, dtype=torch.uint8, device=dev)
d_roff = torch.empty(a.docs + 1, dtype=torch.int64, device=dev); d_tot = torch.zeros(2, dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters):
    e0.record()
    sc.parse_batch_device(reg, d_bytes.data_ptr(), d_off.data_ptr(), a.docs, 0, d_out.data_ptr(), d_toff.data_ptr(), d_res.data_ptr(), a.docs * 12,
                          d_args.data_ptr(), a.docs * 48, d_roff.data_ptr(), d_tot.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    print(f"parse iter {i}: {e0.elapsed_time(e1):.3f} ms  totals={d_tot.tolist()}")
