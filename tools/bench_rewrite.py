#!/usr/bin/env python3
"""Collection rewrite on a 1 GiB HBM-resident corpus of the collection flavour: the one-pass chunk kernel
(obm_rewrite.cuh) next to r01's two passes over the documents and to the scan of the same bytes.  Device time by CUDA
events around the C-ABI calls (3 warm-ups, 10 steps; 1 GiB > L2)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import operator_builder_b200 as ob
from operator_builder_b200 import _native

ndocs, doc_bytes, steps = int(os.environ.get("OBM_ROWS_DOCS", 262144)), 4096, 10
dev = torch.device("cuda:0")
sc = ob.Scanner(0)
L = _native.lib()
st = torch.cuda.current_stream().cuda_stream
n = ndocs * doc_bytes
d_bytes = torch.empty(n + 64, dtype=torch.uint8, device=dev)
d_off = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
sc.generate_corpus_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, doc_bytes, 0, 1, st)
d_rw = torch.empty(n + 64, dtype=torch.uint8, device=dev)
d_noff = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
cap = n // 16
d_out = torch.empty(cap, dtype=torch.int64, device=dev)
d_toff = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def rewrite():
    assert L.obm_rewrite_collection_markers_device(sc.handle, d_bytes.data_ptr(), d_off.data_ptr(), ndocs, d_rw.data_ptr(), n + 64, d_noff.data_ptr(), st) == 0


ms_one = timed(rewrite)
out_one = d_rw[: int(d_noff[-1])].clone()
off_one = d_noff.clone()
os.environ["OBM_REWRITE_TWO_PASS"] = "1"
ms_two = timed(rewrite)
os.environ.pop("OBM_REWRITE_TWO_PASS")
same = bool(torch.equal(out_one, d_rw[: int(d_noff[-1])])) and bool(torch.equal(off_one, d_noff))
ms_scan = timed(lambda: sc.lex_batch_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, n, d_out.data_ptr(), cap, d_toff.data_ptr(), None, None, st))
print(json.dumps({"corpus": f"{ndocs} docs x {doc_bytes} B, collection flavour, HBM resident", "one_pass_ms": ms_one, "one_pass_input_GBps": n / ms_one / 1e6,
                  "two_pass_ms": ms_two, "two_pass_input_GBps": n / ms_two / 1e6, "outputs_identical": same, "out_bytes": int(d_noff[-1]),
                  "scan_ms": ms_scan, "rewrite_plus_scan_over_scan": (ms_one + ms_scan) / ms_scan}))
