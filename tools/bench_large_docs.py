import sys, os, random, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import operator_builder_b200 as ob
rng = random.Random(5)
pool = [b"key: value\n", b"  - item  # +operator-builder:field:name=a.b,type=int,default=3\n", b"\n", b"path: /a/b+c\n", b"x" * 300 + b"\n", b"# plain comment\n"]
def mk(target):
    parts, n = [], 0
    while n < target:
        ln = rng.choice(pool); parts.append(ln); n += len(ln)
    return b"".join(parts)
sc = ob.Scanner(0)
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
for size, cnt in ((100_000, 640), (1_000_000, 64), (16_000_000, 4)):
    docs = [mk(size) for _ in range(cnt)]
    data = np.frombuffer(b"".join(docs), dtype=np.uint8)
    off = np.zeros(cnt + 1, dtype=np.int64); off[1:] = np.cumsum([len(d) for d in docs])
    n = int(off[-1])
    d_bytes = torch.from_numpy(data.copy()).to(dev); d_off = torch.from_numpy(off).to(dev)
    cap = n
    d_out = torch.empty(cap, dtype=torch.int64, device=dev); d_toff = torch.empty(cnt + 1, dtype=torch.int64, device=dev)
    d_status = torch.zeros(4, dtype=torch.int32, device=dev); d_counts = torch.zeros(2, dtype=torch.int64, device=dev)
    for mode in (0, 1):
        sc.set_mode(mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for i in range(3):
            e0.record()
            sc.lex_batch_device(d_bytes.data_ptr(), d_off.data_ptr(), cnt, n, d_out.data_ptr(), cap, d_toff.data_ptr(), d_status.data_ptr(), d_counts.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{cnt} docs x {size} B mode {mode}: {best:.2f} ms  {n/best/1e6:.2f} GB/s tuples={int(d_toff[-1])} status={d_status.tolist()}")
