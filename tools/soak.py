#!/usr/bin/env python3
"""Randomised soak on the GPU: many batches of fuzzed documents (malformed, well-formed, non-ASCII, dense, large),
modes 0 / 1 / 2 must agree tuple for tuple and (sampled) decode to the oracle's lexeme stream.
   python tools/soak.py --seconds 150"""
import argparse, os, random, sys, time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import operator_builder_b200 as ob
import oracle
from tests import corpus_util as cu

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=120)
ap.add_argument("--seed", type=int, default=20260921)
a = ap.parse_args()
rng = random.Random(a.seed)
sc = ob.Scanner(0)
t0 = time.time(); batches = docs_total = bytes_total = checked = 0
while time.time() - t0 < a.seconds:
    kind = rng.randrange(6)
    docs = []
    n = rng.choice([1, 3, 40, 200, 700])
    for _ in range(n):
        r = rng.random()
        if kind == 0: d = cu.fuzz_doc(rng, max_len=rng.choice([5, 60, 400, 3000]), non_ascii=rng.random() < 0.15)
        elif kind == 1: d = cu.fuzz_doc_valid(rng)
        elif kind == 2: d = b"".join(cu.fuzz_doc_valid(rng) for _ in range(rng.randint(1, 12)))  # up to ~40 KB: tile and large paths mixed
        elif kind == 3: d = (b"# +s:a=1 +t:b=\"x y\",c\n" * rng.randint(1, 900)) if r < 0.5 else cu.fuzz_doc(rng, max_len=20000)
        elif kind == 4: d = b"\n".join(rng.choice([b"#", b"# x", b"k: v", b"+a:b", b"'q' // c", b"- /p/q", b""]) for _ in range(rng.randint(0, 3000)))
        else: d = rng.choice([cu.fuzz_doc_valid(rng), cu.fuzz_doc(rng, max_len=300), b"", b"x" * rng.randint(16000, 17000) + b" # +a:b\n"])
        docs.append(d)
    data = np.frombuffer(b"".join(docs) + b"\0", dtype=np.uint8)[:-1].copy()
    off = np.zeros(len(docs) + 1, dtype=np.uint64); off[1:] = np.cumsum([len(d) for d in docs])
    res = []
    for mode in (0, 1, 2):
        sc.set_mode(mode)
        r = sc.lex_batch(data if len(data) else np.zeros(1, np.uint8), off)
        res.append((r.tuples.copy(), r.doc_tuple_off.copy(), dict(r.stats)))
    for k in (1, 2):
        if not (np.array_equal(res[0][0], res[k][0]) and np.array_equal(res[0][1], res[k][1])):
            bad = next(i for i in range(len(docs)) if not np.array_equal(res[0][0][int(res[0][1][i]):int(res[0][1][i + 1])], res[k][0][int(res[k][1][i]):int(res[k][1][i + 1])]) or res[0][1][i + 1] != res[k][1][i + 1])
            open("gpurun_out/soak_fail.bin", "wb").write(docs[bad])
            print(f"MISMATCH batch {batches} kind {kind} mode 0 vs {k} doc {bad} len {len(docs[bad])} ndocs {len(docs)}: {docs[bad][:200]!r}")
            sys.exit(1)
        for key in ("n_markers", "n_lexemes", "n_tuples"):
            assert res[0][2][key] == res[k][2][key], (key, res[0][2], res[k][2])
    for i in rng.sample(range(len(docs)), min(3, len(docs))):  # sampled oracle check
        t = res[0][0][int(res[0][1][i]):int(res[0][1][i + 1])]
        want = oracle.lex_raw(docs[i])
        got = ob.decode_doc_raw(docs[i], t)
        if got != want:
            open("gpurun_out/soak_fail.bin", "wb").write(docs[i])
            print(f"ORACLE MISMATCH batch {batches} kind {kind} doc {i}: {docs[i][:200]!r}"); sys.exit(1)
        checked += 1
    batches += 1; docs_total += len(docs); bytes_total += len(data)
print(f"soak ok: {batches} batches, {docs_total} docs, {bytes_total / 1e6:.1f} MB, {checked} oracle-checked, {time.time() - t0:.0f} s")
