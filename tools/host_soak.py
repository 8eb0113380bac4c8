#!/usr/bin/env python3
"""CPU-only soak of the kernels' logic (tests/hostsim replays both device organisations with the product's own
host/device headers): random batches of malformed / well-formed / dense / non-ASCII / near-tile-size documents at random
alignments, every document compared with the sequential lexer.   python tools/host_soak.py --seconds 600"""
import argparse, os, pickle, random, sys, time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import hostsim, corpus_util as cu

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=300)
ap.add_argument("--seed", type=int, default=99001)
a = ap.parse_args()
rng = random.Random(a.seed)
VALID = [b"\xc3\xa9", b"\xe2\x82\xac", b"\xd9\xa3", b"\xf0\x9f\x98\x80", b"\xe4\xb8\xad", b"\xc2\xbd"]


def mixed(rng):
    k = rng.randrange(7)
    if k == 0: return cu.fuzz_doc(rng, max_len=rng.choice([5, 60, 400, 3000]), non_ascii=rng.random() < 0.3)
    if k == 1: return cu.fuzz_doc_valid(rng)
    if k == 2: return b"".join(cu.fuzz_doc_valid(rng) for _ in range(rng.randint(1, 10)))
    if k == 3: return b"# +s:a=1 +t:b=\"x y\",c\n" * rng.randint(1, 500)
    if k == 4: return b"\n".join(rng.choice([b"#", b"# x", b"k: v", b"+a:b", b"'q' // c", b"- /p/q", b"", "# é".encode(), "k: ü # +a:b=ö".encode()])
                                 for _ in range(rng.randint(0, 1500)))
    if k == 5:
        d = bytearray(cu.fuzz_doc_valid(rng))
        for _ in range(rng.randint(0, 4)):
            if d:
                i = rng.randrange(len(d)); d[i:i] = rng.choice(VALID)
        return bytes(d)
    return rng.choice([b"", b"x" * rng.randint(16000, 17000) + b" # +a:b\n", cu.fuzz_doc(rng, max_len=20000)])


t0 = time.time(); nb = nd = 0
while time.time() - t0 < a.seconds:
    docs = [mixed(rng) for _ in range(rng.choice([1, 5, 40, 120]))]
    skew = rng.randint(0, 15)
    for pipeline in (0, 1):
        tup, toff, st = hostsim.tile_batch(docs, skew, pipeline=pipeline)
        for i, d in enumerate(docs):
            if not np.array_equal(tup[int(toff[i]):int(toff[i + 1])], hostsim.lex_doc(d)):
                pickle.dump((docs, skew, i), open("/tmp/host_soak_fail.pkl", "wb"))
                print("MISMATCH organisation", pipeline, "document", i, len(d), d[:120]); sys.exit(1)
    nb += 1; nd += len(docs)
print(f"host soak ok: {nb} batches, {nd} documents, {time.time() - t0:.0f} s")
