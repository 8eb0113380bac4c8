"""SURVEY.md 8(f) ranks 2 and 4 -- the data formats either side of the scan, on the device.

rank 2  Manifest.LoadContent's collection rewrite (manifests/manifest.go:89-95) == Python's
        bytes.replace twice (Go's strings.ReplaceAll: non-overlapping, left to right)
rank 4  Manifest.ExtractManifests (manifests/manifest.go:57-80), restated line for line below"""
import random

import numpy as np
import pytest

from tests import corpus_util as cu

pytestmark = pytest.mark.gpu


def extract_manifests(content: bytes):
    """manifest.go:57-80, verbatim semantics (strings.Split / TrimRight(" "))"""
    out, cur = [], b""
    for line in content.split(b"\n"):
        if line.rstrip(b" ") == b"---":
            if len(cur) > 0:
                out.append(cur)
                cur = b""
        else:
            cur = cur + b"\n" + line
    if len(cur) > 0:
        out.append(cur)
    return out


def corpus():
    import operator_builder_b200 as ob
    rng = random.Random(12)
    docs = [d for _p, d in cu.fixtures()]
    data0, _ = ob.generate_corpus_host(400, 4096, flavour=1)
    docs += [data0.tobytes()[i * 4096:(i + 1) * 4096] for i in range(400)]
    docs += [b"", b"---", b"---\n", b"\n---\n\n--- \n---x\n ---\na\n---", b"a\n---  \nb\n---\n---\nc", b"collectionField", b"xcollectionFieldcollectionField+operator-builder:collection:fieldx",
             b"+operator-builder:collection:fiel", b"collectionFiel", b"+operator-builder:collection:field", b"++operator-builder:collection:field:name=collectionField\n"]
    for _ in range(300):
        parts = [rng.choice([b"---", b"--- ", b"----", b"a: b", b"", b"# +operator-builder:collection:field:name=x,type=string",
                             b"# +operator-builder:resource:collectionField=p,value=1,include", b"collection", b"Field", b"  ---"]) for _ in range(rng.randint(0, 12))]
        docs.append(b"\n".join(parts) + rng.choice([b"", b"\n"]))
    return docs


def to_dev(docs):
    import torch
    data = np.frombuffer(b"".join(docs) + b"\0", dtype=np.uint8)[:-1].copy()
    off = np.zeros(len(docs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(d) for d in docs])
    dev = torch.device("cuda:0")
    return torch.from_numpy(np.concatenate([data, np.zeros(64, np.uint8)])).to(dev), torch.from_numpy(off).to(dev), off


def test_collection_rewrite():
    import torch
    import operator_builder_b200 as ob
    from operator_builder_b200 import _native
    docs = corpus()
    d_bytes, d_off, off = to_dev(docs)
    sc = ob.Scanner(0)
    L = _native.lib()
    d_out = torch.zeros(int(off[-1]) + 64, dtype=torch.uint8, device=d_bytes.device)
    d_noff = torch.zeros(len(docs) + 1, dtype=torch.int64, device=d_bytes.device)
    st = torch.cuda.current_stream().cuda_stream
    assert L.obm_rewrite_collection_markers_device(sc.handle, d_bytes.data_ptr(), d_off.data_ptr(), len(docs), d_out.data_ptr(), int(off[-1]) + 64,
                                                   d_noff.data_ptr(), st) == 0
    torch.cuda.synchronize()
    noff = d_noff.cpu().numpy()
    out = d_out.cpu().numpy().tobytes()
    changed = 0
    for i, doc in enumerate(docs):
        want = doc.replace(b"+operator-builder:collection:field", b"+operator-builder:field").replace(b"collectionField", b"field")
        got = out[int(noff[i]):int(noff[i + 1])]
        assert got == want, (i, doc[:120])
        changed += want != doc
    assert changed > 400
    sc.close()


def _rewrite_on_device(docs, two_pass=False, byte_shift=0):
    import os
    import torch
    import operator_builder_b200 as ob
    from operator_builder_b200 import _native
    d_bytes, d_off, off = to_dev(docs)
    if byte_shift:  # an input buffer that is not 16-byte aligned: the library takes r01's two passes
        raw = torch.zeros(d_bytes.numel() + 32, dtype=torch.uint8, device=d_bytes.device)
        raw[byte_shift:byte_shift + d_bytes.numel()] = d_bytes
        d_bytes = raw[byte_shift:]
    sc = ob.Scanner(0)
    L = _native.lib()
    d_out = torch.full((int(off[-1]) + 64,), 0xEE, dtype=torch.uint8, device=d_bytes.device)
    d_noff = torch.zeros(len(docs) + 1, dtype=torch.int64, device=d_bytes.device)
    st = torch.cuda.current_stream().cuda_stream
    if two_pass:
        os.environ["OBM_REWRITE_TWO_PASS"] = "1"
    try:
        rc = L.obm_rewrite_collection_markers_device(sc.handle, d_bytes.data_ptr(), d_off.data_ptr(), len(docs), d_out.data_ptr(), int(off[-1]) + 64, d_noff.data_ptr(), st)
    finally:
        os.environ.pop("OBM_REWRITE_TWO_PASS", None)
    assert rc == 0
    torch.cuda.synchronize()
    noff = d_noff.cpu().numpy()
    out = d_out.cpu().numpy().tobytes()
    sc.close()
    return noff, out


def _py_rewrite(doc):
    return doc.replace(b"+operator-builder:collection:field", b"+operator-builder:field").replace(b"collectionField", b"field")


def test_collection_rewrite_chunk_and_document_edges():
    """the one-pass kernel works on 12 KiB chunks of the packed batch: patterns across chunk borders (every shift), patterns cut by
    a document border (no match), dense patterns, empty documents, a batch that ends on a chunk border"""
    rng = random.Random(5)
    P1, P2 = b"+operator-builder:collection:field", b"collectionField"
    CH = 12288
    docs = []
    for shift in range(0, 40):  # a pattern that starts `shift` bytes before a chunk border
        for pat in (P1, P2):
            used = sum(len(d) for d in docs)
            pad = (-used - shift) % CH
            if pad < 8:
                pad += CH
            docs.append(b"a" * (pad - 3) + b"\n# " + pat + b":x=1\n")
    docs += [P2 * 900, (P1 + b"\n") * 400, P2[:7], P2[7:], P1[:20], P1[20:] + P2, b"", b"", P2[:10], b"Field", b"+operator-builder:collection:", b"field",
             b"collectionFcollectionField", b"+operator-builder:+operator-builder:collection:fieldcollectionFieldcollectionField"]
    for _ in range(400):
        n = rng.randint(0, 6)
        docs.append(b"".join(rng.choice([P1, P2, P1[:rng.randint(0, 34)], P2[:rng.randint(0, 15)], b"\n", b"x" * rng.randint(0, 700), b"F", b"+", b"c"]) for _ in range(n)))
    used = sum(len(d) for d in docs)
    docs.append(b"y" * ((-used) % CH))  # the batch ends exactly on a chunk border
    docs += [b""]
    want = [_py_rewrite(d) for d in docs]
    for kw in ({}, {"two_pass": True}, {"byte_shift": 3}):
        noff, out = _rewrite_on_device(docs, **kw)
        assert int(noff[-1]) == sum(len(w) for w in want), kw
        for i, w in enumerate(want):
            assert out[int(noff[i]):int(noff[i + 1])] == w, (kw, i, docs[i][:80])
        assert out[int(noff[-1]):int(noff[-1]) + 16] == b"\xee" * 16, kw  # nothing written past the end
    docs2 = docs[:-2] + [b"z" * 5, P2]  # ... and one that ends inside a pattern's chunk
    noff, out = _rewrite_on_device(docs2)
    for i, d in enumerate(docs2):
        assert out[int(noff[i]):int(noff[i + 1])] == _py_rewrite(d), i


def test_document_split():
    import torch
    import operator_builder_b200 as ob
    from operator_builder_b200 import _native
    docs = corpus()
    d_bytes, d_off, off = to_dev(docs)
    sc = ob.Scanner(0)
    L = _native.lib()
    cap = int(off[-1]) // 4 + len(docs) + 16
    d_rec = torch.zeros(cap * 4, dtype=torch.int32, device=d_bytes.device)
    d_roff = torch.zeros(len(docs) + 1, dtype=torch.int64, device=d_bytes.device)
    st = torch.cuda.current_stream().cuda_stream
    assert L.obm_split_docs_device(sc.handle, d_bytes.data_ptr(), d_off.data_ptr(), len(docs), d_rec.data_ptr(), cap, d_roff.data_ptr(), st) == 0
    torch.cuda.synchronize()
    roff = d_roff.cpu().numpy()
    rec = d_rec.cpu().numpy().view(np.uint32).reshape(-1, 4)
    total = 0
    for i, doc in enumerate(docs):
        want = extract_manifests(doc)
        got = [b"\n" + doc[int(r[1]):int(r[2])] for r in rec[int(roff[i]):int(roff[i + 1])]]
        assert got == want, (i, doc[:120], got, want)
        assert all(int(r[0]) == i for r in rec[int(roff[i]):int(roff[i + 1])])
        total += len(want)
    assert total == int(roff[-1]) and total > 500
    sc.close()
