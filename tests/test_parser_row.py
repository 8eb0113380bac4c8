"""SURVEY.md 8(f) rank 1 -- the parser token contract (8(a) a17) on the tuple stream.

obm_parse_doc (csrc/obm_parse.cpp, mirror of internal/markers/parser/state.go:13-175) over tuples vs the
Python restatement oracle/parser_oracle.py over the oracle lexer's lexemes.  No reference test pins the parser
(SURVEY section 4): both sides follow the source; this checks they agree and that MarkerText -- what
markers/markers.go:198-222 splices back into YAML comments -- is byte-exact."""
import ctypes
import random

import numpy as np
import pytest

from tests import corpus_util as cu
from tests import hostsim


def results_via_tuples(doc, tuples):
    import operator_builder_b200 as ob
    return ob.parse_doc_raw(REG(), doc, tuples)


_REG = None


def REG():
    global _REG
    if _REG is None:
        import operator_builder_b200 as ob
        _REG = ob.Registry()
    return _REG


RES_DT = np.dtype([("doc", "<u4"), ("tuple", "<u4"), ("text_off", "<u4"), ("text_len", "<u4"), ("reg_id", "<u2"), ("nargs", "<u2"),
                   ("arg_base", "<u4"), ("flags", "<u4"), ("aux", "<u4")])
ARG_DT = np.dtype([("name_off", "<u4"), ("val_off", "<u4"), ("val_len", "<u4"), ("name_len", "<u2"), ("kind", "u1"), ("flags", "u1")])
R_HOST = 16
STATS = {"docs": 0, "host": 0}


def format_records(doc, tuples, res, args):
    """compact device records of one document -> the byte format of obm_parse_doc (obm_results_format_doc)"""
    import operator_builder_b200 as ob
    L = ob._native.lib()
    tuples = np.ascontiguousarray(tuples, dtype=np.uint64)
    res = np.ascontiguousarray(res); args = np.ascontiguousarray(args)
    out = ctypes.POINTER(ctypes.c_uint8)(); outlen = ctypes.c_uint64()
    L.obm_results_format_doc.restype = ctypes.c_int64
    L.obm_results_format_doc.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                                         ctypes.c_uint64, ctypes.c_void_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8)), ctypes.POINTER(ctypes.c_uint64)]
    n = L.obm_results_format_doc(REG().handle, doc, len(doc), tuples.ctypes.data, len(tuples), res.ctypes.data, len(res),
                                 args.ctypes.data if len(args) else None, ctypes.byref(out), ctypes.byref(outlen))
    assert n >= 0
    data = ctypes.string_at(out, outlen.value)
    L.obm_free(out)
    return data


def results_via_device_walk(doc, tuples):
    """csrc/obm_parse_dev.h (the code k_parse_docs runs, one thread per document) on the host"""
    import operator_builder_b200 as ob
    L = ob._native.lib()
    tuples = np.ascontiguousarray(tuples, dtype=np.uint64)
    res = np.zeros(len(tuples) + 2, dtype=RES_DT); args = np.zeros(len(tuples) + 2, dtype=ARG_DT)
    nargs = ctypes.c_uint64()
    L.obm_parse_doc_records.restype = ctypes.c_int64
    L.obm_parse_doc_records.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p,
                                        ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
    n = L.obm_parse_doc_records(REG().handle, doc, tuples.ctypes.data, len(tuples), 0, res.ctypes.data, len(res), args.ctypes.data, len(args),
                                ctypes.byref(nargs))
    assert n >= 0
    STATS["docs"] += 1
    STATS["host"] += int(n == 1 and (int(res[0]["flags"]) & R_HOST) != 0)
    return format_records(doc, tuples, res[:n], args[:nargs.value])


def check(oracle, doc, tuples=None):
    from oracle import parser_oracle as po
    want = po.serialize(po.parse(oracle.lex(doc), po.OPERATOR_BUILDER_REGISTRY))
    tuples = hostsim.lex_doc(doc) if tuples is None else tuples
    got = results_via_tuples(doc, tuples)
    assert got == want, (doc, got, want)
    dev = results_via_device_walk(doc, tuples)
    assert dev == want, (doc, dev, want)
    return want


def test_marker_text_and_args(oracle):
    from oracle import parser_oracle as po
    doc = (b"spec:\n  replicas: 2  # +operator-builder:field:name=webstore.replicas,default=2,type=int\n"
           b"  # +operator-builder:field:name=app.label,type=string,default=\"webstore\",description=`multi\n  # line`\n"
           b"# +operator-builder:resource:field=provider,value=\"aws\",include\n# +kubebuilder:validation:Enum=a;b\n"
           b"# +operator-builder:field:name=x,bogus=1,type=bool\n")
    res = po.parse(oracle.lex(doc), po.OPERATOR_BUILDER_REGISTRY)
    oks = [r for r in res if r[0] == "ok"]
    assert [r[1] for r in oks] == [b"+operator-builder:field", b"+operator-builder:field", b"+operator-builder:resource"]
    assert oks[0][2] == b"+operator-builder:field:name=webstore.replicas,default=2,type=int\n"
    assert oks[0][3] == [(b"name", "string", b"webstore.replicas"), (b"default", "int", b"2"), (b"type", "string", b"int")]
    assert oks[1][2] == b"+operator-builder:field:name=app.label,type=string,default=\"webstore\",description=`multi\n line`\n"
    assert oks[2][3][-1] == (b"include", "bool", b"true")
    check(oracle, doc)


def test_errors_and_unknown_markers(oracle):
    for doc in [b"# +operator-builder:field:name=x,default= true\n",       # ParseBool(" true") fails (state.go:113-117)
                b"# +operator-builder:field:name=x,default=1e39\n",          # float32 range (state.go:129-137)
                b"# +operator-builder:field:name=x,default=3.4028235e38\n",
                b"# +operator-builder:field:name='unterminated\n",         # lexer error -> error result
                b"# +operator-builder:field:name=a,,type=int\n# +operator-builder:field:name=b\n",
                b"# a+1 +operator-builder:field:name=x\n",                  # stale '+' poisons the marker name
                b"+operator-builder:resource:include\n", b"+docs: text\n+operator-builder:field:name=z", b""]:
        check(oracle, doc)


def test_fixtures_targeted_fuzz(oracle):
    for _p, doc in cu.fixtures():
        check(oracle, doc)
    for doc in cu.TARGETED + cu.NON_ASCII:
        check(oracle, doc)
    rng = random.Random(31)
    words = [b"+operator-builder:field:", b"+operator-builder:resource:", b"+operator-builder:collection:field:", b"name=", b"type=", b"default=",
             b"include", b"value=", b"field=", b",", b"\n# ", b"x", b"1", b"\"q\"", b"true", b" ", b"`a\n#b`", b"1.5", b"replace=", b"description="]
    for _ in range(3000):
        check(oracle, b"".join(rng.choice(words) for _ in range(rng.randint(1, 14))))
    for _ in range(300):
        check(oracle, cu.fuzz_doc_valid(rng))
    # the device walk handles the regular documents itself; only streams with pseudo-tuples go back to obm_parse_doc
    import operator_builder_b200 as ob
    before = dict(STATS)
    data0, _ = ob.generate_corpus_host(100, 4096, flavour=0)
    for i in range(100):
        check(oracle, data0.tobytes()[i * 4096:(i + 1) * 4096])
    assert STATS["host"] == before["host"], STATS          # synthetic manifests: all on the device walk
    for _p, doc in cu.fixtures():
        check(oracle, doc)
    assert STATS["host"] - before["host"] <= 8, STATS      # the reference's 33 fixtures: the ones with in-band warnings go back


@pytest.mark.gpu
def test_parser_on_gpu_tuples(oracle):
    """same check with tuples produced by the GPU through the C ABI"""
    import numpy as np
    import operator_builder_b200 as ob
    docs = [d for _p, d in cu.fixtures()] + list(cu.TARGETED)
    data0, off0 = ob.generate_corpus_host(200, 4096, flavour=1)
    docs += [data0.tobytes()[i * 4096:(i + 1) * 4096] for i in range(200)]
    data = np.frombuffer(b"".join(docs), dtype=np.uint8)
    off = np.zeros(len(docs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(d) for d in docs])
    sc = ob.Scanner(0)
    res = sc.lex_batch(data, off)
    n_ok = 0
    for i, doc in enumerate(docs):
        want = check(oracle, doc, res.tuples[int(res.doc_tuple_off[i]):int(res.doc_tuple_off[i + 1])])
        n_ok += want.count(b"+operator-builder:")
    assert n_ok > 1500
    sc.close()


@pytest.mark.gpu
def test_device_parser_kernel(oracle):
    """k_parse_docs through obm_parse_batch_device: records of every document == obm_parse_doc == the parser oracle"""
    import torch
    import operator_builder_b200 as ob
    from operator_builder_b200 import _native
    from oracle import parser_oracle as po
    rng = random.Random(78)
    docs = [d for _p, d in cu.fixtures()] + list(cu.TARGETED) + [cu.fuzz_doc_valid(rng) for _ in range(800)]
    docs += [b"# +operator-builder:field:name=x,default= true\n", b"# +operator-builder:field:name=x,default=1e39\n# +operator-builder:field:name=y\n",
             b"# +operator-builder:field:name=x,bogus=1\n# +operator-builder:resource:field=a,value=3.4028235e38,include\n"]
    for fl in (0, 1):
        data0, _ = ob.generate_corpus_host(300, 4096, flavour=fl)
        docs += [data0.tobytes()[i * 4096:(i + 1) * 4096] for i in range(300)]
    data = np.frombuffer(b"".join(docs), dtype=np.uint8)
    off = np.zeros(len(docs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(d) for d in docs])
    sc = ob.Scanner(0)
    res = sc.lex_batch(data, off)
    dev = torch.device("cuda:0")
    d_bytes = torch.from_numpy(data.copy()).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_tup = torch.from_numpy(res.tuples.view(np.int64).copy()).to(dev)
    d_toff = torch.from_numpy(res.doc_tuple_off.view(np.int64).copy()).to(dev)
    cap = len(res.tuples)
    d_res = torch.zeros(cap * 32, dtype=torch.uint8, device=dev)
    d_args = torch.zeros(cap * 16, dtype=torch.uint8, device=dev)
    d_roff = torch.zeros(len(docs) + 1, dtype=torch.int64, device=dev)
    d_tot = torch.zeros(2, dtype=torch.int64, device=dev)
    L = _native.lib()
    L.obm_parse_batch_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    reg = REG()
    st = torch.cuda.current_stream().cuda_stream
    rc = L.obm_parse_batch_device(sc.handle, reg.handle, d_bytes.data_ptr(), d_off.data_ptr(), len(docs), 1000, d_tup.data_ptr(), d_toff.data_ptr(),
                                  d_res.data_ptr(), cap, d_args.data_ptr(), cap, d_roff.data_ptr(), d_tot.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    roff = d_roff.cpu().numpy()
    nres, nargs = [int(x) for x in d_tot.cpu().tolist()]
    assert nres == int(roff[-1])
    R = d_res.cpu().numpy()[:nres * 32].view(RES_DT)
    A = d_args.cpu().numpy()[:nargs * 16].view(ARG_DT)
    host_docs = 0
    for i, doc in enumerate(docs):
        r = R[int(roff[i]):int(roff[i + 1])]
        assert all(int(x) == i + 1000 for x in r["doc"])
        host_docs += int(len(r) == 1 and (int(r[0]["flags"]) & R_HOST) != 0)
        t = res.tuples[int(res.doc_tuple_off[i]):int(res.doc_tuple_off[i + 1])]
        got = format_records(doc, t, r, A)
        want = po.serialize(po.parse(oracle.lex(doc), po.OPERATOR_BUILDER_REGISTRY))
        if got != want:  # which record differs from the host run of the same walk?
            L.obm_parse_doc_records.restype = ctypes.c_int64
            L.obm_parse_doc_records.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p,
                                                ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
            hr = np.zeros(len(t) + 2, dtype=RES_DT); ha = np.zeros(len(t) + 2, dtype=ARG_DT); na = ctypes.c_uint64()
            tt = np.ascontiguousarray(t, dtype=np.uint64)
            n = L.obm_parse_doc_records(reg.handle, doc, tt.ctypes.data, len(tt), i + 1000, hr.ctypes.data, len(hr), ha.ctypes.data, len(ha), ctypes.byref(na))
            dev_args = [A[int(x["arg_base"]):int(x["arg_base"]) + int(x["nargs"])].tolist() for x in r]
            host_args = [ha[int(x["arg_base"]):int(x["arg_base"]) + int(x["nargs"])].tolist() for x in hr[:n]]
            raise AssertionError(f"doc {i}: device records {r.tolist()} args {dev_args}\n host records {hr[:n].tolist()} args {host_args}\n roff {roff[i]}..{roff[i+1]}")
    assert host_docs < len(docs) * 0.75 and nres >= 4800, (host_docs, len(docs), nres)
    sc.close()


@pytest.mark.gpu
def test_marker_index_kernel(oracle):
    """k_marker_index: one record per marker whose definition the parser would load (definition.go:13-21)"""
    import ctypes
    import numpy as np
    import torch
    import operator_builder_b200 as ob
    from operator_builder_b200 import _native
    from oracle import parser_oracle as po
    rng = random.Random(77)
    docs = [d for _p, d in cu.fixtures()] + list(cu.TARGETED) + [cu.fuzz_doc_valid(rng) for _ in range(800)]
    docs += [b"# a+1 +operator-builder:field:name=x\n", b"++operator-builder:field:name=x", b"+operator-builder:fieldx:name=y\n+operator-builder:field:name=z",
             b"+operator-builder:field:\n+operator-builder:resource:include +operator-builder:collection:field:name=q"]
    data0, _ = ob.generate_corpus_host(300, 4096, flavour=1)
    docs += [data0.tobytes()[i * 4096:(i + 1) * 4096] for i in range(300)]
    data = np.frombuffer(b"".join(docs), dtype=np.uint8)
    off = np.zeros(len(docs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(d) for d in docs])
    sc = ob.Scanner(0)
    res = sc.lex_batch(data, off)
    dev = torch.device("cuda:0")
    d_bytes = torch.from_numpy(data.copy()).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_tup = torch.from_numpy(res.tuples.view(np.int64).copy()).to(dev)
    d_toff = torch.from_numpy(res.doc_tuple_off.view(np.int64).copy()).to(dev)
    cap = len(res.tuples)
    d_rec = torch.zeros(cap * 4, dtype=torch.int32, device=dev)
    d_roff = torch.zeros(len(docs) + 1, dtype=torch.int64, device=dev)
    L = _native.lib()
    reg = ob.Registry()
    st = torch.cuda.current_stream().cuda_stream
    rc = L.obm_marker_index_device(sc.handle, reg.handle, d_bytes.data_ptr(), d_off.data_ptr(), len(docs), d_tup.data_ptr(), d_toff.data_ptr(),
                                   d_rec.data_ptr(), cap, d_roff.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    roff = d_roff.cpu().numpy()
    rec = d_rec.cpu().numpy().view(np.uint32).reshape(-1, 4)
    names = [b"+operator-builder:field", b"+operator-builder:collection:field", b"+operator-builder:resource"]
    total = 0
    for i, doc in enumerate(docs):
        prs = po.Parser(oracle.lex(doc), po.OPERATOR_BUILDER_REGISTRY)
        prs.run()
        got = rec[int(roff[i]):int(roff[i + 1])]
        assert [names[int(r[3]) & 0xFFFF] for r in got] == prs.loaded, (doc[:200], got, prs.loaded)
        for r in got:
            assert int(r[0]) == i and doc[int(r[2])] == ord("+")
            t = int(res.tuples[int(res.doc_tuple_off[i]) + int(r[1])])
            assert t >> 59 == 2 and (t & 0xFFFFFFFF) == int(r[2])
        total += len(got)
    assert total == int(roff[-1]) and total > 2400
    sc.close()
