"""SURVEY.md 8(f) rank 1 -- the parser token contract (8(a) a17) on the tuple stream.

obm_parse_doc (csrc/obm_parse.cpp, mirror of internal/markers/parser/state.go:13-175) over tuples vs the
Python restatement oracle/parser_oracle.py over the oracle lexer's lexemes.  No reference test pins the parser
(SURVEY section 4): both sides follow the source; this checks they agree and that MarkerText -- what
markers/markers.go:198-222 splices back into YAML comments -- is byte-exact."""
import random

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


def check(oracle, doc, tuples=None):
    from oracle import parser_oracle as po
    want = po.serialize(po.parse(oracle.lex(doc), po.OPERATOR_BUILDER_REGISTRY))
    got = results_via_tuples(doc, hostsim.lex_doc(doc) if tuples is None else tuples)
    assert got == want, (doc, got, want)
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
