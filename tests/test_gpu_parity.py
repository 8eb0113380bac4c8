"""GPU parity tests proper: everything goes through the C ABI (obm_lex_batch / obm_lex_batch_device)
and is compared lexeme-for-lexeme -- (Type, Value, Pos), stricter than lexer_test.go:429-432 --
with the CPU oracle on the same inputs.  Bit-exact: this is integer/byte/index work."""
import json
import os
import random

import numpy as np
import pytest

from tests import corpus_util as cu

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lexer_golden.json")))


@pytest.fixture(scope="module")
def scanner():
    import operator_builder_b200 as ob
    sc = ob.Scanner(0)
    yield sc
    sc.close()


def pack(docs):
    data = np.frombuffer(b"".join(docs) + b"\0", dtype=np.uint8)[:-1] if docs else np.zeros(0, np.uint8)
    off = np.zeros(len(docs) + 1, dtype=np.uint64)
    if docs:
        off[1:] = np.cumsum([len(d) for d in docs])
    return data, off


def run_and_compare(scanner, oracle, docs, modes=(0, 1, 2, 3)):
    import operator_builder_b200 as ob
    data, off = pack(docs)
    want_stream, want_off, want_n = oracle.lex_batch_raw(data if len(data) else np.zeros(1, np.uint8), off)
    streams = []
    for mode in modes:
        scanner.set_mode(mode)
        res = scanner.lex_batch(data, off)
        streams.append(res.tuples.copy())
        n_lex = 0
        for i, doc in enumerate(docs):
            t = res.tuples[int(res.doc_tuple_off[i]):int(res.doc_tuple_off[i + 1])]
            got = ob.decode_doc_raw(doc, t)
            want = want_stream[int(want_off[i]):int(want_off[i + 1])]
            if got != want:
                from tests import hostsim
                raise AssertionError(f"mode {mode} doc {i} {doc[:200]!r}\n tuples={hostsim.fmt_tuples(t)[:60]}\n"
                                     f" got ={oracle.parse_stream(got)[:40]}\n want={oracle.parse_stream(want)[:40]}")
            n_lex += int(want_n[i])
        assert res.stats["n_lexemes"] == n_lex
        assert res.stats["n_tuples"] == len(res.tuples)
    scanner.set_mode(0)
    for s in streams[1:]:
        assert np.array_equal(streams[0], s), "fast path and exact path tuple streams differ"
    return streams[0]


def test_golden_vectors(scanner, oracle):
    """the reference's 24 vectors, lexer_test.go:28-402, as one batch"""
    docs = [c["input"].encode() for c in GOLDEN["cases"]]
    run_and_compare(scanner, oracle, docs)
    # and with the reference test's own comparison: (Type, Value) until EOF
    import operator_builder_b200 as ob
    data, off = pack(docs)
    res = scanner.lex_batch(data, off)
    for c, lx in zip(GOLDEN["cases"], scanner.lexers(data.tobytes(), off, res)):
        lx.run()
        got = []
        while True:
            lexeme = lx.next_lexeme()
            got.append([int(lexeme.type), lexeme.value.decode()])
            if lexeme.type == ob.LexemeType.EOF:
                break
        assert got == c["expected"], c["name"]


def test_targeted_and_edge_cases(scanner, oracle):
    run_and_compare(scanner, oracle, list(cu.TARGETED))


def test_non_ascii_and_invalid_utf8(scanner, oracle):
    run_and_compare(scanner, oracle, list(cu.NON_ASCII))


def test_reference_fixtures(scanner, oracle):
    fx = cu.fixtures()
    assert len(fx) == 33
    run_and_compare(scanner, oracle, [d for _p, d in fx])


def test_empty_and_ragged_batches(scanner, oracle):
    run_and_compare(scanner, oracle, [])
    run_and_compare(scanner, oracle, [b""])
    run_and_compare(scanner, oracle, [b"", b"", b"+a:b", b"", b"#", b""])
    rng = random.Random(3)
    docs = [cu.fuzz_doc(rng, max_len=rng.choice([0, 1, 7, 63, 64, 65, 500, 5000])) for _ in range(300)]
    run_and_compare(scanner, oracle, docs)


@pytest.mark.parametrize("seed", range(4))
def test_fuzz(scanner, oracle, seed):
    rng = random.Random(900 + seed)
    docs = [cu.fuzz_doc(rng, max_len=400, non_ascii=(seed % 2 == 1)) for _ in range(4000)]
    run_and_compare(scanner, oracle, docs)


@pytest.mark.parametrize("seed", range(3))
def test_fuzz_valid_manifests(scanner, oracle, seed):
    rng = random.Random(600 + seed)
    docs = [cu.fuzz_doc_valid(rng) for _ in range(3000)]
    run_and_compare(scanner, oracle, docs)


def test_c2_10k_docs_bit_exact(scanner, oracle):
    """BASELINE.json configs[1]: 10k synthetic 4 KiB manifests (40,960,000 B), 8 markers/file."""
    import operator_builder_b200 as ob
    data, off = ob.generate_corpus_host(10000, 4096)
    want_stream, want_off, want_n = oracle.lex_batch_raw(data, off)
    res = scanner.lex_batch(data, off)
    assert res.stats["n_markers"] == 80000
    assert res.stats["n_lexemes"] == int(want_n.sum())
    raw = data.tobytes()
    for i in range(10000):
        doc = raw[i * 4096:(i + 1) * 4096]
        got = ob.decode_doc_raw(doc, res.tuples[int(res.doc_tuple_off[i]):int(res.doc_tuple_off[i + 1])])
        assert got == want_stream[int(want_off[i]):int(want_off[i + 1])], i
    for mode in (1, 2, 3):
        scanner.set_mode(mode)
        res1 = scanner.lex_batch(data, off)
        assert np.array_equal(res.tuples, res1.tuples) and np.array_equal(res.doc_tuple_off, res1.doc_tuple_off), mode
    scanner.set_mode(0)


def test_chunked_overlapped_host_path(scanner, oracle):
    """obm_lex_batch pipelines big host batches in chunks (H2D / scan / D2H overlapped): same result as one shot"""
    import operator_builder_b200 as ob
    data, off = ob.generate_corpus_host(3000, 4096)
    rng = random.Random(5)
    extra = [cu.fuzz_doc_valid(rng) for _ in range(500)] + [b"", b"+a:b=1e309\n# +c:d", "# +\u00e9:x=1\n".encode()]
    docs = [data.tobytes()[i * 4096:(i + 1) * 4096] for i in range(3000)] + extra
    pdata, poff = pack(docs)
    one = scanner.lex_batch(pdata, poff)
    old = scanner.set_chunk_bytes(1 << 20)
    try:
        for mode in (0, 1, 3):
            scanner.set_mode(mode)
            res = scanner.lex_batch(pdata, poff)
            assert np.array_equal(res.tuples, one.tuples) and np.array_equal(res.doc_tuple_off, one.doc_tuple_off), mode
            assert res.stats["n_markers"] == one.stats["n_markers"] and res.stats["n_lexemes"] == one.stats["n_lexemes"]
    finally:
        scanner.set_mode(0)
        scanner.set_chunk_bytes(old)
    run_and_compare(scanner, oracle, docs[-600:], modes=(0,))


def test_collection_flavour_and_odd_doc_sizes(scanner, oracle):
    import operator_builder_b200 as ob
    for doc_bytes, n in ((4096, 500), (1000, 700), (37, 100), (16384, 40), (70000, 6)):
        data, off = ob.generate_corpus_host(n, doc_bytes, flavour=1)
        raw = data.tobytes()
        run_and_compare(scanner, oracle, [raw[i * doc_bytes:(i + 1) * doc_bytes] for i in range(n)])


def test_adversarial_sweep(scanner, oracle):
    """BASELINE.json configs[4] (scaled to test size): markers/line 0..64 x line length 16 B..64 KiB."""
    docs = []
    for m in (0, 1, 2, 4, 8, 16, 32, 64):
        for L in (16, 64, 256, 1024, 4096, 16384, 65536):
            line = b"# " + b" ".join(b"+s:a%d=%d" % (k, k) for k in range(m))
            if len(line) + 1 > L:
                line = line[:L - 1]
            pad = L - 1 - len(line)
            variants = [line + b" " + b"x" * (pad - 1) if pad > 0 else line,
                        line + (b" +q:v=\"" + b"y" * (pad - 9) + b"\"" if pad > 9 else b" " * pad),
                        line + (b" +q:v=" + b"z" * (pad - 6) if pad > 6 else b" " * pad)]
            for v in variants:
                docs.append((v + b"\n") * max(1, min(8, 131072 // L)))
    run_and_compare(scanner, oracle, docs)


def test_large_documents(scanner, oracle):
    """documents larger than any tile: 1 MiB and 9 MiB, many lines, markers sprinkled"""
    rng = random.Random(11)
    big = []
    for target in (1 << 20, 9 << 20):
        parts, n = [], 0
        while n < target:
            ln = rng.choice([b"key: value\n", b"  - item  # +operator-builder:field:name=a.b,type=int,default=3\n", b"\n",
                             b"# +x:y=`multi\n  # line`\n", b"path: /a/b+c\n", b"x" * 300 + b"\n"])
            parts.append(ln)
            n += len(ln)
        big.append(b"".join(parts))
    run_and_compare(scanner, oracle, big + [b"+a:b"] + big[:1])


def test_large_documents_chunk_parallel(scanner, oracle):
    """documents above the tile size take the chunk-parallel exact path (csrc/obm_large.h): regular documents validate
    their chunk chain, a back-tick literal across a chunk boundary / a fatal error force the sequential fallback, single
    lines longer than several chunks leave empty chunks; non-ASCII text; small documents in between"""
    rng = random.Random(23)
    pool = [b"key: value\n", b"  - item  # +operator-builder:field:name=a.b,type=int,default=3\n", b"\n", b"path: /a/b+c\n",
            b"x" * 300 + b"\n", b"# plain comment\n", b"a: 'q'  # +x:y=\"s t\",z\n", b"1+1\n", b"++x:y\n", b"# +noscope\n"]
    docs = []
    for target in (17000, 40000, 300000, 2 << 20):
        parts, n = [], 0
        while n < target:
            ln = rng.choice(pool); parts.append(ln); n += len(ln)
        docs.append(b"".join(parts))
        docs.append(b"k: v # +s:a=1\n")
    docs.append(docs[2][:-1])                                                   # no trailing newline
    docs.append(b"k: v\n" * 700 + b"# +x:y=`" + b"z\n" * 3000 + b"`\n" + b"k: v\n" * 100)   # crosses chunk boundaries -> fallback
    docs.append(b"# +a:b=\"unterminated\n" + b"k: v # +c:d=1\n" * 2000)                      # fatal early -> fallback
    docs.append(b"# +a:b=" + b"v" * 30000 + b"\n" + b"k: v # +c:d=1\n" * 10)                   # one line over 7 chunks
    docs.append("# é +a:b=ü\n".encode() * 2500)
    docs.append(b"x" * 16369)
    res = run_and_compare(scanner, oracle, docs)


def test_valid_utf8_documents_stay_line_parallel(scanner, oracle):
    """valid UTF-8 beyond ASCII without Unicode white space: only the lines with such bytes take the Unicode lexer
    (k2_units<true>); invalid UTF-8 / Unicode white space send the document to the sequential lexer; same stream either way"""
    base = b"k: v  # +operator-builder:field:name=a,type=string\n" * 3
    uni = ["é", "中文", "😀", "ß=ü", "# +ключ:значение=да", "x: 'naïve'  # +s:a=\"ö\",b"]
    docs = []
    for u in uni:
        for pad in range(0, 40, 3):
            docs.append(base + b" " * pad + u.encode() + b"\n" + base)
            if "=" not in u:
                docs.append(base[:-1] + u.encode() + b"\n" + b"# " + u.encode() + b" +s:t=1\n")
    import operator_builder_b200 as ob
    data0, _ = ob.generate_corpus_host(64, 4096, 0, 0)
    raw = data0.tobytes()
    for i in range(64):
        d = raw[i * 4096:(i + 1) * 4096]
        docs.append(d.replace(b"plain", "plén".encode(), 1) if i % 3 else d)
    docs += list(cu.NON_ASCII) + ["# caf\u00e9\u00a0+a:b\n".encode(), b"k: v\n\xff\n# +a:b\n"]
    run_and_compare(scanner, oracle, docs)
    # valid UTF-8 without Unicode white space: r01's pipeline (mode 3) keeps such documents line-parallel (only the lines
    # with bytes >= 0x80 take the Unicode lexer); the fused warp kernel (mode 0) keeps them line-parallel unless the text
    # sits on a marker line (Unicode letter / number classes matter there): those documents it lexes sequentially
    plain = docs[:len(docs) - len(cu.NON_ASCII) - 2]
    data, off = pack(plain)
    scanner.set_mode(3)
    assert scanner.lex_batch(data, off).stats["n_docs_exact"] == 0
    scanner.set_mode(0)
    on_marker_line = sum(any(b"+" in ln and any(b >= 0x80 for b in ln) for ln in d.split(b"\n")) for d in plain)
    assert 0 < on_marker_line < sum(any(b >= 0x80 for b in d) for d in plain)
    assert scanner.lex_batch(data, off).stats["n_docs_exact"] == on_marker_line


def test_device_entry_point_and_capacity(scanner, oracle):
    import torch
    import operator_builder_b200 as ob
    ndocs, doc_bytes = 4096, 4096
    dev = torch.device("cuda:0")
    d_bytes = torch.empty(ndocs * doc_bytes, dtype=torch.uint8, device=dev)
    d_off = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
    scanner.generate_corpus_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, doc_bytes, first_doc=0, flavour=0,
                                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host, hoff = ob.generate_corpus_host(ndocs, doc_bytes)
    assert np.array_equal(d_bytes.cpu().numpy(), host), "device and host generators must agree byte for byte"
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), hoff)
    cap = ndocs * doc_bytes // 8
    d_out = torch.zeros(cap, dtype=torch.int64, device=dev)
    d_toff = torch.zeros(ndocs + 1, dtype=torch.int64, device=dev)
    d_status = torch.zeros(4, dtype=torch.int32, device=dev)
    d_counts = torch.zeros(2, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    scanner.lex_batch_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, ndocs * doc_bytes, d_out.data_ptr(), cap,
                             d_toff.data_ptr(), d_status.data_ptr(), d_counts.data_ptr(), st)
    torch.cuda.synchronize()
    ref = scanner.lex_batch(host, hoff)
    n = int(d_toff[-1].item())
    assert n == len(ref.tuples) and int(d_status[0].item()) == 0
    assert np.array_equal(d_out[:n].cpu().numpy().view(np.uint64), ref.tuples)
    assert np.array_equal(d_toff.cpu().numpy().view(np.uint64), ref.doc_tuple_off)
    assert int(d_counts[0].item()) == 8 * ndocs
    # capacity overflow is reported, never written past
    small = 1000
    d_out2 = torch.full((small + 64,), -1, dtype=torch.int64, device=dev)
    scanner.lex_batch_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, ndocs * doc_bytes, d_out2.data_ptr(), small,
                             d_toff.data_ptr(), d_status.data_ptr(), d_counts.data_ptr(), st)
    torch.cuda.synchronize()
    assert int(d_status[0].item()) == 1 and int(d_toff[-1].item()) == n
    assert bool((d_out2[small:] == -1).all())
    # host entry point: OBM_E_CAPACITY with the needed count
    import ctypes
    from operator_builder_b200 import _native
    L = _native.lib()
    cnt = ctypes.c_uint64()
    toff = np.zeros(ndocs + 1, dtype=np.uint64)
    tiny = np.zeros(10, dtype=np.uint64)
    rc = L.obm_lex_batch(scanner.handle, host.ctypes.data, hoff.ctypes.data, ndocs, tiny.ctypes.data, 10, ctypes.byref(cnt),
                         toff.ctypes.data, None)
    assert rc == _native.OBM_E_CAPACITY and cnt.value == n and not tiny.any()


def test_full_size_properties_on_device(scanner):
    """Size-independent properties at a size the oracle cannot check in seconds (1 GiB resident):
    fast path == exact path tuple-for-tuple; per-document counts are a function of the document only
    (a shard generated at a different global index range reproduces the same tuples); totals add up."""
    import torch
    dev = torch.device("cuda:0")
    ndocs, doc_bytes = 262144, 4096
    st = torch.cuda.current_stream().cuda_stream
    d_bytes = torch.empty(ndocs * doc_bytes, dtype=torch.uint8, device=dev)
    d_off = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
    scanner.generate_corpus_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, doc_bytes, 0, 1, st)
    cap = ndocs * doc_bytes // 8
    outs = []
    for mode in (0, 1, 2, 3):
        scanner.set_mode(mode)
        d_out = torch.zeros(cap, dtype=torch.int64, device=dev)
        d_toff = torch.zeros(ndocs + 1, dtype=torch.int64, device=dev)
        d_status = torch.zeros(4, dtype=torch.int32, device=dev)
        d_counts = torch.zeros(2, dtype=torch.int64, device=dev)
        scanner.lex_batch_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, ndocs * doc_bytes, d_out.data_ptr(), cap,
                                 d_toff.data_ptr(), d_status.data_ptr(), d_counts.data_ptr(), st)
        torch.cuda.synchronize()
        assert int(d_status[0].item()) == 0
        assert int(d_counts[0].item()) == 8 * ndocs
        outs.append((d_out, d_toff, int(d_toff[-1].item()), int(d_counts[1].item())))
    scanner.set_mode(0)
    for k in (1, 2, 3):
        assert outs[0][2] == outs[k][2] and outs[0][3] == outs[k][3]
        assert torch.equal(outs[0][1], outs[k][1]) and torch.equal(outs[0][0], outs[k][0])
    # shard property: documents [100000, 100000+4096) regenerated alone give the same tuples
    sub = 4096
    s_bytes = torch.empty(sub * doc_bytes, dtype=torch.uint8, device=dev)
    s_off = torch.empty(sub + 1, dtype=torch.int64, device=dev)
    scanner.generate_corpus_device(s_bytes.data_ptr(), s_off.data_ptr(), sub, doc_bytes, 100000, 1, st)
    s_out = torch.zeros(sub * doc_bytes // 8, dtype=torch.int64, device=dev)
    s_toff = torch.zeros(sub + 1, dtype=torch.int64, device=dev)
    scanner.lex_batch_device(s_bytes.data_ptr(), s_off.data_ptr(), sub, sub * doc_bytes, s_out.data_ptr(), len(s_out),
                             s_toff.data_ptr(), None, None, st)
    torch.cuda.synchronize()
    full_out, full_toff = outs[0][0], outs[0][1]
    a, b = int(full_toff[100000].item()), int(full_toff[100000 + sub].item())
    assert b - a == int(s_toff[-1].item())
    assert torch.equal(full_out[a:b], s_out[:b - a])


@pytest.mark.parametrize("ndocs,flavour,name", [(1048576, 1, "C3: workload-collection spelling, 4 GiB"), (2621440, 0, "C4: 10 GiB")])
def test_full_size_corpus_per_document_hash(scanner, oracle, ndocs, flavour, name):
    """BASELINE.md section 2 at the benched sizes: EVERY document of the resident batch, through a 64-bit hash of its decoded
    lexeme stream (Type, Pos, Value).  GPU side: obm_hash_batch_device over the tuples of the fused warp kernel's scan;
    CPU side: the oracle's own per-document hash (oracle.scan_batch(want_doc_hash=True)) over the same bytes copied back."""
    import ctypes
    import torch
    from operator_builder_b200 import _native
    dev = torch.device("cuda:0")
    doc_bytes = 4096
    free, _total = torch.cuda.mem_get_info()
    if free < ndocs * doc_bytes * 1.8 + (2 << 30):
        pytest.skip("not enough free HBM for " + name)
    st = torch.cuda.current_stream().cuda_stream
    d_bytes = torch.empty(ndocs * doc_bytes, dtype=torch.uint8, device=dev)
    d_off = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
    scanner.generate_corpus_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, doc_bytes, 0, flavour, st)
    cap = ndocs * doc_bytes // 16
    d_out = torch.empty(cap, dtype=torch.int64, device=dev)
    d_toff = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
    d_status = torch.zeros(4, dtype=torch.int32, device=dev)
    d_counts = torch.zeros(2, dtype=torch.int64, device=dev)
    scanner.set_mode(0)
    scanner.lex_batch_device(d_bytes.data_ptr(), d_off.data_ptr(), ndocs, ndocs * doc_bytes, d_out.data_ptr(), cap, d_toff.data_ptr(),
                             d_status.data_ptr(), d_counts.data_ptr(), st)
    d_hash = torch.zeros(ndocs, dtype=torch.int64, device=dev)
    d_nhost = torch.zeros(1, dtype=torch.int32, device=dev)
    L = _native.lib()
    rc = L.obm_hash_batch_device(scanner.handle, d_bytes.data_ptr(), d_off.data_ptr(), ndocs, d_out.data_ptr(), d_toff.data_ptr(), d_hash.data_ptr(),
                                 d_nhost.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    assert int(d_status[0].item()) == 0 and int(d_counts[0].item()) == 8 * ndocs
    assert int(d_nhost.item()) == 0  # synthetic manifests: every document is hashed on the device
    got = d_hash.cpu().numpy().view(np.uint64)
    host = d_bytes.cpu().numpy()
    off = d_off.cpu().numpy().astype(np.uint64)
    del d_bytes, d_out
    torch.cuda.empty_cache()
    import os
    threads = max(1, len(os.sched_getaffinity(0)))
    n_lex, n_mark, _h, want = oracle.scan_batch(host, off, threads, want_doc_hash=True)
    assert n_mark == 8 * ndocs and n_lex == int(d_counts[1].item())
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, (name, len(bad), bad[:10])
