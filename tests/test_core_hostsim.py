"""Host build of the product's lexer core (obm_core.h) + decoder (obm_decode.cpp) vs the oracle.

decode(core(doc)) must equal oracle(doc) on (Type, Value, Pos) for every lexeme -- stricter than
the reference's own test, which ignores Pos (lexer_test.go:429-432).  Also checks that composing a
document from per-line lexers (what the fast kernel does) yields the identical tuple stream.
"""
import json
import os
import random

import numpy as np
import pytest

from tests import corpus_util as cu
from tests import hostsim

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lexer_golden.json")))


def check(oracle, doc, by_lines_too=True):
    want = oracle.lex_raw(doc)
    tup = hostsim.lex_doc(doc)
    got = hostsim.decode(doc, tup)
    if got != want:
        raise AssertionError(f"doc={doc!r}\n tuples={hostsim.fmt_tuples(tup)}\n got ={oracle.parse_stream(got)}\n want={oracle.parse_stream(want)}")
    if by_lines_too:
        tl = hostsim.lex_doc(doc, by_lines=True)
        if not np.array_equal(tl, tup):
            raise AssertionError(f"by-lines stream differs doc={doc!r}\n lines={hostsim.fmt_tuples(tl)}\n doc  ={hostsim.fmt_tuples(tup)}")
    return tup


def test_golden_vectors(oracle):
    for c in GOLDEN["cases"]:
        check(oracle, c["input"].encode())


def test_targeted(oracle):
    for doc in cu.TARGETED:
        check(oracle, doc)


def test_non_ascii(oracle):
    for doc in cu.NON_ASCII:
        check(oracle, doc, by_lines_too=False)


def test_reference_fixtures(oracle):
    fx = cu.fixtures()
    assert len(fx) == 33
    total = 0
    for _path, doc in fx:
        tup = check(oracle, doc)
        total += int(np.sum((tup >> np.uint64(59)) == 2))
    assert total > 70  # 74 operator-builder markers + kubebuilder/docs ones


def test_c1_standalone_sample(oracle):
    """1,906 B, survey model: 139 lexemes, 7 markers, first Comment at {7 16}."""
    doc = dict(cu.fixtures())["test/cases/standalone/.workloadConfig/resources.yaml"]
    assert len(doc) == 1906
    lx = oracle.lex(doc)
    assert len(lx) == 139
    assert sum(1 for t, *_ in lx if t == 2) == 7 and sum(1 for t, *_ in lx if t == 18) == 7
    assert lx[0][0] == 1 and (lx[0][2], lx[0][3]) == (7, 16)
    check(oracle, doc)


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_ascii(oracle, seed):
    rng = random.Random(1000 + seed)
    for _ in range(1500):
        check(oracle, cu.fuzz_doc(rng))


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_non_ascii(oracle, seed):
    rng = random.Random(5000 + seed)
    for _ in range(1500):
        check(oracle, cu.fuzz_doc(rng, non_ascii=True), by_lines_too=False)


def test_strconv_helpers_agree(oracle):
    L, H = oracle.lib(), hostsim.lib()
    rng = random.Random(7)
    cases = [b"1e309", b"1e308", b"1.7976931348623157e308", b"1.7976931348623158e308", b"1.79769313486231580793728971405303e308",
             b"179769313486231580793728971405303415079934132710037826936173778980444968292764750946649017977587207096330286416692887910946555547851940402630657488671505820681908902000708383676273854845817711531764475730270069855571366959622842914819860834936475292719074168444365510704342711559699508093042880177904174497791",
             b"0." + b"0" * 400 + b"1e709", b"0e99999", b"1e10000", b"1e100000", b".", b"-", b"-.", b"1.", b".5", b"-.5e-3", b"1e", b"1e-", b"--1", b"1-1"]
    for _ in range(3000):
        cases.append(bytes(rng.choice(b"0123456789.eE-") for _ in range(rng.randint(1, 12))))
    for s in cases:
        assert L.obo_parse_float_err(s, len(s)) == H.hs_parse_float_err(s, len(s)), s
        assert L.obo_atoi_err(s, len(s)) == H.hs_atoi_err(s, len(s)), s
        # cross-check the syntax/range classes with Python where the grammars coincide
        txt = s.decode()
        try:
            v = float(txt)
            py = 2 if v in (float("inf"), float("-inf")) else 0
        except ValueError:
            py = 1
        if "_" not in txt and "n" not in txt.lower() and "i" not in txt.lower():
            assert L.obo_parse_float_err(s, len(s)) == py, (s, py)


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_valid(oracle, seed):
    rng = random.Random(8800 + seed)
    for _ in range(400):
        check(oracle, cu.fuzz_doc_valid(rng))


def test_large_documents_chunk_parallel_path():
    """obm_large.h: chunks of LCHUNK bytes starting at line boundaries, lexed independently, chain-checked; the result must
    equal the sequential stream whether the chain validates or the document falls back"""
    import random
    from tests import corpus_util as cu
    rng = random.Random(99)
    pool = [b"key: value\n", b"  - item  # +operator-builder:field:name=a.b,type=int,default=3\n", b"\n", b"path: /a/b+c\n",
            b"x" * 300 + b"\n", b"# plain comment\n", b"a: 'q'  # +x:y=\"s t\",z\n", b"1+1\n", b"++x:y\n", b"# +noscope\n"]
    validated = 0
    for trial in range(12):
        parts, n = [], 0
        target = rng.choice([5000, 17000, 40000, 120000])
        multi = trial % 3 == 2  # sprinkle constructs that cross lines (and sometimes chunk boundaries)
        while n < target:
            ln = rng.choice(pool + ([b"# +x:y=`multi\n  # line`\n", b"# +a:b=\n   true\n"] if multi else []))
            parts.append(ln); n += len(ln)
        doc = b"".join(parts)
        if trial % 4 == 1:
            doc = doc[:-1]  # no trailing newline
        got, ok = hostsim.large_doc(doc)
        assert np.array_equal(got, hostsim.lex_doc(doc)), trial
        validated += ok
    assert validated >= 6
    # adversarial: a back-tick literal swallowing a chunk boundary, a fatal error early, a single line longer than many chunks
    for doc in (b"k: v\n" * 700 + b"# +x:y=`" + b"z\n" * 3000 + b"`\n" + b"k: v\n" * 100,
                b"# +a:b=\"unterminated\n" + b"k: v # +c:d=1\n" * 2000,
                b"# +a:b=" + b"v" * 30000 + b"\nk: v # +c:d=1\n" * 10,
                b"", b"\n", b"x" * 4096, b"x" * 4095 + b"\n" + b"# +a:b\n", "# é +a:b=ü\n".encode() * 900):
        got, ok = hostsim.large_doc(doc)
        assert np.array_equal(got, hostsim.lex_doc(doc))


def test_chunk_path_ascii_fast_lexer_on_fuzz():
    """the chunk threads run the ASCII instantiation (word-wise skipping, token loop) line by line from the line START,
    not from the first special byte as the tile path does: any document, any size, must still equal the sequential stream"""
    import random
    from tests import corpus_util as cu
    rng = random.Random(1234)
    docs = list(cu.TARGETED) + [d for _p, d in cu.fixtures()]
    docs += [cu.fuzz_doc(rng, max_len=rng.choice([5, 60, 400, 3000, 9000])) for _ in range(500)]
    docs += [cu.fuzz_doc_valid(rng) for _ in range(200)]
    docs += [b"".join(cu.fuzz_doc(rng, max_len=700) for _ in range(30)) for _ in range(20)]
    for doc in docs:
        got, _ok = hostsim.large_doc(doc)
        assert np.array_equal(got, hostsim.lex_doc(doc)), doc[:300]


def test_decoder_survives_arbitrary_tuple_streams():
    """obm_decode_doc is handed device output by callers of the C ABI: whatever the 64-bit words are (kinds, offsets and
    lengths far outside the document), it must clamp and return, never read outside the document"""
    import random
    rng = random.Random(5)
    L = hostsim.lib()
    for _ in range(5000):
        doc = bytes(rng.randrange(256) for _ in range(rng.randint(0, 60)))
        tup = np.array([(rng.randrange(32) << 59) | ((rng.randrange(1 << 27) if rng.random() < 0.2 else rng.randint(0, 70)) << 32)
                        | (rng.randrange(1 << 32) if rng.random() < 0.2 else rng.randint(0, 70)) for _ in range(rng.randint(0, 12))], dtype=np.uint64)
        assert isinstance(hostsim.decode(doc, tup, L), bytes)
