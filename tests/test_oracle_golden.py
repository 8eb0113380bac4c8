"""Pins the CPU oracle against the reference's own tests.

* 24 golden vectors: internal/markers/lexer/lexer_test.go:28-402 (transcribed by
  tools/extract_golden.py into tests/golden/lexer_golden.json); compared on (Type, Value) exactly
  as lexer_test.go:421-440 does, reading lexemes until EOF.
* primitive expectations: consume_internal_test.go:24-37,71-114,152-191,223-234,268-287 and
  peek_internal_test.go:23-42,69-100,135-161,189-228.
"""
import ctypes
import json
import os

import pytest

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lexer_golden.json")))


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=[c["name"] for c in GOLDEN["cases"]])
def test_golden_vector(oracle, case):
    got = [[t, v.decode("utf-8")] for (t, v, _l, _c) in oracle.lex(case["input"].encode())]
    # the reference loop stops at the first EOF lexeme (lexer_test.go:435)
    assert got[-1][0] == 20
    assert got == case["expected"]


def _prim(oracle, s):
    L = oracle.lib()
    return L, L.obo_prim_new(s.encode(), len(s.encode()))


def _state(L, h):
    buf = ctypes.create_string_buffer(4096)
    n = L.obo_prim_buffer(h, buf, 4096)
    line, col = ctypes.c_int64(), ctypes.c_int64()
    L.obo_prim_pos(h, ctypes.byref(line), ctypes.byref(col))
    return buf.raw[:n].decode(), (line.value, col.value)


def _cstrs(xs):
    arr = (ctypes.c_char_p * max(1, len(xs)))(*[x.encode() for x in xs])
    return arr, len(xs)


# consume_internal_test.go:24-37
@pytest.mark.parametrize("inp,s,buf,pos", [
    ("Hello World", "Hello", "Hello", (1, 6)),
    ("Hello \nWorld", "Hello \nWorld", "Hello \nWorld", (2, 6)),
])
def test_prim_consume(oracle, inp, s, buf, pos):
    L, h = _prim(oracle, inp)
    L.obo_prim_consume(h, s.encode())
    assert _state(L, h) == (buf, pos)
    L.obo_prim_free(h)


# consume_internal_test.go:71-114
@pytest.mark.parametrize("inp,tok,exc,want,buf,pos", [
    ("Hello World", "Hello", [], True, "Hello", (1, 6)),
    ("HelloWorld", "Hello", ["W"], False, "", (1, 1)),
    ("Hello World", "GoodBye", [], False, "", (1, 1)),
    ("Hello \nWorld", "Hello \nWorld", [], True, "Hello \nWorld", (2, 6)),
])
def test_prim_consumed(oracle, inp, tok, exc, want, buf, pos):
    L, h = _prim(oracle, inp)
    arr, n = _cstrs(exc)
    assert bool(L.obo_prim_consumed(h, tok.encode(), arr, n)) == want
    assert _state(L, h) == (buf, pos)
    L.obo_prim_free(h)


# consume_internal_test.go:152-191
@pytest.mark.parametrize("inp,toks,want,buf,pos", [
    ("Hello World", ["Hello"], True, "Hello", (1, 6)),
    ("    Hello World", ["Hello", "World"], True, "    Hello", (1, 10)),
    ("Hello World", ["GoodBye"], False, "", (1, 1)),
    ("   \nWorld", ["World"], True, "   \nWorld", (2, 6)),
])
def test_prim_consumedWhitespaced(oracle, inp, toks, want, buf, pos):
    L, h = _prim(oracle, inp)
    arr, n = _cstrs(toks)
    assert bool(L.obo_prim_consumedWhitespaced(h, arr, n)) == want
    assert _state(L, h) == (buf, pos)
    L.obo_prim_free(h)


# consume_internal_test.go:223-234
@pytest.mark.parametrize("inp,buf,pos", [
    ("   \n\tHello World", "   \n\t", (2, 2)),
    ("Hello World", "", (1, 1)),
])
def test_prim_consumeWhitespace(oracle, inp, buf, pos):
    L, h = _prim(oracle, inp)
    L.obo_prim_consumeWhitespace(h)
    assert _state(L, h) == (buf, pos)
    L.obo_prim_free(h)


# consume_internal_test.go:268-287
@pytest.mark.parametrize("inp,exc,want,buf,pos", [
    ("Hello+World", ["\n", "+"], True, "Hello", (1, 6)),
    ("Hello World", ["H"], False, "", (1, 1)),
])
def test_prim_consumeUntil(oracle, inp, exc, want, buf, pos):
    L, h = _prim(oracle, inp)
    arr = (ctypes.c_int32 * len(exc))(*[ord(c) for c in exc])
    assert bool(L.obo_prim_consumeUntil(h, arr, len(exc))) == want
    assert _state(L, h) == (buf, pos)
    L.obo_prim_free(h)


# peek_internal_test.go:23-42
@pytest.mark.parametrize("inp,want", [("Hello World", ord("H")), ("H", ord("H")), ("\n", ord("\n")), ("", -1)])
def test_prim_peek(oracle, inp, want):
    L, h = _prim(oracle, inp)
    assert L.obo_prim_peek(h) == want
    L.obo_prim_free(h)


# peek_internal_test.go:69-100
@pytest.mark.parametrize("inp,n,want", [
    ("Hello World", 2, [ord("H"), ord("e")]),
    ("H", 2, [ord("H"), -1]),
    ("H\n", 2, [ord("H"), ord("\n")]),
    ("", 2, [-1]),
])
def test_prim_peekN(oracle, inp, n, want):
    L, h = _prim(oracle, inp)
    out = (ctypes.c_int32 * 8)()
    c = L.obo_prim_peekN(h, n, out)
    assert list(out[:c]) == want
    L.obo_prim_free(h)


# peek_internal_test.go:135-161
@pytest.mark.parametrize("inp,tok,exc,want", [
    ("Hello World", "Hello", [], True),
    ("HelloWorld", "Hello", ["W"], False),
    ("HelloWorld", "Goodbye", [], False),
])
def test_prim_peeked(oracle, inp, tok, exc, want):
    L, h = _prim(oracle, inp)
    arr, n = _cstrs(exc)
    assert bool(L.obo_prim_peeked(h, tok.encode(), arr, n)) == want
    L.obo_prim_free(h)


# peek_internal_test.go:189-228
@pytest.mark.parametrize("inp,toks,want", [
    ("  Hello World", ["Hello"], True),
    ("HelloWorld", ["Hello"], True),
    ("HelloWorld", ["Goodbye"], False),
    ("    ", ["Hello"], False),
    ("    \nHello", ["Hello"], True),
])
def test_prim_peekedWhitespaced(oracle, inp, toks, want):
    L, h = _prim(oracle, inp)
    arr, n = _cstrs(toks)
    assert bool(L.obo_prim_peekedWhitespaced(h, arr, n)) == want
    L.obo_prim_free(h)
