"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/obmarkers.h declares, and refuses to lex without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "obmarkers.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(obm_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from operator_builder_b200 import _native
    L = _native.lib()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/obmarkers.h but not exported by libobmarkers.so"
    assert sorted(_native.EXPORTS) == names
    assert L.obm_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import operator_builder_b200 as ob
    with pytest.raises(ob.NativeError) as ei:
        ob.Scanner(0)
    assert ei.value.code == -1  # OBM_E_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_oracle():
    """The product package must not import, link or call anything under oracle/ or tests/."""
    pkg = os.path.join(ROOT, "operator-builder_b200")
    for dirpath, _d, names in os.walk(pkg):
        for n in names:
            if n.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".go")):
                txt = open(os.path.join(dirpath, n), errors="replace").read()
                assert "liblexer_oracle" not in txt and "import oracle" not in txt and "hostsim" not in txt.replace("tests/hostsim", ""), n


def test_stream_mirror_of_next_lexeme():
    """obm_stream_* replays tuples like NewLexer/Run/NextLexeme, incl. the zero Lexeme after close
    (lexer.go:47,51-53).  Tuples here come from the host build of the core (test infrastructure)."""
    import operator_builder_b200 as ob
    from tests import hostsim
    doc = b"# +galaxy:planet=earth,flag\n"
    tup = hostsim.lex_doc(doc)
    lx = ob.Lexer(doc, tup)
    lx.run()
    got = []
    while True:
        lexeme = lx.next_lexeme()
        got.append((lexeme.type.name, lexeme.value))
        if lexeme.type == ob.LexemeType.EOF:
            break
    assert got == [("Comment", b"#"), ("MarkerStart", b"+"), ("Scope", b"galaxy"), ("Separator", b":"), ("Arg", b"planet"),
                   ("ArgAssignment", b"="), ("StringLiteral", b"earth"), ("ArgDelimiter", b","), ("Arg", b"flag"),
                   ("SyntheticBoolLiteral", b"true"), ("MarkerEnd", b"\n"), ("EOF", b"")]
    assert lx.next_lexeme() == ob.ZERO_LEXEME and lx.next_lexeme().type == ob.LexemeType.Error


def test_corpus_generator_shape(oracle):
    """BASELINE.md C2 generator: exactly 4,096 B per doc, ASCII, 8 markers, last byte newline, deterministic."""
    import operator_builder_b200 as ob
    data, off = ob.generate_corpus_host(200, 4096)
    data2, _ = ob.generate_corpus_host(100, 4096, first_doc=100)
    assert np.array_equal(data[100 * 4096:], data2)
    assert data.max() < 0x80
    lex_total = 0
    for d in range(200):
        doc = bytes(data[off[d]:off[d + 1]])
        assert len(doc) == 4096 and doc.endswith(b"\n")
        lx = oracle.lex(doc)
        assert sum(1 for t, *_ in lx if t == 2) == 8, d
        assert lx[-1][0] == 20
        lex_total += len(lx)
    assert 140 <= lex_total / 200 <= 175
    # collection flavour spells the collection markers (collection_field_marker.go:13)
    datac, _ = ob.generate_corpus_host(4, 4096, flavour=1)
    assert b"+operator-builder:collection:field:name=" in bytes(datac) and b"collectionField=" in bytes(datac)
    # odd sizes still come out exact
    for sz in (1, 2, 3, 5, 64, 199, 200, 777, 1536, 65536):
        dd, oo = ob.generate_corpus_host(3, sz)
        assert len(dd) == 3 * sz and bytes(dd)[-1:] == b"\n"
