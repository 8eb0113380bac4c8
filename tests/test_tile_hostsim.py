"""CTA emulation of the tile fast path (obm_tile.h phase functions, tests/hostsim) vs the exact path:
the two tuple streams must be identical, for every batch shape the kernel distinguishes
(sub-batches, skewed starts, large documents, documents with non-ASCII bytes, interacting lines)."""
import json
import os
import random

import numpy as np
import pytest

from tests import corpus_util as cu
from tests import hostsim


def check_batch(docs, skew=0, which=1):
    """every device organisation -- r01's fused tile kernel (0) and two-stage pipeline (1), and the fused warp kernel of
    mode 0 (2; the device source itself, run by 32 fibers per warp) -- against the exact path.  Returns the stats of
    organisation `which`."""
    keep = None
    for pipeline in (0, 1, 2):
        tup, toff, stats = hostsim.tile_batch(docs, skew, pipeline=pipeline)
        for i, doc in enumerate(docs):
            want = hostsim.lex_doc(doc)
            got = tup[int(toff[i]):int(toff[i + 1])]
            if not np.array_equal(got, want):
                raise AssertionError(f"pipeline={pipeline} doc {i} (skew {skew}) {doc[:300]!r}\n tile ={hostsim.fmt_tuples(got)[:50]}\n exact={hostsim.fmt_tuples(want)[:50]}")
        assert int(toff[-1]) == len(tup)
        # marker / lexeme totals are recomputed from the stream: they must agree for every organisation
        kinds = (tup >> np.uint64(59)).astype(np.int64)
        assert int(stats[0]) == int((kinds == 2).sum()), (pipeline, stats)
        assert int(stats[1]) == int(((kinds < 21) | (kinds > 25)).sum()), (pipeline, stats)
        if pipeline == which:
            keep = stats
    return keep


def test_dense_marker_lines_long_lines_and_large_documents():
    """shapes that stress the ordered pipeline's group logic: more marker lines than the staging area holds,
    lines with more tuples than a staging slot, large documents between small ones, interacting lines and
    non-ASCII documents inside dense groups, tiles with more than DMAX documents"""
    long_line = b"# +operator-builder:field:" + b",".join(b"a%d=%d" % (i, i) for i in range(40)) + b"\n"
    dense = b"".join(b"# +a:b:c=%d\n" % i for i in range(700))
    big = b"kind: X\n" + b"".join(b"  key%d: v # +operator-builder:field:name=k%d,type=string\n" % (i, i) for i in range(600))
    assert len(big) > 16368
    multi = b"# +a:b:c=`x\ny`\n# +d:e\n"
    docs = [dense, long_line * 3, big, b"", multi, dense[:3000] + "é".encode() + dense[3000:6000], long_line, big + long_line, b"#x\n" * 2000]
    docs += [b"# +t:%d\n" % i for i in range(300)]  # > DMAX documents in one tile
    docs += [dense, multi * 40, b"+" * 50 + b"\n"]
    for skew in (0, 9):
        check_batch(docs, skew)
    check_batch([long_line * 200])  # every line overflows its staging slot
    # tokens, quoted strings and further markers beyond the 128 bytes the stepper's delimiter mask covers
    far = b"# +a:b:c=\"" + b"x y,z" * 40 + b"\" +d:e:f=" + b"v" * 150 + b" +g:h +i:j:k='" + b"q;" * 70 + b"',l=1.5\n"
    for lead in range(0, 33, 5):
        check_batch([b"x" * lead + b"\n" + far * 3, far[:140] + b"\n" + far, far[:129] + b"\n", far[:127]])
    check_batch([big])  # a batch that is one large document
    check_batch([b""] * 200 + [big] + [b""] * 3)


def test_units_with_more_owning_lines_than_the_owner_table():
    """lines of a few bytes: a unit's documents together have more owning lines than OWN_CAP (256); the fused warp
    kernel then scans and writes the unit one document at a time instead of handing every document to the exact lexer"""
    short = [b"".join(b"# c%d\n" % (i % 7) for i in range(n)) for n in (100, 120, 90, 130, 40, 200)]
    marks = [b"".join(b"# +a:b:c=%d\n" % (i % 10) if i % 3 == 0 else b"# x\n" for i in range(n)) for n in (150, 110, 170)]
    odd = [b"# +a:b=`x\ny`\n" + b"# k\n" * 120, b"# k\n" * 100 + "é # +a:b\n".encode() + b"# k\n" * 30, b"", b"#\n" * 255, b"#\n" * 257, b"+x:y\n" * 140]
    for skew in (0, 5):
        st = check_batch(short * 3, skew, which=2)
        assert int(st[2]) == 0, st  # no document took the exact lexer
        st = check_batch(marks * 3 + short, skew, which=2)
        assert int(st[2]) == 0, st
        check_batch(short[:2] + odd + marks + short[2:] + odd[::-1], skew)
    # the adversarial sweep's 16-byte lines: 4 KiB documents of 256 lines each
    cell = [(b"# " + b"x" * 13 + b"\n") * 256, (b"# +s:a0=0 +s:a\n") * 256, (b"#  +q:v=\"yyyy\"\n") * 256]
    for doc in cell:
        st = check_batch([doc] * 7, 0, which=2)
        assert int(st[2]) == 0, st


def test_non_ascii_neighbour_in_the_same_aligned_word():
    """a document whose last bytes are >= 0x80 followed, inside the same aligned 4-byte word, by an ASCII document that
    starts with a marker: the word-wise skipping of the line lexers must not let the neighbour's bytes hide the '+'"""
    for tail in (b"\xff", b"\xc3\xa9", b"\xf0\x9f\x98\x80", b"\x80\xff\xfe"):
        for pad in range(0, 9):
            docs = [b"x" * pad + b"k: v\n" + tail, b"+lerhacA\n9='", b"k" * pad + tail, b"+a:b=\"q\",c\n", b"y" * pad + tail, b"# +a:b:c='x',d=1\n"]
            for skew in (0, 1, 2, 3, 6):
                check_batch(docs, skew)


def test_targeted_as_one_batch():
    docs = list(cu.TARGETED)
    for skew in (0, 1, 7, 15):
        check_batch(docs, skew)


GO_UNICODE_SPACE = set([0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000]) | set(range(0x2000, 0x200B))


def needs_sequential_lexer(doc: bytes) -> bool:
    """invalid UTF-8 (utf8.DecodeRune) or unicode.IsSpace beyond ASCII: the two things that let a line without tuples
    reach across a newline (csrc/obm_tile.h utf8_plain)"""
    try:
        text = doc.decode("utf-8")
    except UnicodeDecodeError:
        return True
    return any(ord(ch) in GO_UNICODE_SPACE for ch in text)


def test_non_ascii_documents():
    """invalid UTF-8 / Unicode white space -> sequential Unicode lexer for the document; other non-ASCII text stays on the
    line-parallel path (only its non-ASCII lines take the Unicode lexer); r01's fused CTA kernel sends both away"""
    docs = list(cu.NON_ASCII) + [b"# +a:b\n"] * 5
    stats = check_batch(docs)  # stats of the pipeline emulation
    must = sum(needs_sequential_lexer(d) for d in docs)
    assert must >= 8 and must <= int(stats[2]) < len(cu.NON_ASCII)
    # valid text beyond ASCII inside otherwise regular documents, at every alignment of the 32-byte classification words
    base = b"k: v  # +operator-builder:field:name=a,type=string\n" * 3
    uni = ["é", "中文", "😀", "ß=ü", "# +ключ:значение=да", "x: 'naïve'  # +s:a=\"ö\",b"]
    many = []
    for u in uni:
        for pad in range(0, 40, 3):
            many.append(base + b" " * pad + u.encode() + b"\n" + base)
            if "=" not in u:  # appended to a naked string value / inside a comment in front of a marker
                many.append(base[:-1] + u.encode() + b"\n" + b"# " + u.encode() + b" +s:t=1\n")
    stats = check_batch(many, skew=5)
    assert int(stats[2]) == 0
    # the fused warp kernel (mode 0): valid text beyond ASCII keeps a document line-parallel unless it sits on a MARKER line
    # (names, values and the letter after '+' are judged by Unicode classes there); comments and plain YAML do not care
    def marker_line_with_text_beyond_ascii(doc):
        return any(b"+" in ln and any(b >= 0x80 for b in ln) for ln in doc.split(b"\n"))
    stats = check_batch(many, skew=3, which=2)
    assert int(stats[2]) <= sum(marker_line_with_text_beyond_ascii(d) for d in many) < len(many)
    plain = []
    for u in ["é", "中文 # plain", "# café au lait", "// ünïcödé", "k: 'ß'  # größe", "😀😀😀 # x"]:
        for pad in range(0, 40, 3):
            plain.append(base + b" " * pad + u.encode() + b"\n" + base + b"# " + u.encode() * 3 + b"\n")
    stats = check_batch(plain, skew=3, which=2)
    assert int(stats[2]) == 0
    stats = check_batch(docs + many + plain, skew=7, which=2)
    assert must <= int(stats[2])


def test_fixtures_and_golden():
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lexer_golden.json")))
    check_batch([c["input"].encode() for c in golden["cases"]])
    check_batch([d for _p, d in cu.fixtures()], skew=5)


def test_synthetic_corpus():
    import operator_builder_b200 as ob
    for flavour in (0, 1):
        data, off = ob.generate_corpus_host(64, 4096, flavour=flavour)
        raw = data.tobytes()
        stats = check_batch([raw[i * 4096:(i + 1) * 4096] for i in range(64)])
        assert stats[0] == 8 * 64 and stats[2] == 0
    for doc_bytes, n in ((37, 300), (1000, 50), (13296, 3), (13297, 3), (16368, 3), (16369, 3), (70000, 2)):
        data, off = ob.generate_corpus_host(n, doc_bytes)
        raw = data.tobytes()
        check_batch([raw[i * doc_bytes:(i + 1) * doc_bytes] for i in range(n)], skew=3)


def test_many_tiny_and_empty_documents():
    check_batch([b""] * 200)
    check_batch([b"", b"+a:b", b"", b"", b"#", b"\n"] * 60)
    check_batch([b"#\n" * 1500])           # more special lines than the owner queue holds -> exact
    check_batch([b"+a:b\n" * 400, b"x" * 20000, b"# +c:d=1\n" * 100, b""])


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_batches(seed):
    rng = random.Random(4200 + seed)
    for _ in range(40):
        docs = [cu.fuzz_doc(rng, max_len=rng.choice([5, 60, 400, 3000]), non_ascii=rng.random() < 0.15)
                for _ in range(rng.randint(1, 80))]
        check_batch(docs, skew=rng.randint(0, 15))


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_valid_batches(seed):
    """well-formed manifests: almost every document stays on the line-parallel path"""
    rng = random.Random(7700 + seed)
    exact = total = 0
    for _ in range(25):
        docs = [cu.fuzz_doc_valid(rng) for _ in range(rng.randint(1, 40))]
        stats = check_batch(docs, skew=rng.randint(0, 15))
        exact += int(stats[2]); total += len(docs)
    assert exact < total * 0.6


def test_staged_line_views_lookahead_across_the_newline():
    """K2 lexes marker lines from a shared-memory copy that ends W_LOOK bytes after the newline (obm_pipe.h LineView).
    Whitespace-skipping peeks (peek.go:65-89) may look across the newline: every distance around the end of the copy,
    every peeked token, and line tails that leave the lexer in each value / argument state."""
    tails = [b"+a:b:c=", b"+a:b:c= ", b"+a:b:c", b"+a:b:c,", b"+a:b:c=1,", b"+a:b:c=1,d", b"+a:b:c=1,d=", b"+a:b", b"+a:b:", b"+a",
             b"+a:b:c=\"x", b"+a:b:c=`x", b"+a:b:c='x", b"+a:b:c=tru", b"+a:b:c=t", b"# +a:b:c=", b"  x: y # +a:b:c=", b"+a:b:c=1 ", b"+a:b:c=1\t"]
    nexts = [b"true", b"false", b"//x", b"#y", b"x", b"truex", b"tru", b"=1", b",d=2", b"`", b"\"", b""]
    docs = []
    for tail in tails:
        for nxt in nexts:
            for ws in (0, 1, 2, 7, 15, 16, 17, 18, 19, 20, 23, 24, 25, 31, 40, 70):
                for sep in (b" ", b"\t", b"\n"):
                    docs.append(b"k: v\n" + tail + b"\n" + sep * ws + nxt + b"\nrest: 1 # +z:y=2\n")
    random.Random(7).shuffle(docs)
    for lo in range(0, len(docs), 400):
        check_batch(docs[lo:lo + 400], skew=lo % 16)
    # the same at the very end of a document / of the batch (the copy is cut at the document end)
    check_batch([t + b"\n" + b" " * w for t in tails for w in (0, 1, 5, 30)] + [tails[0]])


def test_utf8_check_matches_go_semantics():
    """K1's per-document check (valid per utf8.DecodeRune, no unicode.IsSpace beyond ASCII) against Python's strict decoder
    plus the white-space set: random byte soup biased towards lead / continuation bytes, every alignment"""
    L = hostsim.lib()
    rng = random.Random(31337)
    pieces = [bytes([b]) for b in (0x80, 0xBF, 0xC0, 0xC1, 0xC2, 0xDF, 0xE0, 0xE1, 0xEC, 0xED, 0xEE, 0xEF, 0xF0, 0xF1, 0xF3, 0xF4, 0xF5, 0xFF,
                                   0x85, 0xA0, 0x9A, 0x9F, 0x90, 0x8F, 0xA8, 0xAF, 0x81, 0x8A, 0x8B)]
    pieces += ["é".encode(), "€".encode(), "😀".encode(), "\u00a0".encode(), "\u0085".encode(), "\u1680".encode(), "\u2003".encode(),
               "\u200a".encode(), "\u200b".encode(), "\u2028".encode(), "\u202f".encode(), "\u205f".encode(), "\u3000".encode(), "\u3001".encode(),
               "\ud7ff".encode(), "\ue000".encode(), "\U0010ffff".encode(), b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xe0\x9f\xbf", b"\xf0\x8f\xbf\xbf",
               b"a", b"\n", b" ", b"#", b"+x:y=1"]
    valid = [p for p in pieces if not needs_sequential_lexer(p)]
    n_plain = 0
    for it in range(20000):
        doc = b"".join(rng.choice(valid if it % 2 and rng.random() < 0.9 else pieces) for _ in range(rng.randint(1, 12)))
        if rng.random() < 0.3:
            doc = b"k: v\n" * rng.randint(0, 20) + doc + b"\nz" * rng.randint(0, 40)
        got = L.hs_utf8_plain(doc, len(doc), rng.randint(0, 15))
        want = 0 if needs_sequential_lexer(doc) else 1
        assert got == want, (doc, got, want)
        n_plain += want
    assert 2000 < n_plain < 18000
