"""Shared input generators for the parity tests (TEST INFRASTRUCTURE)."""
import glob
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))

# SURVEY.md Appendix D + targeted artefact cases (stale buffer, column drift, bufio window, ...)
TARGETED = [
    b"", b"\n", b"+", b"++", b"2+2=4", b"#", b"//", b"/", b"# ", b"+a", b"+a:", b"+a:b", b"+a:b=", b"+a:b=,", b"+a:b=1,",
    b"    #+docs: Defines the collection label\n", b"a+b: c\n", b"# +x:y= true\n", b"# +x:y=\n  false,z\n",
    b"# +x:a=1.17.3\n", b"# +x:a=1e+5\n", b"# +x:a=-5,b=.5,c=1e3,d=-\n", b"# +a:b=1 # +c:d\n",
    b"# +hello::x +p:q\n# +a:b\n", b"# ++hello:world", b"x: 1+1\n# +a:b\n", b"+hello\n# +a:b\n", b"+a:flag,other=1",
    b"+a:x=1,flag,other=2 tail", b"+a:x=1, y=2 # c", b"+a:x='q'z", b"+a:x=\"unterminated\n",
    b"# +a:d=`l1\n   # l2\n  // l3`,e\n", b"+a:b=trueish", b"+a:b+c:d", b"k: v # one # two +m:n\n",
    b"+a:b=true", b"+a:b=false,c", b"+a:b=  true\n", b"+a:b=\t\nfalse x", b"+a:b=tru", b"+a:b=1e309", b"+a:b=1e308",
    b"+a:b=1.7976931348623157e308", b"+a:b=1.7976931348623159e308", b"+a:b=0e999999", b"+a:b=-0.0", b"+a:b=1e-400",
    b"+a:b=99999999999999999999", b"+a:b=9223372036854775807", b"+a:b=9223372036854775808", b"+a:b=-9223372036854775808",
    b"+a:b=-9223372036854775809", b"+a:b=007", b"+a:b=1-2", b"+a:b=.", b"+a:b=..", b"+a:b=1.", b"+a:b=.e1", b"+a:b=1e",
    b"+a:b=1E5x", b"+a:b=5x", b"+a:b=5 x", b"+a:b=x;y", b"+a:b;c", b"+a:{", b"+a:b={", b"+a:b=(x)", b"+a:b=[1]",
    b"+a:b='it''s'", b"+a:b=''", b"+a:b=\"\"\n", b"+a:b=``", b"+a:b=`\n`", b"+a:b=`\n#`", b"+a:b=`\n//x`,c=1",
    b"+a:b=`x\n  y\n` trailing", b"+a:b=`unterminated\n# more\n", b"+a:b='x\ny'", b"+a:b=\"x+y\" +c:d=1",
    b"+a:b=1,,c", b"+a:b=1, c", b"+a:b=1,c,", b"+a:b,c", b"+a:b=c:d", b"+a:b=c=d", b"+a::b", b"+a:: b +c:d\n",
    b"+a:b=1+c:d", b"+a+b", b"+a +b:c", b"+1", b"+ a", b"+\n+a:b", b"x+\n# +a:b\n", b"+a\n+b\n+c:d", b"+a=1\n# x",
    b"# +a:b=c # +d:e=f // +g:h\n", b"//+a:b\n//+c:d", b"/ /+a:b", b"/+a:b", b"#\n#\n#+a:b", b"# a # b # c\n",
    b"key: 'a+b' # +x:y\n", b"data: aGVsbG8+d29ybGQ+Zm9v\n# +a:b=c\n", b"a: b\r\n# +c:d=e\r\n# +f:g\r\n",
    b"\t# +a:b\n\x0b\x0c# +c:d", b"+a:b=`" + b"z" * 100 + b"\n" + b" " * 50 + b"#cont`",
    b"+a:b=" + b" " * 4092 + b"true", b"+a:b=" + b" " * 4093 + b"true", b"+a:b=" + b" " * 4091 + b"false",
    b"+a:b=" + b" " * 4092 + b"false", b"+a:b=`x\n" + b" " * 4095 + b"#y`", b"+a:b=`x\n" + b" " * 4096 + b"#y`",
    b"+a:b=`x\n" + b" " * 4094 + b"//y`", b"+a:b=`x\n" + b" " * 4095 + b"//y`",
    b"+" + b"a" * 5000 + b":" + b"b" * 5000 + b"=" + b"c" * 9000 + b"\n",
]

NON_ASCII = [
    "# +\u00e9t\u00e9:nom=caf\u00e9\n".encode(), "+a:b=\u0663\n".encode(), "+a:b=1\u0663\n".encode(), "+\u4e2d:\u6587=\u5b57".encode(),
    b"\xff# +a:b\n", b"\xff\n# +a:b\n", b"+a:b=\xff\xfe\n", b"+a\xc3:b", b"+a:b='\xe2\x82'", b"# \xe2\x82\xac +a:b\n",
    "+a:b=\u00a0true".encode(), "+a:b=\u3000\u3000false,c".encode(), "+a:b=`x\n\u2003# y`".encode(), b"+\xc3\xa9:x",
    "x: 1+1\n\u00e9+\n# +a:b".encode(), b"+a:b=\xf0\x9f\x98\x80", b"\xed\xa0\x80+a:b", b"\xc0\x80+a:b\n", b"+a:b\xff",
    "+a:\u00bd=\u00bd".encode(), "+a:b=\u00b2".encode(), b"\x80\x80\x80#\n+a:b", b"+a:b=`\n\xff#`", b"#\xf4\x90\x80\x80 +a:b",
]


def fixtures():
    """The reference's own manifest fixtures, frozen into tests/golden/fixtures.json by tools/freeze_fixtures.py."""
    path = os.path.join(HERE, "golden", "fixtures.json")
    if not os.path.exists(path):
        return []
    return [(e["path"], e["content"].encode("utf-8")) for e in json.load(open(path))["files"]]


ALPHABET = b"+#/:=, \n\"'`;{}[]()abtruefls0159.-eE\t\r_x"


def fuzz_doc(rng: random.Random, max_len=200, non_ascii=False):
    n = rng.randint(0, max_len)
    out = bytearray()
    while len(out) < n:
        k = rng.random()
        if k < 0.55:
            out.append(rng.choice(ALPHABET))
        elif k < 0.70:
            out += rng.choice([b"+a:", b"+ab:c=", b"# ", b"// ", b"true", b"false", b"=`", b"\n#", b",x=", b"='", b"=\"", b"+x:y:z="])
        elif k < 0.80:
            out += rng.choice([b"\n", b"\n  ", b" \n", b"\n# +"])
        elif non_ascii and k < 0.90:
            out += rng.choice([b"\xff", b"\xc3\xa9", b"\xe2\x82\xac", b"\xc2\xa0", b"\xe3\x80\x80", b"\xd9\xa3", b"\xe2\x82", b"\x80",
                               b"\xf0\x9f\x98\x80", b"\xc2\x85", b"\xed\xa0\x80"])
        else:
            out += bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyzABC0123456789") for _ in range(rng.randint(1, 6)))
    return bytes(out)


def _ident(rng):
    parts = []
    for _ in range(rng.randint(1, 4)):
        parts.append("".join(rng.choice("abcdefghijklmnopqrstuvwxyzABCXYZ0123456789_-") for _ in range(rng.randint(1, 14))))
    return ".".join(parts)


def _value(rng):
    k = rng.random()
    if k < 0.25:
        return _ident(rng)
    if k < 0.45:
        q = rng.choice("\"'`")
        body = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz0123456789 .:/+-,;=#{}[]()") for _ in range(rng.randint(0, 40)))
        return q + body.replace(q, "") + q
    if k < 0.65:
        return str(rng.randint(-500, 100000))
    if k < 0.78:
        return rng.choice(["1.5", "0.25", "-3.75e2", ".5", "1e3", "2.", "007"])
    if k < 0.97:
        return rng.choice(["true", "false", " true", "  false"])
    return rng.choice(["nginx:1.17", "a;b;c", "x=y", "{a}", "", "v1,", "1e309", "`multi\n  # line`", "`a\n//b\nc`"])


def fuzz_doc_valid(rng: random.Random, max_lines=30):
    """Mostly well-formed manifests: YAML-ish lines, comments, markers with 1-3 scopes and 0-5 arguments;
    a few malformed bits sprinkled in so that neighbouring regular lines are exercised too."""
    lines = []
    for _ in range(rng.randint(0, max_lines)):
        k = rng.random()
        indent = " " * rng.choice([0, 2, 4, 6, 8])
        if k < 0.35:
            lines.append(f"{indent}{_ident(rng)}: {rng.choice(['value', '3', '\"q\"', 'a/b', 'it' + chr(39) + 's', 'x+y', ''])}")
        elif k < 0.45:
            lines.append(f"{indent}# {rng.choice(['plain comment', 'see https://x/y', 'a # b', 'c // d', ''])}")
        elif k < 0.5:
            lines.append(rng.choice(["", "---", "  ", "\t", "// go comment", "/ not", "a//b: c"]))
        else:
            scopes = ":".join(_ident(rng).replace(".", "") or "s" for _ in range(rng.randint(1, 3)))
            args = []
            for _ in range(rng.randint(0, 5)):
                a = _ident(rng).replace(".", "")
                args.append(a if rng.random() < 0.2 else f"{a}={_value(rng)}")
            m = "+" + scopes + (":" + ",".join(args) if args else "")
            pre = rng.choice(["# ", "#", "// ", f"{_ident(rng)}: v  # ", "", "#   "])
            post = rng.choice(["", "", "", "", " trailing words", " # +second:marker=1", " +x:y", "  "])
            lines.append(indent + pre + m + post)
    tail = rng.choice(["\n", "", "\n\n"])
    return ("\n".join(lines) + tail).encode()
