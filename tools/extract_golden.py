#!/usr/bin/env python3
"""Transcribe the reference's golden lexer vectors into tests/golden/lexer_golden.json.

Source: /root/reference/internal/markers/lexer/lexer_test.go:28-402 (24 table cases, each
`input -> [(Type, Value)]`; Pos is not asserted there, lexer_test.go:429-432).  Run in the build
container only (the reference tree is not present on the GPU box); the JSON output is committed.
Values are stored as latin-1-safe JSON strings (all vectors are ASCII).
"""
import json
import os
import re
import sys

REF = "/root/reference/internal/markers/lexer/lexer_test.go"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TYPES = ["LexemeError", "LexemeComment", "LexemeMarkerStart", "LexemeScope", "LexemeSeparator", "LexemeArg",
         "LexemeArgAssignment", "LexemeArgDelimiter", "LexemeStringLiteral", "LexemeFloatLiteral",
         "LexemeIntegerLiteral", "LexemeSyntheticBoolLiteral", "LexemeBoolLiteral", "LexemeQuote",
         "LexemeSliceBegin", "LexemeSliceEnd", "LexemeSliceDelimiter", "LexemeNakedSliceDelimiter",
         "LexemeMarkerEnd", "LexemeWarning", "LexemeEOF"]


def go_string_expr(src, i):
    """Evaluate a Go constant string expression (literals joined by +) starting at src[i]."""
    out = []
    n = len(src)
    while True:
        while src[i] in " \t\n":
            i += 1
        c = src[i]
        if c == '"':
            i += 1
            buf = []
            while src[i] != '"':
                if src[i] == "\\":
                    e = src[i + 1]
                    buf.append({"n": "\n", "t": "\t", '"': '"', "\\": "\\", "r": "\r", "'": "'"}[e])
                    i += 2
                else:
                    buf.append(src[i])
                    i += 1
            i += 1
            out.append("".join(buf))
        elif c == "`":
            j = src.index("`", i + 1)
            out.append(src[i + 1:j])
            i = j + 1
        else:
            raise ValueError(f"unexpected {c!r} at {i}")
        k = i
        while k < n and src[k] in " \t\n":
            k += 1
        if k < n and src[k] == "+":
            i = k + 1
            continue
        return "".join(out), i


def main():
    src = open(REF).read()
    body = src[src.index("tests := []struct"):src.index("focused := false")]
    cases = []
    for m in re.finditer(r"\bname:", body):
        name, i = go_string_expr(body, m.end())
        j = body.index("input:", i)
        inp, i = go_string_expr(body, j + len("input:"))
        k = body.index("expected:", i)
        end = body.find("\n\t\t},", k)
        exp_src = body[k:end]
        lexemes = []
        for lm in re.finditer(r"\{Type: lexer\.(\w+), Value: ", exp_src):
            val, _ = go_string_expr(exp_src, lm.end())
            lexemes.append([TYPES.index(lm.group(1)), val])
        line = src[:src.index("tests := []struct") + m.start()].count("\n") + 1
        cases.append({"name": name, "ref_line": line, "input": inp, "expected": lexemes})
    assert len(cases) == 24, len(cases)
    out = os.path.join(ROOT, "tests", "golden", "lexer_golden.json")
    with open(out, "w") as f:
        json.dump({"source": "internal/markers/lexer/lexer_test.go:28-402 @ 2827f233", "cases": cases}, f, indent=1)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    sys.exit(main())
