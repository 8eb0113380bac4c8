#!/usr/bin/env python3
"""SURVEY.md Appendix D (the survey model's derived vectors: Type "Value" @line:col) -> tests/golden/appendix_d.json.

These vectors were produced by an independent restatement written during the survey (it reproduces the reference's 24
golden vectors); no reference test pins them, so they cross-check the oracle's and the product's reading of the same Go
source -- including Pos -- rather than define ground truth.  `<strconv err>` in the survey stands for Go's strconv error
text; it is kept as a placeholder and matched as a wildcard by tests/test_appendix_d.py."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["Error", "Comment", "MarkerStart", "Scope", "Separator", "Arg", "ArgAssignment", "ArgDelimiter", "StringLiteral", "FloatLiteral",
         "IntegerLiteral", "SyntheticBoolLiteral", "BoolLiteral", "Quote", "SliceBegin", "SliceEnd", "SliceDelimiter", "NakedSliceDelimiter",
         "MarkerEnd", "Warning", "EOF"]


def unquote(s):
    """the survey writes values as Go-style double-quoted strings with \\n, \\" and \\\\ escapes"""
    assert s[0] == '"' and s[-1] == '"', s
    out, i, body = [], 0, s[1:-1]
    while i < len(body):
        c = body[i]
        if c == "\\":
            n = body[i + 1]
            out.append({"n": "\n", "t": "\t", '"': '"', "\\": "\\"}[n])
            i += 2
        else:
            out.append(c)
            i += 1
    return "".join(out)


def main():
    text = open(os.path.join(ROOT, "SURVEY.md"), encoding="utf-8").read()
    block = text[text.index("## Appendix D"):]
    block = block[block.index("```") + 3:]
    block = block[:block.index("```")]
    cases, cur = [], None
    for line in block.splitlines():
        if not line.strip():
            continue
        m = re.match(r'^in\s+(".*")\s*$', line)
        if m:
            cur = {"input": unquote(m.group(1)), "expected": []}
            cases.append(cur)
            continue
        m = re.match(r'^\s+(\w+)\s+(".*")\s+@(\d+):(\d+)\s*$', line)
        assert m and cur is not None, line
        cur["expected"].append([NAMES.index(m.group(1)), unquote(m.group(2)), int(m.group(3)), int(m.group(4))])
    out = {"source": "SURVEY.md Appendix D (survey commit 4a58431); unpinned by any reference test", "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "appendix_d.json")
    json.dump(out, open(path, "w"), indent=1, ensure_ascii=False)
    print(f"{len(cases)} cases, {sum(len(c['expected']) for c in cases)} lexemes -> {path}")


if __name__ == "__main__":
    main()
