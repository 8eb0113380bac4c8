#!/usr/bin/env python3
"""tests/golden/fixtures.json (the reference's own manifests) -> tests/golden/node_comments.json: a node tree per manifest
with Head / Line / Foot comment strings, for the batched-inspection mirror test.

NOT yaml.v3's comment attachment (that is a third-party dependency absent from /root/reference, gopkg.in/yaml.v3
v3.0.0-20210107192922-496545a6307b, and unpinned by any reference test): a line-based stand-in that keeps the text of
every comment and hangs it on a plausible node -- a run of comment-only lines becomes the Head comment of the next key
(yaml.v3 keeps the leading '#'), a trailing comment the Line comment of that line's value.  The test only needs realistic
comment strings spread over a mapping / sequence tree; which node owns which comment does not matter for it."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def split_comment(line):
    """YAML-ish: a '#' at the line start or after white space starts a comment (quotes are not tracked: stand-in)"""
    m = re.search(r"(^|\s)#", line)
    if not m:
        return line, ""
    i = m.start() + (0 if m.group(1) == "" else 1)
    return line[:i].rstrip(), line[i:]


def build(text):
    docs, cur, pending = [], None, []

    def new_doc():
        return {"kind": "document", "content": [{"kind": "mapping", "content": []}]}

    for raw in text.split("\n"):
        if raw.strip() == "---":
            if cur is not None:
                docs.append(cur)
            cur, pending = new_doc(), []
            continue
        body, comment = split_comment(raw)
        if not body.strip():
            if comment:
                pending.append(comment)
            continue
        if cur is None:
            cur = new_doc()
        key = {"kind": "scalar", "head": "\n".join(pending), "line": "", "foot": ""}
        is_seq = body.lstrip().startswith("- ")
        val = {"kind": "sequence" if is_seq else "scalar", "head": "", "line": comment, "foot": ""}
        if is_seq:
            val["content"] = [{"kind": "scalar", "head": "", "line": "", "foot": ""}]
        cur["content"][0]["content"] += [key, val]
        pending = []
    if cur is not None:
        if pending:
            cur["content"][0]["foot"] = "\n".join(pending)
        docs.append(cur)
    return docs


def main():
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "fixtures.json")))
    out = {"note": "line-based stand-in for yaml.v3 comment attachment (unpinned); see tools/make_node_comments.py",
           "manifests": [{"path": e["path"], "docs": build(e["content"])} for e in fx["files"]]}
    path = os.path.join(ROOT, "tests", "golden", "node_comments.json")
    json.dump(out, open(path, "w"), indent=0, ensure_ascii=False)
    n = sum(len(m["docs"]) for m in out["manifests"])
    print(f"{len(out['manifests'])} manifests, {n} documents -> {path}")


if __name__ == "__main__":
    main()
