#!/usr/bin/env python3
"""Freeze the reference's manifest fixtures (test/cases/**/.workloadConfig/**/*.yaml, 33 files,
36,010 B) into tests/golden/fixtures.json so that GPU-box tests do not need
/root/reference.  They are test inputs (data), not source code.  Run in the build container."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/test/cases"
files = []
paths = []
for dirpath, _dirs, names in os.walk(REF):  # .workloadConfig is a dot-directory: glob's ** skips it
    paths += [os.path.join(dirpath, n) for n in names if n.endswith(".yaml")]
for p in sorted(paths):
    files.append({"path": os.path.relpath(p, "/root/reference"), "content": open(p, encoding="utf-8").read()})
out = os.path.join(ROOT, "tests", "golden", "fixtures.json")
json.dump({"source": "test/cases/** @ 2827f233", "files": files}, open(out, "w"), indent=0)
print("wrote", out, len(files), "files", sum(len(f["content"].encode()) for f in files), "bytes")
