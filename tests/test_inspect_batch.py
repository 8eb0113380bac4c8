"""the batched per-node inspection (operator-builder_b200/inspect.py, the Python mirror of
go/inspect_batch.go + go/workload_batch.go) against the reference's flow -- one lexer + parser per YAML node over
Head + "\\n" + Line + "\\n" + Foot (inspect/yaml.go:89-95), manifest by manifest (kinds/workload.go:224-228).

The node trees come from tests/golden/node_comments.json (tools/make_node_comments.py): the reference's own 33 manifests
with a line-based stand-in for yaml.v3's comment attachment, which is a third-party dependency and unpinned -- the test
checks the batching (visiting order, one packed batch over every manifest, stream k to visit k), not yaml.v3."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "node_comments.json")))


def tree(d):
    from operator_builder_b200.inspect import Node
    kind = "mapping" if d["kind"] == "mapping" else d["kind"]
    content = [tree(c) for c in d["content"]] if "content" in d else None
    return Node(kind, content, d.get("head", ""), d.get("line", ""), d.get("foot", ""))


def manifests():
    return [[tree(doc) for doc in m["docs"]] for m in GOLD["manifests"]]


class HostScanner:
    """lex_batch through the host build of the product's lexer core (tests/hostsim): the CPU stand-in for Scanner"""

    def lex_batch(self, data, off):
        from operator_builder_b200 import BatchResult
        from tests import hostsim
        raw = bytes(data)
        parts, toff = [], [0]
        for i in range(len(off) - 1):
            t = hostsim.lex_doc(raw[int(off[i]):int(off[i + 1])])
            parts.append(t)
            toff.append(toff[-1] + len(t))
        tup = np.concatenate(parts) if parts else np.zeros(0, np.uint64)
        return BatchResult(tup, np.array(toff, dtype=np.uint64), {})


def reference_flow(oracle, mans):
    """what the reference does: per manifest, per visit, a fresh lexer + parser over the node's comment string"""
    from oracle import parser_oracle as po
    from operator_builder_b200 import inspect as ins
    out = []
    for m in mans:
        per = []
        ins.visit(m, lambda g, n: per.append((po.serialize(po.parse(oracle.lex(n.comment_input()), po.OPERATOR_BUILDER_REGISTRY)), g)))
        out.append(per)
    return out


def compare(got, want):
    assert len(got) == len(want)
    n_results = 0
    for gm, wm in zip(got, want):
        assert len(gm) == len(wm)
        for (gr, gg), (wr, wg) in zip(gm, wm):
            assert gr == wr
            assert len(gg) == len(wg) and all(a is b for a, b in zip(gg, wg))
            n_results += gr.count(b"+operator-builder:")
    return n_results


def test_batched_inspection_equals_per_node_flow(oracle):
    import operator_builder_b200 as ob
    from operator_builder_b200 import inspect as ins
    mans = manifests()
    got = ins.inspect_manifests_batched(HostScanner(), ob.Registry(), mans)
    assert compare(got, reference_flow(oracle, mans)) >= 60
    # the visiting order: a mapping's key and value share one group; every other node is its own
    m = [ins.Node("document", [ins.Node("mapping", [ins.Node(head="# +a:b"), ins.Node("mapping", [ins.Node(line="# k"), ins.Node(line="# v")]),
                                                   ins.Node(head="# k2"), ins.Node("sequence", [ins.Node(line="# item")])])])]
    order = []
    ins.visit(m, lambda g, n: order.append((len(g), n.head or n.line or n.kind)))
    assert order == [(1, "document"), (1, "mapping"), (2, "# +a:b"), (2, "mapping"), (2, "# k"), (2, "# v"), (2, "# k2"), (2, "sequence"), (1, "# item")]


@pytest.mark.gpu
def test_batched_inspection_on_gpu(oracle):
    import operator_builder_b200 as ob
    from operator_builder_b200 import inspect as ins
    mans = manifests()
    sc = ob.Scanner(0)
    try:
        got = ins.inspect_manifests_batched(sc, ob.Registry(), mans)
    finally:
        sc.close()
    assert compare(got, reference_flow(oracle, mans)) >= 60
