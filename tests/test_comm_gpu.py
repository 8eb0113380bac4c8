"""The multi-GPU path behind the C ABI (obm_comm_* / obm_lex_batch_sharded_device): file shards, one NCCL all-gather of
the compact index of registered markers.  nranks = 1 runs on the single-GPU test box; nranks = 2 needs two GPUs (two threads
of this process, one handle + one communicator each -- ncclCommInitRank meets across threads)."""
import ctypes
import random
import threading

import numpy as np
import pytest

from tests import corpus_util as cu

pytestmark = pytest.mark.gpu

REC_DT = np.dtype([("doc", "<u4"), ("tuple", "<u4"), ("off", "<u4"), ("reg_scopes", "<u4")])


def make_docs():
    import operator_builder_b200 as ob
    rng = random.Random(5)
    data0, _ = ob.generate_corpus_host(600, 4096)
    docs = [data0.tobytes()[i * 4096:(i + 1) * 4096] for i in range(600)] + [cu.fuzz_doc_valid(rng) for _ in range(300)]
    rng.shuffle(docs)
    return docs


def run_rank(rank, nranks, uid, docs, out, errors):
    try:
        import torch
        import operator_builder_b200 as ob
        from operator_builder_b200 import shard
        dev = torch.device("cuda", rank)
        torch.cuda.set_device(dev)
        sc = ob.Scanner(rank)
        comm = ob.Comm(sc, uid, rank, nranks)
        d0, d1 = shard.shard_range(len(docs), rank, nranks)
        mine = docs[d0:d1]
        data = np.frombuffer(b"".join(mine) + b"\0", dtype=np.uint8)[:-1]
        off = np.zeros(len(mine) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(d) for d in mine])
        nb = int(off[-1])
        d_bytes = torch.from_numpy(np.concatenate([data, np.zeros(64, np.uint8)])).to(dev)
        d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
        cap = nb + 2 * len(mine) + 64
        d_out = torch.zeros(cap, dtype=torch.int64, device=dev)
        d_toff = torch.zeros(len(mine) + 1, dtype=torch.int64, device=dev)
        d_status = torch.zeros(4, dtype=torch.int32, device=dev)
        d_counts = torch.zeros(2, dtype=torch.int64, device=dev)
        rcap = 16 * len(docs)
        d_idx = torch.zeros(rcap * 16, dtype=torch.uint8, device=dev)
        d_all = torch.zeros(nranks * rcap * 16, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        reg = ob.Registry()
        per_rank, stride = comm.lex_batch_sharded_device(reg, d_bytes.data_ptr(), d_off.data_ptr(), len(mine), nb, d0, d_out.data_ptr(), cap,
                                                         d_toff.data_ptr(), d_status.data_ptr(), d_counts.data_ptr(), d_idx.data_ptr(), rcap,
                                                         d_all.data_ptr(), nranks * rcap, st)
        torch.cuda.synchronize(dev)
        allr = d_all.cpu().numpy()
        slots = [allr[r * stride * 16:(r * stride + per_rank[r]) * 16].view(REC_DT).copy() for r in range(nranks)]
        own = d_idx.cpu().numpy()[:per_rank[rank] * 16].view(REC_DT).copy()
        out[rank] = (per_rank, stride, slots, own, int(d_toff[-1].item()))
        comm.close()
        sc.close()
    except Exception as e:  # noqa: BLE001
        errors.append((rank, repr(e)))


@pytest.mark.parametrize("nranks", [1, 2])
def test_sharded_step_gathers_every_ranks_results(nranks):
    import torch
    import operator_builder_b200 as ob
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    docs = make_docs()
    uid = ob.Comm.unique_id()
    out, errors = {}, []
    threads = [threading.Thread(target=run_rank, args=(r, nranks, uid, docs, out, errors)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert len(out) == nranks
    # every rank holds every rank's records, in global document order, with global document ids
    for r in range(nranks):
        per_rank, stride, slots, own, _ = out[r]
        assert per_rank == out[0][0] and stride == max(per_rank)
        for q in range(nranks):
            assert np.array_equal(slots[q], out[q][3]), (r, q)
    merged = np.concatenate(out[0][2])
    assert list(merged["doc"]) == sorted(merged["doc"]) and int(merged["doc"].max()) < len(docs)
    # ... and they are exactly the markers whose definition the parser would load (definition.go:13-21), per document
    import oracle
    from oracle import parser_oracle as po
    names = [b"+operator-builder:field", b"+operator-builder:collection:field", b"+operator-builder:resource"]
    k = 0
    for i, doc in enumerate(docs):
        prs = po.Parser(oracle.lex(doc), po.OPERATOR_BUILDER_REGISTRY)
        prs.run()
        n = len(prs.loaded)
        got = merged[k:k + n]
        assert len(got) == n and all(int(x) == i for x in got["doc"]), (i, doc[:100])
        assert [names[int(x) & 0xFFFF] for x in got["reg_scopes"]] == prs.loaded
        assert all(doc[int(o)] == ord("+") for o in got["off"])
        k += n
    assert k == len(merged)
