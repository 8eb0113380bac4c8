"""N>1 host logic on CPU: world_size-2 gloo run of the file sharding + the index exchange step.

Each rank builds ITS shard of the synthetic corpus by global document index (the same generator the GPU kernel
runs, host twin obm_generate_corpus_host), gets per-document tuple counts from the CPU checker (tests/hostsim
replays the device code on the host; this is test infrastructure, not a product path), exchanges counts and
checks the global index against a single-process run over the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NDOCS, DOC_BYTES = 37, 1024  # odd on purpose: uneven shards


def _counts(d0, d1):
    import operator_builder_b200 as ob
    from tests import hostsim
    data, off = ob.generate_corpus_host(d1 - d0, DOC_BYTES, d0, 0)
    docs = [bytes(data[off[i]:off[i + 1]]) for i in range(d1 - d0)]
    return docs, torch.tensor([len(hostsim.lex_doc(d)) for d in docs], dtype=torch.int32)


def _worker(rank, world, port, ndocs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from operator_builder_b200 import shard
        d0, d1 = shard.shard_range(ndocs, rank, world)
        docs, counts = _counts(d0, d1)
        allc = shard.exchange_counts(counts, ndocs, rank, world)
        rank_of, local_off = shard.global_index(allc, ndocs, world)
        q.put((rank, d0, d1, allc.tolist(), rank_of.tolist(), local_off.tolist(), [hash(d) for d in docs] and [len(d) for d in docs]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ndocs", [NDOCS, 36])
def test_shard_and_index_exchange_world2(ndocs):
    from operator_builder_b200 import shard
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, world, port, ndocs, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    _, ref_counts = _counts(0, ndocs)
    ref = ref_counts.tolist()
    covered = []
    for rank, d0, d1, allc, rank_of, local_off, lens in res:
        assert allc == ref  # every rank holds the same global counts, equal to the single-process run
        assert all(l == DOC_BYTES for l in lens)
        covered += list(range(d0, d1))
        for d in range(ndocs):
            assert rank_of[d] == shard.owner_of(d, ndocs, world)
            e0 = shard.shard_range(ndocs, rank_of[d], world)[0]
            assert local_off[d] == sum(ref[e0:d])
    assert covered == list(range(ndocs))  # shards are disjoint and complete


def test_shard_range_properties():
    from operator_builder_b200 import shard
    for n in (0, 1, 7, 8, 1000, 2621440):
        for w in (1, 2, 3, 4, 8):
            rs = [shard.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
            for d in {0, n // 2, n - 1} - {-1}:
                if 0 <= d < n:
                    a, b = shard.shard_range(n, shard.owner_of(d, n, w), w)
                    assert a <= d < b
