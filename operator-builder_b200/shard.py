"""File sharding of a manifest batch over ranks (one process per GPU) and the one exchange step of the path.

Documents are independent (reference: internal/markers/lexer/lexer.go:27-40 builds one lexer per input; the
driver loop internal/workload/v1/kinds/workload.go:224-285 visits manifests one by one), so the batch is cut
into contiguous ranges of whole files and no data-path collective is needed for the scan itself.  What every
rank needs afterwards is the GLOBAL doc_tuple_off index (where each document's tuples live and on which
rank): an all-gather of 4 bytes per document.  Tuples stay resident on the rank that produced them.

torch.distributed is plumbing here: "nccl" on the GPU box, "gloo" in the CPU tests (tests/test_shard_gloo.py).
"""
import torch
import torch.distributed as dist


def shard_range(ndocs: int, rank: int, world: int):
    """Contiguous range [d0, d1) of documents owned by `rank`; ranges differ by at most one document."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return ndocs * rank // world, ndocs * (rank + 1) // world


def owner_of(doc: int, ndocs: int, world: int) -> int:
    """Rank whose shard_range contains `doc` (inverse of shard_range)."""
    if not (0 <= doc < ndocs):
        raise ValueError("doc out of range")
    r = min(world - 1, (doc * world) // max(ndocs, 1))
    while doc < shard_range(ndocs, r, world)[0]:
        r -= 1
    while doc >= shard_range(ndocs, r, world)[1]:
        r += 1
    return r


def counts_from_offsets(tuple_off: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """Per-document tuple counts (int32) from a doc_tuple_off array (int64[ndocs+1])."""
    if out is None:
        out = torch.empty(tuple_off.numel() - 1, dtype=torch.int32, device=tuple_off.device)
    torch.sub(tuple_off[1:], tuple_off[:-1], out=out)
    return out


def exchange_counts(counts_local: torch.Tensor, ndocs: int, rank: int, world: int, counts_all: torch.Tensor = None, group=None):
    """All-gather of per-document tuple counts.  Returns int32[ndocs] in global document order.

    Even shards go through one all_gather_into_tensor straight into `counts_all`; uneven shards are padded to
    the largest shard and compacted afterwards."""
    if world == 1:
        return counts_local
    sizes = [shard_range(ndocs, r, world)[1] - shard_range(ndocs, r, world)[0] for r in range(world)]
    if counts_local.numel() != sizes[rank]:
        raise ValueError("counts_local does not match this rank's shard")
    if min(sizes) == max(sizes):
        if counts_all is None:
            counts_all = torch.empty(ndocs, dtype=counts_local.dtype, device=counts_local.device)
        dist.all_gather_into_tensor(counts_all, counts_local, group=group)
        return counts_all
    pad = max(sizes)
    src = torch.zeros(pad, dtype=counts_local.dtype, device=counts_local.device)
    src[: sizes[rank]] = counts_local
    g = torch.empty(pad * world, dtype=counts_local.dtype, device=counts_local.device)
    dist.all_gather_into_tensor(g, src, group=group)
    return torch.cat([g[r * pad: r * pad + sizes[r]] for r in range(world)])


def global_index(counts_all: torch.Tensor, ndocs: int, world: int):
    """(rank_of_doc int32[ndocs], local_tuple_off int64[ndocs]): where document d's tuples start inside the
    tuple buffer of the rank that owns it."""
    rank_of = torch.empty(ndocs, dtype=torch.int32, device=counts_all.device)
    local_off = torch.empty(ndocs, dtype=torch.int64, device=counts_all.device)
    for r in range(world):
        d0, d1 = shard_range(ndocs, r, world)
        rank_of[d0:d1] = r
        c = counts_all[d0:d1].to(torch.int64)
        local_off[d0:d1] = torch.cumsum(c, 0) - c
    return rank_of, local_off
