"""Range sharding of a block batch across GPUs.

Blocks are independent, so GPU g of G owns one contiguous range and no data-path collective is
needed.  The only cross-rank traffic is metadata: per-rank packed sizes, exchanged once when the
caller wants every rank's compacted output laid out in one stream (`packed_offsets`)."""
from __future__ import annotations


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous [lo, hi) of n blocks owned by `rank`; sizes differ by at most one block"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def packed_offsets(local_total: int, group=None) -> tuple[int, int]:
    """(this rank's byte offset in the concatenation of every rank's packed output, grand total).
    One all_gather of a single int64 per rank — 8 x 8 bytes on an 8-GPU node."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([local_total], dtype=torch.int64, device=dev)
    alls = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(alls, mine, group=group)
    totals = [int(t.item()) for t in alls]
    return sum(totals[:rank]), sum(totals)


def max_over_ranks(seconds: float, group=None) -> float:
    """device-time aggregation rule of the bench contract: a multi-GPU step takes as long as its slowest rank"""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
