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


def stitch_packed(packed, local_total: int, dst_rank: int = 0, group=None):
    """One process per GPU: every rank holds its packed shard (`packed[:local_total]`, a uint8 tensor on its GPU); rank
    `dst_rank` gets the shards of all ranks back to back, in rank order, in one tensor on ITS GPU -- the device-side
    counterpart of the gather-write a host caller does with `*_compact_host_multi`'s pieces.
    Exchange: one all_gather of an int64 per rank (`packed_offsets`), then one point-to-point transfer per rank straight
    into the destination tensor at the computed offset (NCCL send/recv: GPU-to-GPU over NVLink, no host copy; gloo on CPU
    tensors in the tests).  -> (stitched tensor or None on the other ranks, this rank's offset, grand total)"""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([local_total], dtype=torch.int64, device=dev)
    alls = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(alls, mine, group=group)
    totals = [int(t.item()) for t in alls]
    offs = [sum(totals[:r]) for r in range(world)]
    out = None
    ops = []
    if rank == dst_rank:
        out = torch.empty(max(sum(totals), 1), dtype=torch.uint8, device=packed.device)
        out[offs[rank]:offs[rank] + totals[rank]].copy_(packed[:totals[rank]])
        for r in range(world):
            if r != rank and totals[r]:
                ops.append(dist.P2POp(dist.irecv, out[offs[r]:offs[r] + totals[r]], r, group))
    elif local_total:
        ops.append(dist.P2POp(dist.isend, packed[:local_total].contiguous(), dst_rank, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if out is not None:
        out = out[:sum(totals)]
    return out, offs[rank], sum(totals)


def max_over_ranks(seconds: float, group=None) -> float:
    """device-time aggregation rule of the bench contract: a multi-GPU step takes as long as its slowest rank"""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def frame_boundaries(buf) -> list[tuple[int, int]]:
    """[(start, end)) of every frame (skippable frames included) in a buffer of concatenated LZ4 frames — the walk
    LZ4FrameInputStream.nextFrameInfo / readHeader / readBlock make (LZ4FrameInputStream.java:132-321), sizes only:
    no payload byte is read, no checksum verified (the decoder does that).  Sharding metadata, not a codec: a frame's
    content checksum chains over all its blocks (`:264-273`), so config 3 is split across GPUs BY FRAME, never by block."""
    b = memoryview(buf).cast("B")
    n, ip, out = len(b), 0, []

    def le32(p):
        if n - p < 4:
            raise EOFError("Stream ended prematurely")
        return b[p] | (b[p + 1] << 8) | (b[p + 2] << 16) | (b[p + 3] << 24)

    while ip < n:
        start = ip
        magic = le32(ip); ip += 4
        if (magic >> 4) == (0x184D2A50 >> 4):                     # skippable frame: 4-byte size, payload
            ip += 4 + le32(ip)
        elif magic == 0x184D2204:
            if n - ip < 3:
                raise EOFError("Stream ended prematurely")
            flg = b[ip]
            ip += 2 + (8 if flg & 8 else 0) + 1                   # FLG, BD, [content size], HC
            while True:
                word = le32(ip); ip += 4
                size = word & 0x7FFFFFFF
                if size == 0:
                    break                                         # EndMark
                ip += size + (4 if flg & 0x10 else 0)             # payload, [block checksum]
            if flg & 4:
                ip += 4                                           # content checksum
        else:
            raise IOError("Stream unsupported (invalid magic bytes)")
        if ip > n:
            raise EOFError("Stream ended prematurely")
        out.append((start, ip))
    return out


def shard_frames(sizes, world: int) -> list[tuple[int, int]]:
    """Contiguous frame ranges [lo, hi) for ranks 0..world-1, balanced by BYTES (frames differ in size, so an equal
    frame count can leave one GPU with most of the work): rank r's range ends at the first frame boundary at or past
    r+1 shares of the total.  Every frame lands on exactly one rank; trailing ranks may be empty when frames are few."""
    total = sum(sizes)
    cuts, acc, k = [0], 0, 0
    for r in range(1, world):
        target = total * r / world
        while k < len(sizes) and acc + sizes[k] / 2 <= target:    # a frame goes to the side its midpoint is on
            acc += sizes[k]; k += 1
        cuts.append(k)
    cuts.append(len(sizes))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]
