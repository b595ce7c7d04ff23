"""The N>1 host path on CPU: two gloo ranks shard one block list by range, each produces its packed
output independently (here with the CPU checker standing in for the device kernels — this test is about
the sharding / offset / timing plumbing, not the codec), and the metadata exchange stitches the
pieces into exactly what a single rank would have produced."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_file, n, out_dir):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from lz4java_b200.sharding import shard_range, packed_offsets, max_over_ranks
    dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
    P = O.Port()
    data = P.datagen(n * 4096, 0.5, 0.0, 9)
    lo, hi = shard_range(n, world, rank)
    pieces = [P.compress(data[b * 4096:(b + 1) * 4096]) for b in range(lo, hi)]
    local = b"".join(pieces)
    off, total = packed_offsets(len(local))
    slowest = max_over_ranks(0.5 + rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([lo, hi, off, total, len(local), slowest], dtype=np.float64))
    open(os.path.join(out_dir, f"r{rank}.bin"), "wb").write(local)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_range_sharding(tmp_path, port):
    world, n = 2, 37
    rendezvous = str(tmp_path / "rdv")
    mp.spawn(_worker, args=(world, rendezvous, n, str(tmp_path)), nprocs=world, join=True)
    data = port.datagen(n * 4096, 0.5, 0.0, 9)
    single = b"".join(port.compress(data[b * 4096:(b + 1) * 4096]) for b in range(n))
    stitched = bytearray(len(single))
    covered = 0
    for r in range(world):
        lo, hi, off, total, ln, slowest = np.load(tmp_path / f"r{r}.npy")
        assert int(total) == len(single)
        assert slowest == 0.5 + (world - 1)                      # max over ranks
        blob = open(tmp_path / f"r{r}.bin", "rb").read()
        stitched[int(off):int(off) + len(blob)] = blob
        covered += int(hi - lo)
    assert covered == n and bytes(stitched) == single


def _stitch_worker(rank, world, port_file, n, out_dir, dst_rank):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from lz4java_b200.sharding import shard_range, stitch_packed
    dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
    P = O.Port()
    data = P.datagen(n * 4096, 0.5, 0.0, 9)
    lo, hi = shard_range(n, world, rank)
    local = b"".join(P.compress(data[b * 4096:(b + 1) * 4096]) for b in range(lo, hi))
    packed = torch.zeros(len(local) + 100, dtype=torch.uint8)         # a shard buffer larger than what was packed into it
    packed[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8) if local else packed[:0]
    out, off, total = stitch_packed(packed, len(local), dst_rank=dst_rank)
    assert (out is not None) == (rank == dst_rank)
    if out is not None:
        open(os.path.join(out_dir, "stitched.bin"), "wb").write(out.numpy().tobytes())
    np.save(os.path.join(out_dir, f"s{rank}.npy"), np.array([off, total, len(local)], dtype=np.float64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,dst_rank", [(2, 37, 0), (3, 2, 1)])
def test_stitch_packed_shards_into_one_stream(tmp_path, port, world, n, dst_rank):
    """the device-side stitch's exchange (sharding.stitch_packed) over gloo: one int64 per rank, then every rank's packed
    shard straight into the destination rank's tensor at its computed offset; with 3 ranks and 2 blocks one shard is empty"""
    mp.spawn(_stitch_worker, args=(world, str(tmp_path / "rdv"), n, str(tmp_path), dst_rank), nprocs=world, join=True)
    data = port.datagen(n * 4096, 0.5, 0.0, 9)
    single = b"".join(port.compress(data[b * 4096:(b + 1) * 4096]) for b in range(n))
    assert open(tmp_path / "stitched.bin", "rb").read() == single
    acc = 0
    for r in range(world):
        off, total, ln = np.load(tmp_path / f"s{r}.npy")
        assert int(off) == acc and int(total) == len(single)
        acc += int(ln)
