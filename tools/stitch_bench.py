"""Device-side stitch of per-GPU packed shards over NVLink, measured (NOT run yet: written after round 2's GPU budget was
spent; the same code paths run in tests/test_sharding_gloo.py over gloo and in test_device_side_compaction_and_stitch on the
emulator build / one GPU).  Every rank compresses its range of 64 KiB blocks, packs it on its GPU (b200lz4_compact_dev) and
the shards are stitched into ONE stream on rank 0's GPU (sharding.stitch_packed: an 8-byte all_gather + one NCCL send/recv per
rank at the computed offset).  Rank 0 then decodes every block of the stitched stream and compares XXH64s with what the
owning ranks hashed before compressing.  Prints one JSON line: bytes stitched, GB/s into rank 0 (device events, max over ranks).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/stitch_bench.py [blocks_per_gpu]
With one process (no torchrun) and several visible GPUs it measures the one-process flavour instead (b200lz4_stitch_shards_dev:
cudaMemcpyPeerAsync per shard)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import lz4java_b200 as L
from lz4java_b200 import sharding
from oracle import oracle as O            # the data generator only (tools/ is test infrastructure)

B = L.batch
BS = 65536
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def make_shard(dev, seed):
    """-> (src, packed, packed_off, clen, total, hashes) on `dev`"""
    torch.cuda.set_device(dev); L._native.lib().b200lz4_set_device(dev.index)
    base_n = min(nblk, 2048)
    host = O.best_available().datagen(base_n * BS, 0.5, 0.0, seed)
    src = torch.from_numpy(host).to(dev).repeat((nblk + base_n - 1) // base_n)[: nblk * BS].contiguous()
    v = src.view(nblk, BS); idx = torch.arange(nblk, device=dev, dtype=torch.int64)
    for k in range(3):
        v[:, k] ^= ((idx >> (8 * k)) & 0xFF).to(torch.uint8)
    bound = L.max_compressed_length(BS); stride = (bound + 15) // 16 * 16
    soff = idx * BS; slen = torch.full((nblk,), BS, device=dev, dtype=torch.int32)
    coff = idx * stride; ccap = torch.full((nblk,), bound, device=dev, dtype=torch.int32)
    slots = torch.empty(nblk * stride, device=dev, dtype=torch.uint8); clen = torch.zeros(nblk, device=dev, dtype=torch.int32)
    hashes = torch.zeros(nblk, device=dev, dtype=torch.int64)
    B.xxh64_batch_dev(src, soff, slen, hashes, 0)
    B.compress_fast_batch_dev(src, soff, slen, slots, coff, ccap, clen, BS)
    packed = torch.empty(nblk * stride, device=dev, dtype=torch.uint8)
    poff = torch.zeros(nblk, device=dev, dtype=torch.int64); tot = torch.zeros(1, device=dev, dtype=torch.int64)
    B.compact_dev(slots, coff, clen, packed, poff, tot)
    torch.cuda.synchronize(dev)
    del slots
    return src, packed, poff, clen, int(tot.item()), hashes


def verify(dev, stitched, offs, lens, hashes):
    torch.cuda.set_device(dev); L._native.lib().b200lz4_set_device(dev.index)
    n = offs.numel()
    out = torch.empty(n * BS, device=dev, dtype=torch.uint8); res = torch.zeros(n, device=dev, dtype=torch.int32)
    doff = torch.arange(n, device=dev, dtype=torch.int64) * BS; dlen = torch.full((n,), BS, device=dev, dtype=torch.int32)
    B.decompress_safe_batch_dev(stitched, offs, lens, out, doff, dlen, res)
    got = torch.zeros(n, device=dev, dtype=torch.int64)
    B.xxh64_batch_dev(out, doff, dlen, got, 0)
    torch.cuda.synchronize(dev)
    return bool((res == BS).all().item()) and bool(torch.equal(got, hashes))


def main_ranks():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("NCCL_DEBUG", "WARN")
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    src, packed, poff, clen, total, hashes = make_shard(dev, 100 + rank)
    best = None
    for it in range(4):                                         # first pass warms NCCL's p2p channels up
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, my_off, grand = sharding.stitch_packed(packed, total, dst_rank=0)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it: best = min(best or 1e9, float(t.item()))
    # block table of the stitched stream on rank 0: offsets shifted by the shard's position, lengths, hashes
    offs_all = [torch.zeros_like(poff) for _ in range(world)]; lens_all = [torch.zeros_like(clen) for _ in range(world)]
    hash_all = [torch.zeros_like(hashes) for _ in range(world)]
    dist.all_gather(offs_all, poff + my_off); dist.all_gather(lens_all, clen); dist.all_gather(hash_all, hashes)
    if rank == 0:
        ok = verify(dev, out, torch.cat(offs_all), torch.cat(lens_all), torch.cat(hash_all))
        moved = grand - total                                  # rank 0's own shard is a local copy
        print(json.dumps({"stitch": "sharding.stitch_packed (NCCL send/recv into rank 0)", "gpus": world, "blocks_per_gpu": nblk,
                          "stitched_bytes": grand, "bytes_over_nvlink": moved, "seconds": best,
                          "GBps_into_rank0": moved / best / 1e9, "verified": ok}))
    dist.barrier(); dist.destroy_process_group()


def main_one_process():
    ndev = torch.cuda.device_count()
    shards = [make_shard(torch.device("cuda", g), 100 + g) for g in range(ndev)]
    totals = [s[4] for s in shards]
    dst = torch.empty(sum(totals) + 16, device=torch.device("cuda", 0), dtype=torch.uint8)
    import time
    best = None
    for it in range(4):
        t0 = time.perf_counter()
        pos = B.stitch_shards_dev([s[1] for s in shards], totals, dst)
        dt = time.perf_counter() - t0                           # the call returns when every copy has landed
        if it: best = min(best or 1e9, dt)
    d0 = torch.device("cuda", 0)
    offs = torch.cat([(s[2] + int(p)).to(d0) for s, p in zip(shards, pos)]); lens = torch.cat([s[3].to(d0) for s in shards])
    hs = torch.cat([s[5].to(d0) for s in shards])
    ok = verify(d0, dst, offs, lens, hs)
    moved = sum(totals[1:])
    print(json.dumps({"stitch": "b200lz4_stitch_shards_dev (cudaMemcpyPeerAsync per shard)", "gpus": ndev, "blocks_per_gpu": nblk,
                      "stitched_bytes": sum(totals), "bytes_over_nvlink": moved, "seconds": best,
                      "GBps_into_gpu0": moved / best / 1e9 if moved else None, "verified": ok}))


if __name__ == "__main__":
    main_ranks() if world > 1 else main_one_process()
