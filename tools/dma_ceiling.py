"""What the host side of an 8-GPU box can move: every rank (one per GPU, torchrun) copies 1 GiB pinned buffers H2D and D2H
at the same time, node-local or not, and the ranks are timed together.  This is the ceiling of any end-to-end codec call:
bench.py's e2e leg moves (1 + 1/ratio) bytes each way per payload byte.  Also a STREAM-style copy of host memory on
the rank's cores (what the pinned buffers' DRAM channels give).
    python -m torch.distributed.run --nproc-per-node 8 ... tools/dma_ceiling.py [--no-numa]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
numa = bench.numa_bind(local) if "--no-numa" not in sys.argv else {"node": None}
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1); h2.fill_(2)
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, it=6):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / it], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)


a = timed(lambda: d.copy_(h, non_blocking=True)); b = timed(lambda: h2.copy_(d2, non_blocking=True)); c = timed(both)
# host memory copy on this rank's cores (numpy memcpy, one thread per rank)
x = np.ones(1 << 28, dtype=np.uint8); y = np.empty_like(x)
t0 = time.perf_counter()
for _ in range(4): np.copyto(y, x)
hc = (time.perf_counter() - t0) / 4
if rank == 0:
    print(json.dumps({"gpus": world, "numa": numa, "h2d_GBps_total": world * n / a / 1e9, "d2h_GBps_total": world * n / b / 1e9,
                      "both_directions_GBps_total": 2 * world * n / c / 1e9, "per_gpu_both_GBps": 2 * n / c / 1e9,
                      "host_memcpy_one_thread_GBps_rw": 2 * (1 << 28) / hc / 1e9,
                      "note": "pinned 1 GiB buffers per direction per GPU, all ranks at once, max over ranks"}))
if world > 1: dist.destroy_process_group()
