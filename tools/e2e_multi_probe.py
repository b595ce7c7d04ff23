"""End to end through the C ABI from ONE process driving several GPUs (the single-JVM case): the range-sharded host batch
calls b200lz4_compress_fast_compact_host_multi / b200lz4_decompress_fast_batch_host_multi over G = 1, 2, 4, 8 devices,
host buffers pinned with b200lz4_host_register, wall clock.  Strong scaling: the same NBLK blocks are split G ways.

  NBLK=65536 python tools/e2e_multi_probe.py          (needs `gpurun --gpus N`; on one GPU it lists device 0 twice as a
                                                        functional check of the sharded path, which is not a scaling number)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _variant  # noqa: F401  (B200LZ4_TEST_SO development switch)
import lz4java_b200 as L
from oracle import oracle as O

BLOCK = 65536


def main():
    nblk = int(os.environ.get("NBLK", 65536))
    iters = int(os.environ.get("ITERS", 3))
    lib = L._native.lib()
    ndev = L._native.check(lib.b200lz4_device_count())
    chk = O.best_available()
    base = chk.datagen(min(nblk, 4096) * BLOCK, 0.5, 0.0, 2)
    src = np.empty(nblk * BLOCK, dtype=np.uint8)
    for lo in range(0, len(src), len(base)):
        src[lo:lo + len(base)] = base[:len(src) - lo]
    bound = L.max_compressed_length(BLOCK)
    comp = np.empty(nblk * ((bound + 15) // 16 * 16), dtype=np.uint8)
    out = np.empty(nblk * BLOCK, dtype=np.uint8)
    for a in (src, comp, out):
        L._native.check(lib.b200lz4_host_register(a.ctypes.data, a.nbytes))
    soff, slen = L.batch.uniform_layout(nblk, BLOCK)
    coff, ccap = L.batch.uniform_layout(nblk, bound)
    lists = [g for g in (1, 2, 4, 8) if g <= ndev] or [1]
    if ndev == 1:
        lists.append([0, 0])
    for devs in lists:
        best_c = best_d = 1e30
        for _ in range(iters + 1):                                   # first pass: contexts + staging buffers
            t0 = time.perf_counter()
            ooff, clen, sbase, stotal = L.batch.compress_fast_compact_host_multi(src, soff, slen, comp, devs, BLOCK)
            t1 = time.perf_counter()
            res = L.batch.decompress_fast_batch_host_multi(comp, ooff, clen, out, soff, slen, devs)
            t2 = time.perf_counter()
            best_c, best_d = min(best_c, t1 - t0), min(best_d, t2 - t1)
        assert (res == clen).all() and (out == src).all(), "round trip mismatch"
        gib = nblk * BLOCK / 2 ** 30
        print(f"devices {devs}: compress {gib / best_c:.1f} GiB/s, decompress {gib / best_d:.1f} GiB/s, "
              f"round trip {gib / (best_c + best_d):.1f} GiB/s (ratio {nblk * BLOCK / int(clen.sum()):.3f}; packed per shard: only compressed bytes cross PCIe)", flush=True)
    for a in (src, comp, out):
        lib.b200lz4_host_unregister(a.ctypes.data)


if __name__ == "__main__":
    main()
