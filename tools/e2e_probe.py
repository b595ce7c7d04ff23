import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import _variant  # noqa: F401  (B200LZ4_TEST_SO development switch)
import lz4java_b200 as L
from oracle import oracle as O
B = L.batch; BLOCK = 65536
chk = O.best_available()
n = 65536; nbytes = n * BLOCK
host = chk.datagen(4096 * BLOCK, 0.5, 0.0, 2)
src = torch.empty(nbytes, dtype=torch.uint8).pin_memory().numpy()
for r in range(nbytes // len(host)): src[r * len(host):(r + 1) * len(host)] = host
bound = L.max_compressed_length(BLOCK)
comp = torch.empty(n * bound, dtype=torch.uint8).pin_memory().numpy()
out = torch.empty(nbytes, dtype=torch.uint8).pin_memory().numpy()
soff, slen = B.uniform_layout(n, BLOCK)
def T(fn, it=3):
    fn(); t0 = time.perf_counter()
    for _ in range(it): r = fn()
    return (time.perf_counter() - t0) / it, r
tc, (ooff, olen, total) = T(lambda: B.compress_fast_compact_host(src, soff, slen, comp, BLOCK))
td, res = T(lambda: B.decompress_fast_batch_host(comp, ooff, olen, out, soff, slen))
print(f"chunk {os.environ.get('B200LZ4_CHUNK_MB','256')} MiB: compress_compact_host {tc*1e3:.1f} ms ({nbytes/tc/2**30:.1f} GiB/s in, H2D {nbytes/tc/1e9:.1f} GB/s)  decompress_host {td*1e3:.1f} ms ({nbytes/td/2**30:.1f} GiB/s out)  total {total/1e9:.2f} GB", flush=True)
# slot-layout compress (no compaction) for comparison
coff, ccap = B.uniform_layout(n, bound)
ts, _ = T(lambda: B.compress_fast_batch_host(src, soff, slen, comp, coff, ccap, BLOCK))
print(f"  compress_batch_host (slots, D2H whole slots) {ts*1e3:.1f} ms", flush=True)

# concurrency test: compress loop and decompress loop in two threads at once
comp2 = torch.empty(n * bound, dtype=torch.uint8).pin_memory().numpy()
B.compress_fast_compact_host(src, soff, slen, comp2, BLOCK)
res = {}
def loopA():
    L._native.lib().b200lz4_set_device(0)
    B.compress_fast_compact_host(src, soff, slen, comp2, BLOCK)
    t0 = time.perf_counter()
    for _ in range(4): B.compress_fast_compact_host(src, soff, slen, comp2, BLOCK)
    res["A"] = (time.perf_counter() - t0) / 4
def loopB():
    t0 = time.perf_counter()
    for _ in range(4): B.decompress_fast_batch_host(comp, ooff, olen, out, soff, slen)
    res["B"] = (time.perf_counter() - t0) / 4
ta = threading.Thread(target=loopA); ta.start(); time.sleep(0.3); loopB(); ta.join()
print(f"  concurrent: compress {res['A']*1e3:.1f} ms/call, decompress {res['B']*1e3:.1f} ms/call (alone: {tc*1e3:.1f} / {td*1e3:.1f})", flush=True)
