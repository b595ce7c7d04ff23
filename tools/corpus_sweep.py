"""SURVEY.md §8(d) corpus corners, device-resident, 64 KiB blocks: RDG_genBuffer P=0.20 / 0.50 / 0.80, random bytes, zeros.
For each: fast-compress GiB/s + ratio (next to the reference's ratio on a 64 MiB sample of the same bytes), safe- and
fast-decompress GiB/s, bit-exact round trip.  Development/record tool (bench.py is the contract):
    python tools/corpus_sweep.py > gpurun_out/corpus_sweep.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import _variant  # noqa: F401  (B200LZ4_TEST_SO development switch)
import lz4java_b200 as L
from oracle import oracle as O

BS = 65536
GIB = float(1 << 30)

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 1e3)
    return best

def main():
    chk = O.best_available()
    dev = torch.device("cuda:0")
    nblk = int(os.environ.get("NBLK", 16384)); base_n = 4096
    B = L.batch
    bound = L.max_compressed_length(BS); stride = (bound + 15) // 16 * 16
    soff = torch.arange(nblk, device=dev, dtype=torch.int64) * BS
    slen = torch.full((nblk,), BS, device=dev, dtype=torch.int32)
    coff = torch.arange(nblk, device=dev, dtype=torch.int64) * stride
    ccap = torch.full((nblk,), bound, device=dev, dtype=torch.int32)
    comp = torch.zeros(nblk * stride, device=dev, dtype=torch.uint8)
    clen = torch.zeros(nblk, device=dev, dtype=torch.int32)
    out = torch.zeros(nblk * BS, device=dev, dtype=torch.uint8)
    res = torch.zeros(nblk, device=dev, dtype=torch.int32)
    N = nblk * BS
    rows = {}
    corpora = [("rdg_p20", lambda: chk.datagen(base_n * BS, 0.20, 0.0, 2)), ("rdg_p50", lambda: chk.datagen(base_n * BS, 0.50, 0.0, 2)),
               ("rdg_p80", lambda: chk.datagen(base_n * BS, 0.80, 0.0, 2)),
               ("random", lambda: np.random.default_rng(5).integers(0, 256, base_n * BS, dtype=np.uint8)),
               ("zeros", lambda: np.zeros(base_n * BS, dtype=np.uint8))]
    for name, gen in corpora:
        host = np.ascontiguousarray(gen())
        src = torch.from_numpy(host).to(dev).repeat(nblk // base_n).contiguous()
        if name != "zeros":                      # make tiled blocks distinct (ratio unchanged); zeros stay zeros
            v = src.view(nblk, BS); idx = torch.arange(nblk, device=dev, dtype=torch.int64)
            for k in range(4): v[:, k] ^= ((idx >> (8 * k)) & 0xFF).to(torch.uint8)
        tc = timeit(lambda: B.compress_fast_batch_dev(src, soff, slen, comp, coff, ccap, clen, BS))
        C = int(clen.sum().item())
        ts = timeit(lambda: B.decompress_safe_batch_dev(comp, coff, clen, out, soff, slen, res))
        ok = bool((res == BS).all().item()) and bool(torch.equal(out, src))
        out.zero_()
        tf = timeit(lambda: B.decompress_fast_batch_dev(comp, coff, ccap, out, soff, slen, res))
        ok = ok and bool((res == clen).all().item()) and bool(torch.equal(out, src))
        sample = host[: 1024 * BS]
        ref_c = sum(len(chk.compress(sample[i * BS:(i + 1) * BS].tobytes())) for i in range(1024))
        rows[name] = {"compress_GiBps": N / tc / GIB, "ratio": N / C, "reference_ratio_64MiB_sample": 1024 * BS / ref_c,
                      "decompress_safe_GiBps": N / ts / GIB, "decompress_fast_GiBps": N / tf / GIB, "roundtrip_bit_exact": ok}
        print(name, json.dumps(rows[name]), file=sys.stderr, flush=True)
        del src
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "blocks": nblk, "block_bytes": BS, "corpora": rows}, indent=1))

if __name__ == "__main__":
    main()
