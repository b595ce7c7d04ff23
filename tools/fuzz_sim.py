"""Long-running CPU fuzz of the kernels' own source under the SIMT emulator (tests/simt) against the oracle.
Not part of the test suite (tests/test_kernel_logic_cpu.py is the bounded version); run by hand:
    python tools/fuzz_sim.py <seed> <minutes>          # e.g. 12 processes with different seeds
Exits non-zero and prints a reproducer (hex) on the first disagreement."""
import ctypes, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import oracle as O
import corpus
import test_kernel_logic_cpu as T

PAD = T.PAD


def gen(rng, n):
    kind = rng.randrange(8)
    if kind == 0: return rng.randbytes(n)
    if kind == 1: return bytes(rng.randrange(4) for _ in range(n))
    if kind == 2: return (bytes(rng.randrange(256) for _ in range(rng.randrange(1, 40))) * (n // 1 + 1))[:n]
    if kind == 3: return bytes([rng.randrange(256)]) * n
    parts = []; tot = 0
    while tot < n:
        k = rng.randrange(6)
        if k == 0 or not parts: p = rng.randbytes(rng.randrange(1, 60))
        elif k == 1: p = bytes([rng.randrange(256)]) * rng.randrange(4, 400)
        elif k == 2:
            whole = b"".join(parts); a = rng.randrange(len(whole)); p = whole[a:a + rng.randrange(4, 300)]
        elif k == 3:
            prev = b"".join(parts[-2:]); a = rng.randrange(len(prev)); p = prev[a:a + rng.randrange(4, 40)]
        elif k == 4: p = bytes(rng.randrange(256) for _ in range(rng.randrange(2, 9))) * rng.randrange(2, 60)
        else: p = bytes(rng.randrange(3) for _ in range(rng.randrange(10, 200)))
        parts.append(p); tot += len(p)
    return b"".join(parts)[:n]


def main(seed, minutes):
    rng = random.Random(seed)
    chk = O.best_available()
    dec = T._build("dec_harness.cpp", f"libdecsim_{seed}.so"); 
    for f in (dec.sim_decompress_safe, dec.sim_decompress_fast):
        f.restype = ctypes.c_int; f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    cmp_ = T._build("comp_harness.cpp", f"libcompsim_{seed}.so")
    cmp_.sim_compress_fast.restype = ctypes.c_int
    cmp_.sim_compress_fast.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    t_end = time.time() + minutes * 60
    stats = {"decodes": 0, "malformed": 0, "compress": 0}
    while time.time() < t_end:
        n = rng.choice([rng.randrange(0, 40), rng.randrange(0, 700), rng.randrange(0, 6000), rng.randrange(0, 30000), rng.randrange(60000, 66000)])
        d = gen(rng, n)
        c = chk.compress(d)
        for b in (1, 0):
            r, o = T.run_safe(dec, c, len(d) + rng.choice([0, 0, 1, 7, 64, 300]), b)
            assert r == len(d) and o == d, ("safe", b, seed, len(d), r, c.hex() if len(c) < 400 else len(c))
            r, o = T.run_fast(dec, c, len(d), b, readable=c + rng.choice([b"", bytes(64), b"\x01\x00" + c[:2000], rng.randbytes(700)]))
            assert r == len(c) and o == d, ("fast", b, seed, len(d), r, c.hex() if len(c) < 400 else len(c))
            stats["decodes"] += 2
        for m in corpus.mutate(c, rng, 4):
            if not m: continue
            cap = rng.choice([len(d), len(d) + 1, max(0, len(d) - 1), len(d) + 70, len(d) // 2])
            want, out = chk.decompress_safe(m, cap)
            for b in (1, 0):
                r, o = T.run_safe(dec, m, cap, b)
                assert r == want and (want < 0 or o == out), ("safe-malformed", b, seed, cap, r, want, m.hex() if len(m) < 600 else len(m))
            padded = m + bytes(len(d) + len(d) // 255 + 64)
            want, out = chk.decompress_fast(padded, len(d))
            for b in (1, 0):
                r, o = T.run_fast(dec, padded, len(d), b)
                assert r == want and (want < 0 or o == out), ("fast-malformed", b, seed, r, want, m.hex() if len(m) < 600 else len(m))
            stats["malformed"] += 4
        if len(d) <= 66000:
            for variant in list(T.COMPRESS_KINDS):
                if variant != "long" and len(d) >= 65536 + 11: continue
                bound = chk.compress_bound(len(d))
                cap = rng.choice([bound, bound, rng.randrange(0, bound + 1), len(c), max(0, len(c) - 1), -1])
                r, cc = T.run_compress(cmp_, d, cap, variant, shift=rng.randrange(4))
                assert 0 <= r <= max(cap, 0), (variant, seed, len(d), cap, r)
                if r > 0:
                    rr, o = chk.decompress_safe(cc, len(d))
                    assert rr == len(d) and o == d, ("compress", variant, seed, len(d), cap)
                    r2, o2 = T.run_safe(dec, cc, len(d), 1)
                    assert r2 == len(d) and o2 == d, ("compress->sim decode", variant, seed, len(d))
                else:
                    assert cap < bound, ("refused at bound", variant, seed, len(d))
                stats["compress"] += 1
    print(seed, stats, flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]), float(sys.argv[2]))
