"""Open-ended fuzz of the library's HOST layer on the CPU emulator build (tests/simt/build_sim_library.sh): random batch
layouts — gaps between source blocks, scattered or back-to-back destination slots, empty and tiny blocks, 1 MiB pipeline
chunks so that small batches cross chunk boundaries, random device lists over pretend GPUs — through the single-GPU and the
range-sharded multi-GPU entry points, compact and slot layouts, against the CPU checker.

  SIMT_DEVICES=3 B200LZ4_CHUNK_MB=1 B200LZ4_TEST_SO=tests/simt/_build/libb200lz4_sim.so python tools/fuzz_host_sim.py SEED SECONDS
"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _variant  # noqa: F401
import lz4java_b200 as L
from oracle import oracle as O


def main():
    seed, budget = int(sys.argv[1]), float(sys.argv[2])
    assert "sim" in os.environ.get("B200LZ4_TEST_SO", ""), "meant for the emulator build (or its sanitizer twin)"
    rng = random.Random(seed)
    port = O.Port()
    B = L.batch
    ndev = L._native.lib().b200lz4_device_count()
    t0 = time.time(); cases = 0
    while time.time() - t0 < budget:
        n = rng.choice([1, 2, 3, 5, 9, 17, 40])
        big = rng.random() < 0.3
        datas = []
        for _ in range(n):
            ln = rng.choice([0, 1, 12, 13, rng.randrange(0, 300), rng.randrange(0, 5000), rng.randrange(0, 70000 if big else 9000)])
            kind = rng.randrange(3)
            datas.append(port.datagen(ln, rng.choice([0.2, 0.5, 0.9]), 0.0, rng.randrange(1 << 30)).tobytes() if kind == 0 else
                         rng.randbytes(ln) if kind == 1 else bytes(rng.choice(b"ab") for _ in range(ln)))
        # source arena with random gaps and alignment
        soff, pos = [], rng.randrange(0, 16)
        for d in datas:
            soff.append(pos); pos += len(d) + rng.choice([0, 0, 1, 7, 100])
        src = np.frombuffer(rng.randbytes(pos + 64), dtype=np.uint8).copy()
        for o, d in zip(soff, datas):
            src[o:o + len(d)] = np.frombuffer(d, dtype=np.uint8)
        soff = np.array(soff, dtype=np.uint64); slen = np.array([len(d) for d in datas], dtype=np.int32)
        # destination slots: back to back or with gaps (the scattered-dst bounce path); gap bytes must stay untouched
        caps = [L.max_compressed_length(len(d)) for d in datas]
        gaps = rng.random() < 0.5
        coff, pos = [], rng.randrange(0, 16)
        for c in caps:
            coff.append(pos); pos += c + (rng.choice([0, 3, 64]) if gaps else 0)
        coff = np.array(coff, dtype=np.uint64); ccap = np.array(caps, dtype=np.int32)
        devs = rng.choice([1, [0, 0], [rng.randrange(ndev) for _ in range(rng.randrange(1, 5))], list(range(ndev))])
        comp = np.full(pos + 64, 0xC3, dtype=np.uint8)
        multi = rng.random() < 0.6
        clen = (B.compress_fast_batch_host_multi(src, soff, slen, comp, coff, ccap, devs) if multi else
                B.compress_fast_batch_host(src, soff, slen, comp, coff, ccap))
        used = np.zeros(len(comp), dtype=bool)
        for k, d in enumerate(datas):
            o = int(coff[k]); used[o:o + caps[k]] = True
            c = comp[o:o + int(clen[k])].tobytes()
            assert int(clen[k]) > 0 and port.decompress_safe(c, len(d)) == (len(d), d), (seed, cases, k)
        assert (comp[~used] == 0xC3).all(), (seed, cases, "bytes between the slots were touched")
        # packed layout must hold the same streams
        pk = np.zeros(sum((c + 15) // 16 * 16 for c in caps) + 64, dtype=np.uint8)
        if multi:
            ooff, olen, sbase, stotal = B.compress_fast_compact_host_multi(src, soff, slen, pk, devs)
        else:
            ooff, olen, total = B.compress_fast_compact_host(src, soff, slen, pk)
            assert total == int(olen.sum())
        assert (olen == clen).all()
        for k in range(n):
            assert pk[int(ooff[k]):int(ooff[k]) + int(olen[k])].tobytes() == comp[int(coff[k]):int(coff[k]) + int(clen[k])].tobytes(), (seed, cases, k)
        # decode from the packed layout, both decoders, scattered output slots
        doff, pos = [], rng.randrange(0, 16)
        for d in datas:
            doff.append(pos); pos += len(d) + (rng.choice([0, 5, 64]) if gaps else 0)
        doff = np.array(doff, dtype=np.uint64)
        out = np.full(pos + 64, 0x3C, dtype=np.uint8)
        if multi:
            r = B.decompress_safe_batch_host_multi(pk, ooff, olen, out, doff, slen, devs)
        else:
            r = B.decompress_safe_batch_host(pk, ooff, olen, out, doff, slen)
        out2 = np.full(pos + 64, 0x3C, dtype=np.uint8)
        avail = np.minimum(olen.astype(np.int64) + 32, len(pk) - ooff.astype(np.int64)).astype(np.int32)
        r2 = (B.decompress_fast_batch_host_multi(pk, ooff, avail, out2, doff, slen, devs) if multi else
              B.decompress_fast_batch_host(pk, ooff, avail, out2, doff, slen))
        usedo = np.zeros(len(out), dtype=bool)
        for k, d in enumerate(datas):
            o = int(doff[k]); usedo[o:o + len(d)] = True
            assert int(r[k]) == len(d) and int(r2[k]) == int(olen[k]), (seed, cases, k, int(r[k]), int(r2[k]))
            assert out[o:o + len(d)].tobytes() == d and out2[o:o + len(d)].tobytes() == d, (seed, cases, k)
        assert (out[~usedo] == 0x3C).all() and (out2[~usedo] == 0x3C).all(), (seed, cases, "bytes between output slots were touched")
        h = B.xxh64_batch_host_multi(src, soff, slen, devs, 5) if multi else B.xxh64_batch_host(src, soff, slen, 5)
        for k in (0, n - 1):
            assert int(h[k]) == port.xxh64(datas[k], 5)
        cases += 1
    print("seed", seed, "cases", cases, "ok", flush=True)


if __name__ == "__main__":
    main()
