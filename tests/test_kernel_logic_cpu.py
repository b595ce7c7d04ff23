"""The CUDA decoder kernels' own source, run on the CPU by a SIMT emulator (tests/simt/simt.h), against the oracle.

No GPU is involved and nothing here is a product path: lz4-java_b200/csrc/lz4_decompress.cu is compiled as host C++
with -DB200_HOST_SIM (every CUDA thread of a CTA becomes a coroutine; warp collectives and barriers are emulated),
and the four kernels — safe/fast x batched/sequential — are fuzzed on inputs the GPU tests also use.  This checks the
kernels' logic (token walk, margins, dependency rounds, the reference's accept/reject rules and return codes), not
timing or the GPU memory model.  It exists because the build box has no GPU: a logic bug in the batched decoder
(a long sequence early in a batch pushing later ones past the output margin) was found late in round 1 by a GPU
sweep; on the pre-fix source this file's test_walk_stops_at_the_stream_end fails on the same six blocks a CPU model
predicted, on the fixed source it passes."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import corpus

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PAD = 4096


@pytest.fixture(scope="module")
def sim():
    out = os.path.join(HERE, "simt", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libdecsim.so")
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-Wno-attributes", "-DB200_HOST_SIM",
           "-I" + os.path.join(HERE, "simt"), "-I" + os.path.join(ROOT, "lz4-java_b200", "csrc"),
           os.path.join(HERE, "simt", "dec_harness.cpp"), "-o", so]
    subprocess.run(cmd, check=True, capture_output=True)
    lib = ctypes.CDLL(so)
    for f in (lib.sim_decompress_safe, lib.sim_decompress_fast):
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return lib


def _src(b):
    a = np.zeros(len(b) + 2 * PAD, dtype=np.uint8)
    a[PAD:PAD + len(b)] = np.frombuffer(b, dtype=np.uint8)
    return a


def run_safe(sim, c, cap, batched):
    s = _src(c); d = np.full(max(cap, 0) + 2 * PAD, 0x55, dtype=np.uint8)
    r = sim.sim_decompress_safe(s.ctypes.data + PAD, len(c), d.ctypes.data + PAD, cap, batched)
    assert (d[:PAD] == 0x55).all() and (d[PAD + max(cap, 0):] == 0x55).all(), "wrote outside [dst, dst+cap)"
    return r, d[PAD:PAD + max(r, 0)].tobytes()


def run_fast(sim, c, n, batched, readable=None):
    """readable: bytes the decoder may read (>= the stream); default = the stream itself"""
    data = c if readable is None else readable
    s = _src(data); d = np.full(n + 2 * PAD, 0x55, dtype=np.uint8)
    r = sim.sim_decompress_fast(s.ctypes.data + PAD, len(data), d.ctypes.data + PAD, n, batched)
    assert (d[:PAD] == 0x55).all() and (d[PAD + n:] == 0x55).all(), "wrote outside [dst, dst+n)"
    return r, d[PAD:PAD + n].tobytes()


@pytest.mark.parametrize("batched", [1, 0], ids=["batched", "sequential"])
def test_corpus_round_trips(sim, port, batched):
    for name, d in corpus.blocks(port):
        if len(d) > 300000:
            continue
        c = port.compress(d)
        r, o = run_safe(sim, c, len(d), batched)
        assert r == len(d) and o == d, (name, r)
        r, o = run_fast(sim, c, len(d), batched)
        assert r == len(c) and o == d, (name, r)


@pytest.mark.parametrize("batched", [1, 0], ids=["batched", "sequential"])
def test_safe_decoder_return_codes_on_malformed_input(sim, port, batched):
    """same accept/reject set and the same negative codes as the reference (lz4.c:2337), like the GPU test"""
    rng = random.Random(99)
    cases = []
    for name, d in corpus.blocks(port, big=False)[::3]:
        c = port.compress(d); n = len(d)
        for cap in (n, n - 1, n + 1, n + 64, max(0, n - 13), 0, n // 2):
            cases.append((c, cap))
        for cut in (1, 3, 8):
            if len(c) > cut:
                cases.append((c[:-cut], n))
        cases.append((c + b"\x10\x41", n))
        for m in corpus.mutate(c, rng, 6):
            cases.append((m, rng.choice([n, n + 1, n - 1, n + 70])))
    for v in corpus.MALFORMED:
        for cap in (20, 64, 200):
            cases.append((v, cap))
    negatives = 0
    for k, (c, cap) in enumerate(cases):
        if not c:
            continue
        want, out = port.decompress_safe(c, cap)
        r, o = run_safe(sim, c, cap, batched)
        assert r == want, (k, len(c), cap, r, want, c[:16].hex())
        if want >= 0:
            assert o == out, k
        else:
            negatives += 1
    assert negatives > 50


@pytest.mark.parametrize("batched", [1, 0], ids=["batched", "sequential"])
def test_fast_decoder_on_malformed_input(sim, port, batched):
    rng = random.Random(7)
    for name, d in corpus.blocks(port, big=False)[::4]:
        c = port.compress(d); n = len(d)
        cases = [(c, dl) for dl in (n, n - 1, n + 1, n + 5, max(0, n - 12)) if dl >= 0]
        cases += [(m, n) for m in corpus.mutate(c, rng, 4)]
        for cc, dl in cases:
            if not cc:
                continue
            padded = cc + bytes(dl + dl // 255 + 64)          # the reference's unbounded reads stay defined
            want, out = port.decompress_fast(padded, dl)
            r, o = run_fast(sim, padded, dl, batched)
            assert r == want, (name, len(cc), dl, r, want)
            if want >= 0:
                assert o == out, name


def _tail_heavy_blocks():
    rng = random.Random(8080)
    items = []
    for trial in range(96):
        hist = bytes(rng.randrange(256) for _ in range(3000))
        parts = [hist]

        def short():
            parts.append(bytes(rng.randrange(256) for _ in range(rng.randrange(1, 7))))
            a = rng.randrange(0, 2900); parts.append(hist[a:a + rng.randrange(5, 13)])
        for _ in range(rng.randrange(40, 90)): short()
        a = rng.randrange(0, 500)
        parts.append(bytes(rng.randrange(256) for _ in range(2))); parts.append(hist[a:a + rng.randrange(1100, 1700)])
        for _ in range(rng.randrange(3, 26)): short()
        parts.append(bytes(rng.randrange(256) for _ in range(rng.randrange(5, 15))))
        items.append(b"".join(parts))
    return items


@pytest.mark.parametrize("batched", [1, 0], ids=["batched", "sequential"])
def test_walk_stops_at_the_stream_end(sim, port, batched):
    """the data of tests/test_gpu_parity.py::test_decompress_fast_does_not_walk_past_the_stream: what follows each
    stream starts with a valid-looking offset and continues with another block's sequences"""
    items = _tail_heavy_blocks()
    comp = [port.compress(d) for d in items]
    for k, (d, c) in enumerate(zip(items, comp)):
        readable = c + b"\x01\x00" + comp[(k + 1) % len(comp)][9:3000] + bytes(64)
        r, o = run_fast(sim, c, len(d), batched, readable=readable)
        assert r == len(c) and o == d, (k, r, len(c))
        r, o = run_safe(sim, c, len(d) + (k % 3) * 40, batched)
        assert r == len(d) and o == d, (k, r)


@pytest.mark.parametrize("batched", [1, 0], ids=["batched", "sequential"])
def test_dependency_patterns_and_extremes(sim, port, batched):
    rng = random.Random(4242)
    items = [bytes(70000), bytes([7]) * 40000 + bytes(rng.randrange(256) for _ in range(300)) + bytes([9]) * 3000,
             bytes(rng.randrange(256) for _ in range(20000))]
    for trial in range(10):
        parts = []
        while sum(map(len, parts)) < 20000 + 3000 * trial:
            kind = rng.randrange(7)
            if kind == 0:
                parts.append(bytes([rng.randrange(256)]) * rng.randrange(5, 700))
            elif kind == 1:
                pat = bytes(rng.randrange(256) for _ in range(rng.randrange(2, 9))); parts.append(pat * rng.randrange(3, 120))
            elif kind == 2:
                parts.append(bytes(rng.randrange(256) for _ in range(rng.randrange(1, 80))))
            elif kind == 3 and parts:
                prev = b"".join(parts[-3:]); a = rng.randrange(len(prev)); parts.append(prev[a:a + rng.randrange(4, 60)])
            elif kind == 4 and parts:
                whole = b"".join(parts); a = rng.randrange(len(whole)); parts.append(whole[a:a + rng.randrange(4, 400)])
            elif kind == 5:
                pat = bytes(rng.randrange(256) for _ in range(rng.randrange(33, 200))); parts.append(pat * rng.randrange(2, 6))
            else:
                parts.append(bytes(rng.randrange(4) for _ in range(rng.randrange(20, 300))))
        items.append(b"".join(parts))
    for k, d in enumerate(items):
        c = port.compress(d)
        r, o = run_safe(sim, c, len(d), batched)
        assert r == len(d) and o == d, (k, r)
        r, o = run_fast(sim, c, len(d), batched, readable=c + bytes(3000))
        assert r == len(c) and o == d, (k, r)
