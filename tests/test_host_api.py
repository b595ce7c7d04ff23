"""Host-side logic of the net.jpountz mirror that runs before any native call: argument checking
(SafeUtils.checkRange / ByteBufferUtils.checkNotReadOnly semantics), level clamping, layouts."""
import numpy as np
import pytest


def test_check_range_semantics(b200):
    from lz4java_b200.lz4 import _check_range, _view
    a = _view(bytes(10))
    _check_range(a, 0, 10)
    _check_range(a, 9, 1)
    _check_range(a, 5, 0)
    with pytest.raises(ValueError):          # IllegalArgumentException("lengths must be >= 0"), SafeUtils.java:35
        _check_range(a, 0, -1)
    with pytest.raises(IndexError):          # ArrayIndexOutOfBoundsException
        _check_range(a, 10, 1)
    with pytest.raises(IndexError):
        _check_range(a, 5, 6)
    with pytest.raises(IndexError):
        _check_range(a, -1, 1)


def test_read_only_dest_is_rejected_before_native_call(b200):
    """LZ4Test.java:421-454: ReadOnlyBufferException before anything native happens (works without a GPU)"""
    comp = b200.LZ4Compressor()
    with pytest.raises(BufferError):
        comp.compress(b"abcd" * 10, 0, 40, bytes(100), 0, 100)
    ro = np.zeros(100, dtype=np.uint8)
    ro.flags.writeable = False
    with pytest.raises(BufferError):
        b200.LZ4SafeDecompressor().decompress(b"\x00", 0, 1, ro, 0, 100)
    with pytest.raises(BufferError):
        b200.LZ4FastDecompressor().decompress(b"\x00", 0, ro, 0, 0)


def test_range_errors_before_native_call(b200):
    comp = b200.LZ4Compressor()
    with pytest.raises(IndexError):
        comp.compress(b"abcd", 2, 10, bytearray(100), 0, 100)
    with pytest.raises(ValueError):
        comp.compress(b"abcd", 0, -1, bytearray(100), 0, 100)
    with pytest.raises(IndexError):
        b200.XXHash32().hash(b"abcd", 3, 2, 0)


def test_uniform_layout(b200):
    off, ln = b200.batch.uniform_layout(4, 65536, 65824)
    assert off.dtype == np.uint64 and ln.dtype == np.int32
    assert list(off) == [0, 65824, 131648, 197472] and list(ln) == [65536] * 4


def test_shard_ranges_cover_and_ascend(b200):
    from lz4java_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_range(n, world, r) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            for (a, b), (c, d) in zip(pieces, pieces[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1


def test_frame_boundaries_and_byte_balanced_frame_shards(port):
    """config 3 shards BY FRAME (a content checksum chains over a frame's blocks): the boundary walk finds every frame
    of a concatenated stream whatever its flags, and the ranges handed to the ranks cover each frame once, balanced by
    bytes; each rank's slice decodes (CPU checker here) to exactly its part of the whole"""
    from lz4java_b200.sharding import frame_boundaries, shard_frames
    import random
    rng = random.Random(11)
    skip = bytes([0x53, 0x2A, 0x4D, 0x18, 5, 0, 0, 0, 9, 8, 7, 6, 5])
    parts, plain = [], []
    for k in range(23):
        n = rng.choice([0, 1, 100, 65536, 70000, 300000])
        d = port.datagen(n, 0.5, 0.0, k).tobytes() if k % 5 else rng.randbytes(n)      # some frames hold stored blocks
        parts.append(port.frame_compress(d, rng.choice([4, 5, 7]), rng.choice([0, 1, 3, 5, 7]))); plain.append(d)
        if k % 7 == 3:
            parts.append(skip); plain.append(b"")
    stream = b"".join(parts)
    bounds = frame_boundaries(stream)
    assert [e - s for s, e in bounds] == [len(p) for p in parts] and bounds[0][0] == 0 and bounds[-1][1] == len(stream)
    sizes = [e - s for s, e in bounds]
    for world in (1, 2, 3, 8, 64):
        ranges = shard_frames(sizes, world)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == len(sizes)
        assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
        got = b""
        for lo, hi in ranges:
            if hi > lo:
                piece = stream[bounds[lo][0]:bounds[hi - 1][1]]
                want = b"".join(plain[lo:hi])
                r, out = port.frame_decompress(piece, len(want) + 8)
                assert r == len(want) and out == want
                got += out
        assert got == b"".join(plain)
        if world <= 3:                                            # balanced: no rank above its share by more than one frame
            share = len(stream) / world
            assert max(sum(sizes[lo:hi]) for lo, hi in ranges) <= share + max(sizes)
    for bad in (stream[:-3], b"\x01\x02\x03\x04" + stream, stream[:bounds[2][0] + 5]):
        with pytest.raises((EOFError, IOError)):
            frame_boundaries(bad)


def test_bench_host_helpers(monkeypatch):
    """bench.py cannot run here (no GPU), but its host-side sizing logic can: thread counts tried for the CPU legs under a
    cgroup quota, and the memory room the e2e leg sizes its pinned buffers by"""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "cpu_threads", lambda: 128)
    monkeypatch.setattr(bench, "cpu_quota", lambda: "1600000 100000")
    assert bench.quota_cpus() == 16 and bench.thread_candidates() == [128, 64, 32, 16]
    monkeypatch.setattr(bench, "cpu_quota", lambda: "max 100000")
    assert bench.quota_cpus() is None and bench.thread_candidates() == [128, 64]
    monkeypatch.setattr(bench, "cpu_quota", lambda: "-1")
    assert bench.thread_candidates() == [128, 64]
    monkeypatch.setattr(bench, "cpu_quota", lambda: None)
    assert bench.thread_candidates() == [128, 64]
    monkeypatch.setattr(bench, "cpu_threads", lambda: 8)
    monkeypatch.setattr(bench, "cpu_quota", lambda: "1600000 100000")
    assert bench.thread_candidates() == [8, 4]                      # quota above the visible CPUs: nothing to add
    room = bench.host_memory_budget()
    assert room is None or room > (1 << 28)
    # the bench arms keep the contract's keys
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"ms_per_step"', '"higher_is_better"', '"scaling"', '"vs_baseline"', '"dtype"',
                '"roofline"', '"cpu_baseline"', '"e2e"', '"gpu_launches"', '"clocks"', '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"impl": "reference"'):
        assert key in src, key
