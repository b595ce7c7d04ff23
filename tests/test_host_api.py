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
