"""Host-side mirror of the net.jpountz.lz4 block API over libb200lz4 (the "B200" backend).

Same names, argument meaning and error behaviour as the reference's Java classes so the parity
tests read like the reference's own (src/test/net/jpountz/lz4/LZ4Test.java):

  LZ4Factory.b200Instance()      <- LZ4Factory.nativeInstance()        LZ4Factory.java:91-96
  LZ4Compressor.compress(...)    <- LZ4JNICompressor.compress          LZ4JNICompressor.java:35-43
  LZ4FastDecompressor.decompress <- LZ4JNIFastDecompressor.decompress  LZ4JNIFastDecompressor.java:36-44
  LZ4SafeDecompressor.decompress <- LZ4JNISafeDecompressor.decompress  LZ4JNISafeDecompressor.java:35-43
  LZ4Exception                   <- LZ4Exception.java:22

This environment has no JDK; the Java twin of this file lives in lz4-java_b200/java/ (see
INTEGRATION.md).  Java's byte[] is a Python bytes-like here; dest must be writable
(bytearray / numpy uint8 / writable memoryview).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

MAX_INPUT_SIZE = 0x7E000000          # LZ4Utils / lz4.h:211
DEFAULT_COMPRESSION_LEVEL = 9        # LZ4Constants.java:23
MAX_COMPRESSION_LEVEL = 17           # LZ4Constants.java:24


class LZ4Exception(RuntimeError):
    """LZ4Exception.java:22 — unchecked."""


def _view(buf, writable=False) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8 or not buf.flags.c_contiguous:
            raise TypeError("need a contiguous uint8 array")
        if writable and not buf.flags.writeable:
            raise BufferError("ReadOnlyBufferException")        # ByteBufferUtils.java:99-103
        return buf
    mv = memoryview(buf)
    if writable and mv.readonly:
        raise BufferError("ReadOnlyBufferException")
    return np.frombuffer(mv, dtype=np.uint8)


def _check_range(arr: np.ndarray, off: int, length: int | None = None):
    # SafeUtils.checkRange (SafeUtils.java:24-42)
    if length is None:
        if off < 0 or off >= len(arr):
            raise IndexError(off)
        return
    if length < 0:
        raise ValueError("lengths must be >= 0")
    if length > 0:
        if off < 0 or off >= len(arr):
            raise IndexError(off)
        if off + length - 1 >= len(arr):
            raise IndexError(off + length - 1)


def max_compressed_length(length: int) -> int:
    """LZ4Utils.maxCompressedLength (LZ4Utils.java:32-41); equals LZ4_compressBound."""
    if length < 0:
        raise ValueError("length must be >= 0, got " + str(length))
    if length >= MAX_INPUT_SIZE:
        raise ValueError("length must be < " + str(MAX_INPUT_SIZE))
    return length + length // 255 + 16


def _addr(a: np.ndarray, off: int) -> int:
    return a.ctypes.data + off


class LZ4Compressor:
    """LZ4Compressor.java — fast-scan block compressor (the GPU warp-greedy parser)."""

    def __init__(self, level: int | None = None):
        self._level = level

    def maxCompressedLength(self, length: int) -> int:
        return max_compressed_length(length)

    def compress(self, src, srcOff=None, srcLen=None, dest=None, destOff=0, maxDestLen=None):
        """compress(src, srcOff, srcLen, dest, destOff, maxDestLen) -> compressed length
        (LZ4Compressor.java:59); compress(src) -> bytes (LZ4Compressor.java:129-136)."""
        s = _view(src)
        if dest is None:                       # convenience overload: allocates maxCompressedLength
            so = 0 if srcOff is None else srcOff
            sl = len(s) - so if srcLen is None else srcLen
            out = bytearray(self.maxCompressedLength(sl))
            n = self.compress(s, so, sl, out, 0, len(out))
            return bytes(out[:n])
        d = _view(dest, writable=True)
        if maxDestLen is None:
            maxDestLen = len(d) - destOff
        _check_range(s, srcOff, srcLen)
        _check_range(d, destOff, maxDestLen)
        L = N.lib()
        if self._level is None:
            r = L.b200lz4_compress_default(_addr(s, srcOff), _addr(d, destOff), srcLen, maxDestLen)
        else:
            r = L.b200lz4_compress_HC(_addr(s, srcOff), _addr(d, destOff), srcLen, maxDestLen, self._level)
        N.check(r)
        if r <= 0:
            raise LZ4Exception("maxDestLen is too small")      # LZ4JNICompressor.java:39-41
        return r


class LZ4FastDecompressor:
    """LZ4FastDecompressor.java — needs the exact decompressed length, returns bytes READ."""

    def decompress(self, src, srcOff=0, dest=None, destOff=0, destLen=None):
        s = _view(src)
        if dest is None or isinstance(dest, int):              # decompress(src, destLen) -> bytes
            n = destLen if destLen is not None else dest
            if n is None:
                raise TypeError("destLen required")
            out = bytearray(n)
            self.decompress(s, srcOff, out, 0, n)
            return bytes(out)
        d = _view(dest, writable=True)
        if destLen is None:
            destLen = len(d) - destOff
        _check_range(s, srcOff)
        _check_range(d, destOff, destLen)
        r = N.lib().b200lz4_decompress_fast_bounded(_addr(s, srcOff), len(s) - srcOff, _addr(d, destOff), destLen)
        N.check(r)
        if r < 0:
            raise LZ4Exception("Error decoding offset " + str(srcOff - r) + " of input buffer")   # LZ4JNIFastDecompressor.java:40-42
        return r


class LZ4SafeDecompressor:
    """LZ4SafeDecompressor.java — knows the compressed length, returns bytes WRITTEN."""

    def decompress(self, src, srcOff=0, srcLen=None, dest=None, destOff=0, maxDestLen=None):
        s = _view(src)
        if srcLen is None:
            srcLen = len(s) - srcOff
        if dest is None or isinstance(dest, int):              # decompress(src, maxDestLen) -> bytes
            cap = maxDestLen if maxDestLen is not None else dest
            if cap is None:
                raise TypeError("maxDestLen required")
            out = bytearray(cap)
            n = self.decompress(s, srcOff, srcLen, out, 0, cap)
            return bytes(out[:n])
        d = _view(dest, writable=True)
        if maxDestLen is None:
            maxDestLen = len(d) - destOff
        _check_range(s, srcOff, srcLen)
        _check_range(d, destOff, maxDestLen)
        r = N.lib().b200lz4_decompress_safe(_addr(s, srcOff), _addr(d, destOff), srcLen, maxDestLen)
        N.check(r)
        if r < 0:
            raise LZ4Exception("Error decoding offset " + str(srcOff - r) + " of input buffer")   # LZ4JNISafeDecompressor.java:39-41
        return r


class LZ4Factory:
    """LZ4Factory.java — entry point; b200Instance() sits next to nativeInstance()/safeInstance()."""

    _instance = None

    def __init__(self):
        self._fast = LZ4Compressor()
        self._hc = {lvl: LZ4Compressor(lvl) for lvl in range(1, MAX_COMPRESSION_LEVEL + 1)}
        self._fast_dec = LZ4FastDecompressor()
        self._safe_dec = LZ4SafeDecompressor()
        # the reference's constructor self-test (LZ4Factory.java:204-220): a 20-byte block must
        # round-trip through both decompressors
        original = bytes([ord(c) for c in "abcdefghijklmnopqrstuvwxyz"[:20]])
        comp = self._fast.compress(original)
        if self._fast_dec.decompress(comp, 0, None, 0, len(original)) != original:
            raise AssertionError("fast decompressor self-test failed")
        if self._safe_dec.decompress(comp, 0, len(comp), None, 0, len(original)) != original:
            raise AssertionError("safe decompressor self-test failed")

    @classmethod
    def b200Instance(cls) -> "LZ4Factory":
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def fastCompressor(self) -> LZ4Compressor:
        return self._fast

    def highCompressor(self, compressionLevel: int = DEFAULT_COMPRESSION_LEVEL) -> LZ4Compressor:
        # LZ4Factory.java:263-270: >17 -> 17, <1 -> 9
        if compressionLevel > MAX_COMPRESSION_LEVEL:
            compressionLevel = MAX_COMPRESSION_LEVEL
        elif compressionLevel < 1:
            compressionLevel = DEFAULT_COMPRESSION_LEVEL
        return self._hc[compressionLevel]

    def fastDecompressor(self) -> LZ4FastDecompressor:
        return self._fast_dec

    def safeDecompressor(self) -> LZ4SafeDecompressor:
        return self._safe_dec

    def unknownSizeDecompressor(self) -> LZ4SafeDecompressor:
        """deprecated alias of safeDecompressor() (LZ4Factory.java:292-301)"""
        return self._safe_dec

    def decompressor(self) -> LZ4FastDecompressor:
        """deprecated alias of fastDecompressor() (LZ4Factory.java:303-311)"""
        return self._fast_dec

    def __str__(self):
        return "LZ4Factory:B200"
