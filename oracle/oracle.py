"""
oracle/oracle.py — ctypes bindings for the CPU checkers.  TEST INFRASTRUCTURE, NOT PRODUCT.

Two libraries:
  * ``port``  = oracle/liboracle.so      (the plain-C restatements in this directory)
  * ``ref``   = oracle/_ref/liblz4ref.so (the reference's own lz4 1.9.4 / xxHash 0.6.5
                sources compiled in place by oracle/Makefile; present whenever it was
                built in the dev container — the .so travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "liblz4ref.so")

u8p = C.POINTER(C.c_uint8)


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (always) and _ref/liblz4ref.so (when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _as_u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        assert b.dtype == np.uint8
        return np.ascontiguousarray(b)
    return np.frombuffer(bytes(b), dtype=np.uint8)


class _Lib:
    """Common surface over both libraries (names differ, semantics identical)."""

    kind = "?"

    def compress_bound(self, n: int) -> int: ...


class Port(_Lib):
    kind = "port"

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build()
        L = self.L = C.CDLL(PORT_SO)
        L.orc_lz4_compress_bound.restype = C.c_int
        L.orc_lz4_compress_bound.argtypes = [C.c_int]
        for f in (L.orc_lz4_compress_default, L.orc_lz4_decompress_safe):
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_lz4_decompress_fast.restype = C.c_int
        L.orc_lz4_decompress_fast.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_xxh32.restype = C.c_uint32
        L.orc_xxh32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.orc_datagen.restype = None
        L.orc_datagen.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_uint32]
        for nm, st in (("32", C.c_uint32), ("64", C.c_uint64)):
            getattr(L, f"orc_xxh{nm}_state_size").restype = C.c_size_t
            getattr(L, f"orc_xxh{nm}_reset").argtypes = [C.c_void_p, st]
            getattr(L, f"orc_xxh{nm}_update").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
            getattr(L, f"orc_xxh{nm}_digest").argtypes = [C.c_void_p]
            getattr(L, f"orc_xxh{nm}_digest").restype = st
        L.orc_frame_bound.restype = C.c_size_t
        L.orc_frame_bound.argtypes = [C.c_size_t, C.c_int]
        L.orc_frame_compress.restype = C.c_size_t
        L.orc_frame_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.orc_frame_decompress.restype = C.c_int64
        L.orc_frame_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_lz4block_bound.restype = C.c_size_t
        L.orc_lz4block_bound.argtypes = [C.c_size_t, C.c_int]
        L.orc_lz4block_compress.restype = C.c_int64
        L.orc_lz4block_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_lz4block_decompress.restype = C.c_int64
        L.orc_lz4block_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_with_length_compress.restype = C.c_int
        L.orc_with_length_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_with_length_decompress.restype = C.c_int
        L.orc_with_length_decompress.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_cpu_bench.restype = C.c_double
        L.orc_cpu_bench.argtypes = [C.c_int, C.c_void_p] + [C.c_void_p] * 7 + [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]

    # ---- block codec
    def compress_bound(self, n):
        return self.L.orc_lz4_compress_bound(n)

    def compress(self, src, cap: int | None = None) -> bytes:
        s = _as_u8(src)
        cap = self.compress_bound(len(s)) if cap is None else cap
        d = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.L.orc_lz4_compress_default(_ptr(s), _ptr(d), len(s), cap)
        return None if r <= 0 else d[:r].tobytes()

    def decompress_safe(self, src, cap: int):
        """returns (ret, bytes[:max(ret,0)])"""
        s = _as_u8(src)
        d = np.zeros(max(cap, 1), dtype=np.uint8)
        r = self.L.orc_lz4_decompress_safe(_ptr(s), _ptr(d), len(s), cap)
        return r, d[: max(r, 0)].tobytes()

    def decompress_fast(self, src, n: int):
        """returns (ret, bytes[:n]); src is padded so a hostile stream cannot run off the end"""
        s = np.concatenate([_as_u8(src), np.zeros(n + n // 255 + 64, dtype=np.uint8)])
        d = np.zeros(max(n, 1), dtype=np.uint8)
        r = self.L.orc_lz4_decompress_fast(_ptr(s), _ptr(d), n)
        return r, d[:n].tobytes()

    def xxh32(self, buf, seed=0):
        s = _as_u8(buf)
        return self.L.orc_xxh32(_ptr(s), len(s), seed & 0xFFFFFFFF)

    def xxh64(self, buf, seed=0):
        s = _as_u8(buf)
        return self.L.orc_xxh64(_ptr(s), len(s), seed & 0xFFFFFFFFFFFFFFFF)

    def xxh_stream(self, bits: int, chunks, seed=0):
        nm = str(bits)
        st = C.create_string_buffer(getattr(self.L, f"orc_xxh{nm}_state_size")())
        getattr(self.L, f"orc_xxh{nm}_reset")(st, seed)
        for ch in chunks:
            a = _as_u8(ch)
            getattr(self.L, f"orc_xxh{nm}_update")(st, _ptr(a), len(a))
        return getattr(self.L, f"orc_xxh{nm}_digest")(st)

    def datagen(self, size: int, match_proba=0.5, lit_proba=0.0, seed=0) -> np.ndarray:
        out = np.empty(size, dtype=np.uint8)
        self.L.orc_datagen(_ptr(out), size, match_proba, lit_proba, seed)
        return out

    # ---- frame container
    def frame_compress(self, src, bs_code=7, flags=1) -> bytes:
        s = _as_u8(src)
        cap = self.L.orc_frame_bound(len(s), bs_code)
        d = np.empty(cap, dtype=np.uint8)
        scratch = np.empty(self.compress_bound(1 << (8 + 2 * bs_code)), dtype=np.uint8)
        r = self.L.orc_frame_compress(_ptr(s), len(s), _ptr(d), cap, bs_code, flags, _ptr(scratch))
        assert r > 0
        return d[:r].tobytes()

    def frame_decompress(self, src, cap: int):
        s = _as_u8(src)
        d = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.L.orc_frame_decompress(_ptr(s), len(s), _ptr(d), cap)
        return r, d[: max(r, 0)].tobytes()

    # ---- lz4-java's pure-Java backend restated (lz4_java_port_oracle.c; PARITY UNPINNED: no JVM here)
    def java_compress(self, src, cap: int | None = None):
        """LZ4JavaSafeCompressor.compress -> bytes, or None where Java throws LZ4Exception"""
        s = _as_u8(src)
        cap = self.compress_bound(len(s)) if cap is None else cap
        d = np.empty(max(cap, 1), dtype=np.uint8)
        self.L.orc_java_compress.restype = C.c_int
        r = self.L.orc_java_compress(_ptr(s), len(s), _ptr(d), cap)
        return None if r < 0 else d[:r].tobytes()

    def java_decompress_safe(self, src, cap: int):
        """LZ4JavaSafeSafeDecompressor -> (bytes written or -1, output)"""
        s = _as_u8(src)
        d = np.zeros(max(cap, 1) + 8, dtype=np.uint8)
        self.L.orc_java_decompress_safe.restype = C.c_int
        r = self.L.orc_java_decompress_safe(_ptr(s), len(s), _ptr(d), cap)
        return r, d[: max(r, 0)].tobytes()

    def java_decompress_fast(self, src, n: int):
        """LZ4JavaSafeFastDecompressor -> (bytes read or -1, output); len(src) is the readable array length"""
        s = _as_u8(src)
        d = np.zeros(max(n, 1) + 8, dtype=np.uint8)
        self.L.orc_java_decompress_fast.restype = C.c_int
        r = self.L.orc_java_decompress_fast(_ptr(s), len(s), _ptr(d), n)
        return r, d[:n].tobytes() if r >= 0 else b""

    # ---- lz4-java's own containers
    def lz4block_compress(self, src, block_size=65536) -> bytes:
        s = _as_u8(src)
        d = np.empty(self.L.orc_lz4block_bound(len(s), block_size) + 64, dtype=np.uint8)
        scratch = np.empty(self.compress_bound(block_size) + 64, dtype=np.uint8)
        r = self.L.orc_lz4block_compress(_ptr(s), len(s), _ptr(d), block_size, _ptr(scratch))
        return d[:r].tobytes()

    def lz4block_decompress(self, src, cap: int, stop_on_empty_block: bool = True):
        s = _as_u8(src)
        d = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.L.orc_lz4block_decompress(_ptr(s), len(s), _ptr(d), cap, int(stop_on_empty_block))
        return r, d[: max(r, 0)].tobytes()

    def with_length_compress(self, src) -> bytes:
        s = _as_u8(src)
        cap = self.compress_bound(len(s)) + 4
        d = np.empty(cap, dtype=np.uint8)
        r = self.L.orc_with_length_compress(_ptr(s), _ptr(d), len(s), cap)
        return d[:r].tobytes()

    def with_length_decompress(self, src, cap: int):
        s = np.concatenate([_as_u8(src), np.zeros(cap + cap // 255 + 64, dtype=np.uint8)])
        d = np.zeros(max(cap, 1), dtype=np.uint8)
        r = self.L.orc_with_length_decompress(_ptr(s), _ptr(d), cap)
        n = int.from_bytes(bytes(s[:4]), "little") if r >= 0 else 0
        return r, d[:n].tobytes()

    # ---- function addresses for the pthread harness
    def fn_addr(self, op: str) -> int:
        name = {"compress": "orc_lz4_compress_default", "dec_safe": "orc_lz4_decompress_safe",
                "dec_fast": "orc_lz4_decompress_fast", "xxh32": "orc_xxh32", "xxh64": "orc_xxh64"}[op]
        return C.cast(getattr(self.L, name), C.c_void_p).value


class Ref(_Lib):
    """The reference's own C, compiled from /root/reference by oracle/Makefile."""
    kind = "reference"

    def __init__(self):
        if not os.path.exists(REF_SO):
            if os.path.isdir("/root/reference/src/lz4/lib"):
                build()
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO)
        L = self.L = C.CDLL(REF_SO)
        L.LZ4_compressBound.restype = C.c_int
        L.LZ4_compressBound.argtypes = [C.c_int]
        L.LZ4_versionNumber.restype = C.c_int
        for f in (L.LZ4_compress_default, L.LZ4_decompress_safe):
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.LZ4_compress_HC.restype = C.c_int
        L.LZ4_compress_HC.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.LZ4_decompress_fast.restype = C.c_int
        L.LZ4_decompress_fast.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.XXH32.restype = C.c_uint32
        L.XXH32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
        L.XXH64.restype = C.c_uint64
        L.XXH64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.RDG_genBuffer.restype = None
        L.RDG_genBuffer.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_uint]
        L.LZ4F_compressFrameBound.restype = C.c_size_t
        L.LZ4F_compressFrameBound.argtypes = [C.c_size_t, C.c_void_p]
        L.LZ4F_compressFrame.restype = C.c_size_t
        L.LZ4F_compressFrame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.LZ4F_isError.restype = C.c_uint
        L.LZ4F_isError.argtypes = [C.c_size_t]
        L.LZ4F_createDecompressionContext.restype = C.c_size_t
        L.LZ4F_createDecompressionContext.argtypes = [C.POINTER(C.c_void_p), C.c_uint]
        L.LZ4F_freeDecompressionContext.argtypes = [C.c_void_p]
        L.LZ4F_decompress.restype = C.c_size_t
        L.LZ4F_decompress.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p]

    def version(self):
        return self.L.LZ4_versionNumber()

    def compress_bound(self, n):
        return self.L.LZ4_compressBound(n)

    def compress(self, src, cap=None):
        s = _as_u8(src)
        cap = self.compress_bound(len(s)) if cap is None else cap
        d = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.L.LZ4_compress_default(_ptr(s), _ptr(d), len(s), cap)
        return None if r <= 0 else d[:r].tobytes()

    def compress_hc(self, src, level=9):
        s = _as_u8(src)
        cap = self.compress_bound(len(s))
        d = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.L.LZ4_compress_HC(_ptr(s), _ptr(d), len(s), cap, level)
        return None if r <= 0 else d[:r].tobytes()

    def decompress_safe(self, src, cap):
        s = _as_u8(src)
        # guard bytes after dst: the reference promises never to write past cap
        d = np.zeros(max(cap, 1) + 64, dtype=np.uint8)
        r = self.L.LZ4_decompress_safe(_ptr(s), _ptr(d), len(s), cap)
        return r, d[: max(r, 0)].tobytes()

    def decompress_fast(self, src, n):
        s = np.concatenate([_as_u8(src), np.zeros(n + n // 255 + 64, dtype=np.uint8)])
        d = np.zeros(max(n, 1) + 64, dtype=np.uint8)
        r = self.L.LZ4_decompress_fast(_ptr(s), _ptr(d), n)
        return r, d[:n].tobytes()

    def xxh32(self, buf, seed=0):
        s = _as_u8(buf)
        return self.L.XXH32(_ptr(s), len(s), seed & 0xFFFFFFFF)

    def xxh64(self, buf, seed=0):
        s = _as_u8(buf)
        return self.L.XXH64(_ptr(s), len(s), seed & 0xFFFFFFFFFFFFFFFF)

    def datagen(self, size, match_proba=0.5, lit_proba=0.0, seed=0):
        out = np.empty(size, dtype=np.uint8)
        self.L.RDG_genBuffer(_ptr(out), size, match_proba, lit_proba, seed)
        return out

    def frame_compress(self, src, bs_code=7, flags=1) -> bytes:
        """LZ4F_compressFrame with independent blocks (lz4frame.h:175-196 prefs layout)."""
        s = _as_u8(src)

        class FrameInfo(C.Structure):
            _fields_ = [("blockSizeID", C.c_int), ("blockMode", C.c_int), ("contentChecksumFlag", C.c_int),
                        ("frameType", C.c_int), ("contentSize", C.c_ulonglong), ("dictID", C.c_uint),
                        ("blockChecksumFlag", C.c_int)]

        class Prefs(C.Structure):
            _fields_ = [("frameInfo", FrameInfo), ("compressionLevel", C.c_int), ("autoFlush", C.c_uint),
                        ("favorDecSpeed", C.c_uint), ("reserved", C.c_uint * 3)]
        p = Prefs()
        p.frameInfo.blockSizeID = bs_code
        p.frameInfo.blockMode = 1                      # LZ4F_blockIndependent
        p.frameInfo.contentChecksumFlag = 1 if flags & 1 else 0
        p.frameInfo.blockChecksumFlag = 1 if flags & 2 else 0
        p.frameInfo.contentSize = len(s) if flags & 4 else 0
        cap = self.L.LZ4F_compressFrameBound(len(s), C.byref(p))
        d = np.empty(cap, dtype=np.uint8)
        r = self.L.LZ4F_compressFrame(_ptr(d), cap, _ptr(s), len(s), C.byref(p))
        assert not self.L.LZ4F_isError(r)
        return d[:r].tobytes()

    def frame_decompress(self, src, cap):
        s = _as_u8(src)
        d = np.empty(max(cap, 1), dtype=np.uint8)
        ctx = C.c_void_p()
        assert not self.L.LZ4F_isError(self.L.LZ4F_createDecompressionContext(C.byref(ctx), 100))
        ip = op = 0
        try:
            while ip < len(s):
                ssz = C.c_size_t(len(s) - ip)
                dsz = C.c_size_t(cap - op)
                r = self.L.LZ4F_decompress(ctx, C.c_void_p(d.ctypes.data + op), C.byref(dsz),
                                           C.c_void_p(s.ctypes.data + ip), C.byref(ssz), None)
                if self.L.LZ4F_isError(r):
                    return -1, b""
                ip += ssz.value
                op += dsz.value
                if ssz.value == 0 and dsz.value == 0:
                    break
        finally:
            self.L.LZ4F_freeDecompressionContext(ctx)
        return op, d[:op].tobytes()

    def fn_addr(self, op: str) -> int:
        name = {"compress": "LZ4_compress_default", "dec_safe": "LZ4_decompress_safe",
                "dec_fast": "LZ4_decompress_fast", "xxh32": "XXH32", "xxh64": "XXH64"}[op]
        return C.cast(getattr(self.L, name), C.c_void_p).value


_OPS = {"compress": 0, "dec_safe": 1, "dec_fast": 2, "xxh32": 3, "xxh64": 4}


def cpu_bench(lib: _Lib, op: str, src: np.ndarray, src_off, src_len, dst, dst_off, dst_cap,
              threads: int, passes: int = 3):
    """Time ``op`` over a batch of independent blocks on ``threads`` host threads.

    Returns (best_seconds, median_seconds, results[int64])."""
    port = Port()
    n = len(src_off)
    src_off = np.ascontiguousarray(src_off, dtype=np.uint64)
    src_len = np.ascontiguousarray(src_len, dtype=np.int32)
    dst_off = np.ascontiguousarray(dst_off if dst_off is not None else np.zeros(n), dtype=np.uint64)
    dst_cap = np.ascontiguousarray(dst_cap if dst_cap is not None else np.zeros(n), dtype=np.int32)
    res = np.zeros(n, dtype=np.int64)
    med = C.c_double()
    best = port.L.orc_cpu_bench(_OPS[op], C.c_void_p(lib.fn_addr(op)), _ptr(src), _ptr(src_off), _ptr(src_len),
                                _ptr(dst) if dst is not None else None, _ptr(dst_off), _ptr(dst_cap), _ptr(res),
                                n, threads, passes, C.byref(med))
    return best, med.value, res


def best_available() -> _Lib:
    """The reference build when present, else the port."""
    try:
        return Ref()
    except (FileNotFoundError, OSError):
        return Port()
