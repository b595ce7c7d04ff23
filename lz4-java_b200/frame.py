"""LZ4 Frame batch decoding on the B200 backend — the semantics of LZ4FrameInputStream.read()
(src/java/net/jpountz/lz4/LZ4FrameInputStream.java:132-321) for whole buffers of concatenated frames.

The reference's stream class decodes one block per native call; here one call indexes the container on
the host and decodes every block of every frame in batched GPU launches, verifying the header, block
and content XXH32 checksums on the device."""
from __future__ import annotations

import numpy as np

from . import _native as N
from .lz4 import _view

ERRORS = {
    -1: "Stream ended prematurely",                 # LZ4FrameInputStream.PREMATURE_EOS
    -2: "Stream unsupported (invalid magic bytes)",  # NOT_SUPPORTED
    -3: "Stream frame descriptor corrupted",         # DESCRIPTOR_HASH_MISMATCH
    -4: "Block size exceeded max",
    -5: "Block checksum mismatch",                   # BLOCK_HASH_MISMATCH
    -6: "Error decoding block",
    -7: "Content checksum mismatch",
    -8: "Size check mismatch",
    -9: "destination too small",
    -10: "unsupported frame descriptor",
}


class LZ4FrameError(IOError):
    """what LZ4FrameInputStream throws as java.io.IOException"""

    def __init__(self, code: int):
        super().__init__(ERRORS.get(code, f"B200 backend error {code}"))
        self.code = code


def decompress_frames(src, max_decoded: int, read_single_frame: bool = False) -> bytes:
    """decode every frame in `src` (concatenated / skippable frames allowed) -> the decoded stream;
    read_single_frame: stop behind the first non-skippable frame like LZ4FrameInputStream(in, true) (:83-91)"""
    s = _view(src)
    out = np.empty(max(max_decoded, 1), dtype=np.uint8)
    if read_single_frame:
        r = N.lib().b200lz4f_decompress_host_single(s.ctypes.data, len(s), out.ctypes.data, max_decoded, None)
    else:
        r = N.lib().b200lz4f_decompress_host(s.ctypes.data, len(s), out.ctypes.data, max_decoded)
    N.check(r)
    if r < 0:
        raise LZ4FrameError(int(r))
    return out[:r].tobytes()


def expected_content_size(src) -> int:
    """LZ4FrameInputStream.getExpectedContentSize (:416-428): the content size the first non-skippable frame declares, -1 if none"""
    import ctypes
    s = _view(src)
    size = ctypes.c_int64(-1)
    r = N.lib().b200lz4f_expected_content_size(s.ctypes.data, len(s), ctypes.byref(size))
    N.check(r)
    if r < 0:
        raise LZ4FrameError(int(r))
    return int(size.value)


def compress_frame(src, block_size_code: int = 4, content_checksum=True, block_checksum=False, content_size=False, hc_level: int = 0) -> bytes:
    """one LZ4 frame as LZ4FrameOutputStream writes it (LZ4FrameOutputStream.java:178-251), whole buffer at once;
    hc_level: the stream's compressor argument (:132-133) -- 0 = fastCompressor(), 1..17 = highCompressor(level)"""
    s = _view(src)
    flags = (1 if content_checksum else 0) | (2 if block_checksum else 0) | (4 if content_size else 0)
    L = N.lib()
    cap = L.b200lz4f_compress_bound(len(s), block_size_code)
    if cap == 0:
        raise ValueError("block_size_code must be 4..7 (64 KiB .. 4 MiB)")
    out = np.empty(cap, dtype=np.uint8)
    r = L.b200lz4f_compress_host_hc(s.ctypes.data, len(s), out.ctypes.data, cap, block_size_code, flags, hc_level)
    N.check(r)
    if r < 0:
        raise LZ4FrameError(int(r))
    return out[:r].tobytes()


# ---- lz4-java's private "LZ4Block" container (LZ4BlockOutputStream / LZ4BlockInputStream)
def compress_lz4block(src, block_size: int = 1 << 16, hc_level: int = 0) -> bytes:
    s = _view(src)
    L = N.lib()
    cap = L.b200lz4block_compress_bound(len(s), block_size)
    if cap == 0:
        raise ValueError("blockSize must be >= 64 and <= 32 MiB")            # LZ4BlockOutputStream.java:58-66
    out = np.empty(cap, dtype=np.uint8)
    r = L.b200lz4block_compress_host_hc(s.ctypes.data, len(s), out.ctypes.data, cap, block_size, hc_level)
    N.check(r)
    if r < 0:
        raise LZ4FrameError(int(r))
    return out[:r].tobytes()


def decompress_lz4block(src, max_decoded: int, stop_on_empty_block: bool = True) -> bytes:
    """LZ4BlockInputStream(in, stopOnEmptyBlock) read to its end (LZ4BlockInputStream.java:60-72, default true :100-104)."""
    s = _view(src)
    out = np.empty(max(max_decoded, 1), dtype=np.uint8)
    r = N.lib().b200lz4block_decompress_host(s.ctypes.data, len(s), out.ctypes.data, max_decoded, int(bool(stop_on_empty_block)), None)
    N.check(r)
    if r == -1:
        raise EOFError("Stream ended prematurely")                           # LZ4BlockInputStream.java:197
    if r < 0:
        raise IOError("Stream is corrupted" if r == -2 else f"error {r}")    # LZ4BlockInputStream.java:203,...
    return out[:r].tobytes()


# ---- LZ4CompressorWithLength / LZ4DecompressorWithLength
def compress_with_length(src) -> bytes:
    s = _view(src)
    cap = len(s) + len(s) // 255 + 16 + 4
    out = np.empty(cap, dtype=np.uint8)
    r = N.lib().b200lz4_compress_with_length(s.ctypes.data if len(s) else None, out.ctypes.data, len(s), cap)
    N.check(r)
    if r <= 0:
        from .lz4 import LZ4Exception
        raise LZ4Exception("maxDestLen is too small")
    return out[:r].tobytes()


def decompress_with_length(src) -> bytes:
    s = _view(src)
    from .lz4 import LZ4Exception
    if len(s) < 4:
        raise LZ4Exception("Error decoding offset 0 of input buffer")
    n = N.lib().b200lz4_decompressed_length(s.ctypes.data)
    if n < 0:
        raise LZ4Exception("negative length")
    out = np.empty(max(n, 1), dtype=np.uint8)
    r = N.lib().b200lz4_decompress_with_length(s.ctypes.data, len(s), out.ctypes.data, n)
    N.check(r)
    if r < 0:
        raise LZ4Exception("Error decoding offset " + str(-r) + " of input buffer")
    return out[:n].tobytes()
