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


def decompress_frames(src, max_decoded: int) -> bytes:
    """decode every frame in `src` (concatenated / skippable frames allowed) -> the decoded stream"""
    s = _view(src)
    out = np.empty(max(max_decoded, 1), dtype=np.uint8)
    r = N.lib().b200lz4f_decompress_host(s.ctypes.data, len(s), out.ctypes.data, max_decoded)
    N.check(r)
    if r < 0:
        raise LZ4FrameError(int(r))
    return out[:r].tobytes()
