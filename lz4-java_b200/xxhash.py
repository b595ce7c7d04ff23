"""Host-side mirror of net.jpountz.xxhash over libb200lz4.

  XXHashFactory.b200Instance()   <- XXHashFactory.nativeInstance()   XXHashFactory.java:91-96
  XXHash32.hash / XXHash64.hash  <- XXHash32JNI / XXHash64JNI        XXHash32JNI.java:29-49
  StreamingXXHash32 / 64         <- StreamingXXHash32JNI.java:26-91 (long state handle, close())
"""
from __future__ import annotations

from . import _native as N
from .lz4 import _view, _check_range, _addr


class XXHash32:
    def hash(self, buf, off: int = 0, length: int | None = None, seed: int = 0) -> int:
        a = _view(buf)
        if length is None:
            length = len(a) - off
        _check_range(a, off, length)
        return N.checked_value(N.lib().b200xxh32(_addr(a, off) if length else None, length, seed & 0xFFFFFFFF))


class XXHash64:
    def hash(self, buf, off: int = 0, length: int | None = None, seed: int = 0) -> int:
        a = _view(buf)
        if length is None:
            length = len(a) - off
        _check_range(a, off, length)
        return N.checked_value(N.lib().b200xxh64(_addr(a, off) if length else None, length, seed & 0xFFFFFFFFFFFFFFFF))


class _Streaming:
    _bits = 32

    def __init__(self, seed: int):
        self.seed = seed
        L = N.lib()
        self._L = L
        mask = 0xFFFFFFFF if self._bits == 32 else 0xFFFFFFFFFFFFFFFF
        self._state = getattr(L, f"b200xxh{self._bits}_create")(seed & mask)
        if not self._state:
            raise N.B200Error("libb200lz4: " + N.last_error())

    def _check(self):
        if not self._state:
            raise AssertionError("Already finalized")            # StreamingXXHash32JNI.java:41-45

    def reset(self):
        self._check()
        mask = 0xFFFFFFFF if self._bits == 32 else 0xFFFFFFFFFFFFFFFF
        getattr(self._L, f"b200xxh{self._bits}_reset")(self._state, self.seed & mask)

    def update(self, buf, off: int = 0, length: int | None = None):
        self._check()
        a = _view(buf)
        if length is None:
            length = len(a) - off
        _check_range(a, off, length)
        if length:
            N.check(getattr(self._L, f"b200xxh{self._bits}_update")(self._state, _addr(a, off), length))

    def getValue(self) -> int:
        self._check()
        return N.checked_value(getattr(self._L, f"b200xxh{self._bits}_digest")(self._state))

    def close(self):
        if self._state:
            getattr(self._L, f"b200xxh{self._bits}_free")(self._state)
            self._state = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StreamingXXHash32(_Streaming):
    _bits = 32

    def asChecksum(self) -> int:
        """java.util.zip.Checksum view masks to 28 bits (StreamingXXHash32.java:106)."""
        return self.getValue() & 0xFFFFFFF


class StreamingXXHash64(_Streaming):
    _bits = 64


class XXHashFactory:
    _instance = None

    def __init__(self):
        self._h32, self._h64 = XXHash32(), XXHash64()
        # XXHashFactory.java:184-203: one-shot must equal streaming on 100 bytes
        import random
        r = random.Random(0)
        data = bytes(r.randrange(256) for _ in range(100))
        seed = r.randrange(1 << 31)
        for one, stream in ((self._h32, StreamingXXHash32(seed)), (self._h64, StreamingXXHash64(seed))):
            stream.update(data, 0, len(data))
            if one.hash(data, 0, len(data), seed) != stream.getValue():
                raise AssertionError("XXHash self-test failed")
            stream.close()

    @classmethod
    def b200Instance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def hash32(self):
        return self._h32

    def hash64(self):
        return self._h64

    def newStreamingHash32(self, seed: int):
        return StreamingXXHash32(seed)

    def newStreamingHash64(self, seed: int):
        return StreamingXXHash64(seed)
