"""Digests of the fast compressor's output on the seeded test corpus (tests/corpus.py), written by the CPU emulator build
of the kernel source (tests/simt) at a point where that output had been checked byte for byte against the round-1 kernel
it replaced (lz4_compress_fast3_kernel, itself validated on the GPU against oracle/_ref's decoders).  The stream is a
*different valid parse* than the reference's, so there is no reference-held vector for it: this fixture pins "the same
bytes as before" across refactors, and "the GPU emits what the emulator emits" (tests/test_gpu_parity.py).
    python tests/golden/make_fast_streams.py > tests/golden/fast_streams.json"""
import ctypes, hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
import corpus
from oracle import oracle as O


def main():
    port = O.best_available()
    lib = ctypes.CDLL(os.path.join(os.path.dirname(HERE), "simt", "_build", "libcompsim.so"))
    lib.sim_compress_fast.restype = ctypes.c_int
    lib.sim_compress_fast.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    out = {}
    for name, d in corpus.blocks(port):
        if len(d) >= 65536 + 11:
            continue
        cap = port.compress_bound(len(d))
        a = np.zeros(len(d) + 8192, dtype=np.uint8); a[4096:4096 + len(d)] = np.frombuffer(d, dtype=np.uint8)
        o = np.zeros(cap + 64, dtype=np.uint8)
        r = lib.sim_compress_fast(a.ctypes.data + 4096, len(d), o.ctypes.data, cap, 3)
        assert r > 0 and port.decompress_safe(o[:r].tobytes(), len(d)) == (len(d), d), name
        out[name] = {"n": len(d), "c": r, "sha256": hashlib.sha256(o[:r].tobytes()).hexdigest()}
    print(json.dumps({"kernel": "lz4_compress_wide_kernel<13>", "note": __doc__.split("\n")[0], "streams": out}, indent=0))


if __name__ == "__main__":
    main()
