"""Generates tests/golden/kat.json from the REFERENCE's own C (oracle/_ref/liblz4ref.so, built from
/root/reference by oracle/Makefile).  Run once in the dev container:  python tests/golden/make_golden.py
The fixture pins the oracle restatements (and, on the GPU box, the CUDA path) to reference outputs
without needing /root/reference at test time."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O      # noqa: E402
import corpus                       # noqa: E402


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    R = O.Ref()
    assert R.version() == 10904, "fixtures must come from lz4 1.9.4"
    kat = {"lz4_version": R.version(), "datagen": [], "xxh": [], "compress": [], "malformed_safe": [], "malformed_fast": []}
    for size, mp, seed in [(0, 0.5, 0), (1, 0.5, 0), (100, 0.5, 3), (70000, 0.5, 0), (300000, 0.2, 7), (300000, 0.8, 9), (1 << 20, 0.5, 2)]:
        kat["datagen"].append({"size": size, "match_proba": mp, "seed": seed, "sha256": sha(R.datagen(size, mp, 0.0, seed))})
    stream = R.datagen(200000, 0.5, 0.0, 77).tobytes()
    for n in list(range(0, 70)) + [255, 256, 4095, 4096, 4097, 65536, 100001]:
        for seed in (0, 0x9747B28C):
            kat["xxh"].append({"len": n, "seed": seed, "xxh32": R.xxh32(stream[:n], seed), "xxh64": R.xxh64(stream[:n], seed * 0x100000001)})
    for name, d in corpus.blocks(R):
        c = R.compress(d)
        kat["compress"].append({"name": name, "len": len(d), "in_sha256": sha(d), "clen": len(c), "c_sha256": sha(c)})
    for v in corpus.MALFORMED:
        for cap in (0, 12, 20, 64, 100, 200):
            kat["malformed_safe"].append({"hex": v.hex(), "cap": cap, "ret": R.decompress_safe(v, cap)[0]})
        for n in (13, 20, 64):
            kat["malformed_fast"].append({"hex": v.hex(), "n": n, "ret": R.decompress_fast(v, n)[0]})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json")
    json.dump(kat, open(out, "w"), indent=0)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
