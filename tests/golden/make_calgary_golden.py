"""Generates tests/golden/calgary_lz4.json in the dev container:  python tests/golden/make_calgary_golden.py
Real-data fixtures that travel to the GPU box (where /root/reference does not exist): 64 KiB cuts of the reference's
Calgary test files (src/test-resources/calgary/{book1,geo,pic}, the files LZ4Test.java round-trips), stored as the
REFERENCE's own output for them — LZ4_compress_default (fast) and LZ4_compress_HC level 9 from oracle/_ref — with the
sha256 of the original bytes.  A decoder under test must turn each stream back into bytes with that digest."""
import base64
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O      # noqa: E402

BASE = "/root/reference/src/test-resources/calgary"


def main():
    R = O.Ref()
    assert R.version() == 10904
    out = {"lz4_version": R.version(), "blocks": []}
    for f, cuts in (("book1", (0, 7)), ("geo", (0, 1)), ("pic", (0, 5))):
        data = open(os.path.join(BASE, f), "rb").read()
        for k in cuts:
            d = data[k * 65536:(k + 1) * 65536]
            c, h = R.compress(d), R.compress_hc(d, 9)
            assert R.decompress_safe(c, len(d))[1] == d and R.decompress_safe(h, len(d))[1] == d
            out["blocks"].append({"name": f"{f}@{k * 65536}", "len": len(d), "sha256": hashlib.sha256(d).hexdigest(),
                                  "fast_b64": base64.b64encode(c).decode(), "hc9_b64": base64.b64encode(h).decode()})
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "calgary_lz4.json")
    json.dump(out, open(p, "w"), indent=0)
    print("wrote", p, os.path.getsize(p), "bytes", [(b["name"], b["len"], len(b["fast_b64"]) * 3 // 4, len(b["hc9_b64"]) * 3 // 4) for b in out["blocks"]])


if __name__ == "__main__":
    main()
