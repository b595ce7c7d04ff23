"""Diagnostic for a round-trip mismatch seen once on RDG P=0.80 (tools/corpus_sweep.py): repeat the sweep's exact
procedure, attribute every bad block to the compressor (CPU checker rejects the stream) or to a decoder."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import _variant  # noqa: F401  (B200LZ4_TEST_SO development switch)
import lz4java_b200 as L
from oracle import oracle as O
BS = 65536
chk = O.best_available(); dev = torch.device("cuda:0"); B = L.batch; lib = L._native.lib()
nblk = 16384; base_n = 4096
bound = L.max_compressed_length(BS); stride = (bound + 15) // 16 * 16
soff = torch.arange(nblk, device=dev, dtype=torch.int64) * BS
slen = torch.full((nblk,), BS, device=dev, dtype=torch.int32)
coff = torch.arange(nblk, device=dev, dtype=torch.int64) * stride
ccap = torch.full((nblk,), bound, device=dev, dtype=torch.int32)
comp = torch.zeros(nblk * stride, device=dev, dtype=torch.uint8)
clen = torch.zeros(nblk, device=dev, dtype=torch.int32)
out = torch.zeros(nblk * BS, device=dev, dtype=torch.uint8)
res = torch.zeros(nblk, device=dev, dtype=torch.int32)
knob = ctypes.c_int.in_dll(lib, "b200lz4_decompress_batch_below")
def parse(c, upto_out):
    ip = 0; op = 0; seqs = []
    while ip < len(c):
        t0 = ip; tok = int(c[ip]); ip += 1; lit = tok >> 4
        if lit == 15:
            while True:
                s = int(c[ip]); ip += 1; lit += s
                if s != 255: break
        ip += lit; op_l = op; op += lit
        if ip >= len(c): seqs.append((t0, lit, 0, 0, op_l, op)); break
        off = int(c[ip]) | (int(c[ip + 1]) << 8); ip += 2; ml = tok & 15
        if ml == 15:
            while True:
                s = int(c[ip]); ip += 1; ml += s
                if s != 255: break
        ml += 4
        seqs.append((t0, lit, ml, off, op_l, op)); op += ml
        if op_l > upto_out + 400: break
    return seqs
for mp in (0.5, 0.8):
    host = np.ascontiguousarray(chk.datagen(base_n * BS, mp, 0.0, 2))
    src = torch.from_numpy(host).to(dev).repeat(nblk // base_n).contiguous()
    v = src.view(nblk, BS); idx = torch.arange(nblk, device=dev, dtype=torch.int64)
    for k in range(4): v[:, k] ^= ((idx >> (8 * k)) & 0xFF).to(torch.uint8)
    hsrc = src.cpu().numpy()
    for rep in range(3):
        for _ in range(3): B.compress_fast_batch_dev(src, soff, slen, comp, coff, ccap, clen, BS)
        torch.cuda.synchronize()
        for mode, below in (("batched", 1 << 30), ("sequential", 0)):
            knob.value = below
            for kind in ("safe", "fast"):
                out.zero_()
                if kind == "safe": B.decompress_safe_batch_dev(comp, coff, clen, out, soff, slen, res)
                else: B.decompress_fast_batch_dev(comp, coff, ccap, out, soff, slen, res)
                torch.cuda.synchronize()
                eq = (out.view(nblk, BS) == src.view(nblk, BS)).all(dim=1)
                rbad = int((res != BS).sum().item()) if kind == "safe" else int((res != clen).sum().item())
                badb = torch.nonzero(~eq).flatten().cpu().numpy()
                print(f"P={mp} rep {rep} {mode} {kind}: {len(badb)} blocks differ, {rbad} wrong return values; first: {badb[:6]}", flush=True)
                if len(badb) and mode == "batched" and kind == "safe":
                    for b in [int(x) for x in badb[:2]]:
                        cl = int(clen[b].item())
                        c = comp[b * stride: b * stride + cl].cpu().numpy()
                        r, o = chk.decompress_safe(c.tobytes(), BS)
                        cpu_ok = (r == BS and o == hsrc[b * BS:(b + 1) * BS].tobytes())
                        og = out[b * BS:(b + 1) * BS].cpu().numpy(); s = hsrc[b * BS:(b + 1) * BS]
                        d = np.nonzero(og != s)[0]
                        print(f"  block {b}: CPU checker decodes the GPU stream correctly: {cpu_ok} (r={r}); GPU first diff at {int(d[0])}, ndiff {len(d)}, last {int(d[-1])}, res {int(res[b].item())}, clen {cl}")
                        print("   got ", og[d[0] - 4:d[0] + 20].tolist()); print("   want", s[d[0] - 4:d[0] + 20].tolist())
                        if cpu_ok:
                            for q in [q for q in parse(c, int(d[0])) if q[5] + q[2] >= d[0] - 200 and q[4] <= d[0] + 60]:
                                print("    tok@%d lit=%d ml=%d off=%d  lit_out=%d match_out=%d..%d" % (q[0], q[1], q[2], q[3], q[4], q[5], q[5] + q[2]))
                        else:
                            o = np.frombuffer(o, dtype=np.uint8) if r > 0 else None
                            if o is not None and len(o) == BS:
                                dd = np.nonzero(o != s)[0]; print("   CPU decode of the GPU stream differs from the source first at", int(dd[0]))
                                for q in [q for q in parse(c, int(dd[0])) if q[5] + q[2] >= dd[0] - 100 and q[4] <= dd[0] + 40]:
                                    print("    tok@%d lit=%d ml=%d off=%d  lit_out=%d match_out=%d..%d" % (q[0], q[1], q[2], q[3], q[4], q[5], q[5] + q[2]))
    del src
