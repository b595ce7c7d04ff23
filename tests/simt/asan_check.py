"""AddressSanitizer pass over the emulated kernels with EXACT-SIZE buffers (manual tool; ~10 minutes):
every buffer is a malloc of its nominal size plus 4 bytes in front and 7 behind — the slack the kernels document
(they read the aligned 32-bit words that contain their first and last byte) — so any wider access is reported.

    python tests/simt/asan_check.py          # builds ASan variants of dec_harness / comp_harness and re-execs itself under libasan

Round 1: 176 corpus items x (valid + 3 mutated streams) x 4 decoder kernels and 1232 compressions over 7 kernel variants with
full and random capacities: no report.  Round 2 (the wide compressor in its three- and two-warp builds, the long-block kernel, the
rewritten decoder walk): see the last line this prints; DESIGN.md quotes it."""
import ctypes, os, random, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(harness, so):
    out = os.path.join(HERE, "_build", so)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-fsanitize=address", "-fno-omit-frame-pointer",
                    "-Wno-unknown-pragmas", "-Wno-attributes", "-DB200_HOST_SIM", "-I" + HERE, "-I" + os.path.join(ROOT, "lz4-java_b200", "csrc"),
                    os.path.join(HERE, harness), "-o", out], check=True)
    return out


def main():
    if "libasan" not in os.environ.get("LD_PRELOAD", ""):
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
        build("dec_harness.cpp", "libdecsim_asan.so"); build("comp_harness.cpp", "libcompsim_asan.so")
        sys.exit(subprocess.run([sys.executable, os.path.abspath(__file__)], env=env).returncode)
    from oracle import oracle as O
    import corpus
    chk = O.Port()
    dec = ctypes.CDLL(os.path.join(HERE, "_build", "libdecsim_asan.so"))
    cmp_ = ctypes.CDLL(os.path.join(HERE, "_build", "libcompsim_asan.so"))
    for f in (dec.sim_decompress_safe, dec.sim_decompress_fast):
        f.restype = ctypes.c_int; f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    cmp_.sim_compress_fast.restype = ctypes.c_int
    cmp_.sim_compress_fast.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    libc = ctypes.CDLL(None); libc.malloc.restype = ctypes.c_void_p; libc.malloc.argtypes = [ctypes.c_size_t]; libc.free.argtypes = [ctypes.c_void_p]

    def decode(c, n, safe, batched):
        sp = libc.malloc(len(c) + 11); dp = libc.malloc(n + 8)
        ctypes.memmove(sp + 4, c, len(c))
        r = (dec.sim_decompress_safe if safe else dec.sim_decompress_fast)(sp + 4, len(c), dp + 4, n, batched)
        out = ctypes.string_at(dp + 4, n) if r >= 0 else b""
        libc.free(sp); libc.free(dp)
        return r, out

    rng = random.Random(1); nd = nc = 0
    variants = {"wide3": 3, "wide2": 2, "long": 0}             # comp_harness.cpp: warps of the wide kernel, 0 = the long-block kernel
    for name, d in corpus.blocks(chk, big=False):
        c = chk.compress(d)
        for b in (1, 0):
            r, o = decode(c, len(d), True, b); assert r == len(d) and o == d, (name, b)
            r, o = decode(c, len(d), False, b); assert r == len(c) and o == d, (name, b)
            for m in corpus.mutate(c, rng, 3):
                if m: decode(m, len(d), True, b)
            nd += 1
        bound = chk.compress_bound(len(d))
        for vn, kind in variants.items():
            if kind and len(d) > 65536:
                continue
            for cap in (bound, rng.randrange(0, bound + 1), -1):
                s_ = libc.malloc(len(d) + 11); d_ = libc.malloc(max(cap, 0) + 8)
                ctypes.memmove(s_ + 4, d, len(d))
                r = cmp_.sim_compress_fast(s_ + 4, len(d), d_ + 4, cap, kind)
                assert 0 <= r <= max(cap, 0)
                if r > 0:
                    rr, o = chk.decompress_safe(ctypes.string_at(d_ + 4, r), len(d)); assert rr == len(d) and o == d, (name, vn)
                libc.free(s_); libc.free(d_); nc += 1
    print(f"asan: {nd} decoder items, {nc} compressions: no report")


if __name__ == "__main__":
    main()
