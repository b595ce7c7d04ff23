/*
 * oracle/lz4_block_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A plain-C restatement of the LZ4 block algorithms the reference's JNI backend
 * executes (lz4 1.9.4 vendored at /root/reference/src/lz4/lib).  It exists so the
 * CUDA kernels have a readable CPU checker that travels to the GPU box; it is
 * only ever imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  The product (lz4-java_b200/) never
 * links or calls anything in this directory.
 *
 * Parity is PINNED: tests/test_oracle_pin.py checks every function here against
 * the reference's own C sources compiled in place (oracle/_ref/liblz4ref.so) and
 * against the committed fixtures in tests/golden/ that were produced by that
 * library (tests/golden/make_golden.py).
 *
 * Functions and the reference lines they restate:
 *   orc_lz4_compress_bound      lz4.h:212  LZ4_COMPRESSBOUND / lz4.c:709 LZ4_compressBound
 *   orc_lz4_compress_default    lz4.c:1435 -> 1416 -> 1346 -> 1308 -> 910-1302
 *                               (LZ4_compress_generic_validated, noDict, acceleration 1,
 *                                byU16 below 65547 input bytes, byU32+hash5 otherwise)
 *   orc_lz4_decompress_safe     lz4.c:2345 -> 1936-2339 (decode_full_block, noDict)
 *   orc_lz4_decompress_fast     lz4.c:2362 -> 1794-1891 (LZ4_decompress_unsafe_generic)
 *
 * Byte-for-byte the same results as the reference are expected from all four
 * (including negative return codes of the decoders); the compressor restatement
 * reproduces the reference's exact parse so ratio comparisons are against the
 * reference's own bytes.
 */
#include <stdint.h>
#include <string.h>
#include <stddef.h>

#define ORC_MINMATCH      4
#define ORC_MFLIMIT       12
#define ORC_LASTLITERALS  5
#define ORC_MAX_INPUT     0x7E000000
#define ORC_MAX_DISTANCE  65535
#define ORC_64K_LIMIT     (65536 + (ORC_MFLIMIT - 1))   /* lz4.c:689 */
#define ORC_SKIP_TRIGGER  6                             /* lz4.c:690 */

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

int orc_lz4_compress_bound(int n)
{   /* lz4.h:212 */
    if (n < 0 || (unsigned)n > (unsigned)ORC_MAX_INPUT) return 0;
    return n + n / 255 + 16;
}

/* ------------------------------------------------------------------ compress */

typedef struct { int is16; uint16_t t16[8192]; uint32_t t32[4096]; } orc_table;

static uint32_t tbl_hash(const orc_table* t, const uint8_t* p)
{
    if (t->is16)                       /* lz4.c:756-762, hashLog+1 = 13 bits */
        return (rd32(p) * 2654435761U) >> (32 - 13);
    /* lz4.c:764-780: 64-bit build hashes 5 bytes into 12 bits */
    return (uint32_t)(((rd64(p) << 24) * 889523592379ULL) >> (64 - 12));
}
static uint32_t tbl_get(const orc_table* t, uint32_t h) { return t->is16 ? t->t16[h] : t->t32[h]; }
static void tbl_put(orc_table* t, uint32_t h, uint32_t idx)
{ if (t->is16) t->t16[h] = (uint16_t)idx; else t->t32[h] = idx; }

static unsigned count_common(const uint8_t* a, const uint8_t* b, const uint8_t* alimit)
{   /* lz4.c:659-682 expressed bytewise: number of equal bytes, a stops at alimit */
    const uint8_t* s = a;
    while (a < alimit && *a == *b) { a++; b++; }
    return (unsigned)(a - s);
}

int orc_lz4_compress_default(const uint8_t* src, uint8_t* dst, int srcSize, int dstCap)
{
    static __thread orc_table tbl;                     /* the 16 KiB LZ4_stream_t of lz4.c:1423 */
    int limited;
    const uint8_t *ip, *anchor, *iend, *mfl1, *mlimit, *base;
    uint8_t *op, *olimit, *token;
    uint32_t fwdH;

    if (srcSize < 0 || (unsigned)srcSize > (unsigned)ORC_MAX_INPUT) return 0;   /* lz4.c:1324 */
    limited = dstCap < orc_lz4_compress_bound(srcSize);                         /* lz4.c:1352 */
    if (srcSize == 0) {                                                         /* lz4.c:1325-1336 */
        if (limited && dstCap <= 0) return 0;
        dst[0] = 0; return 1;
    }
    memset(&tbl, 0, sizeof tbl);
    tbl.is16 = srcSize < ORC_64K_LIMIT;                                         /* lz4.c:1353 */

    base = src; ip = src; anchor = src; iend = src + srcSize;
    mfl1 = iend - ORC_MFLIMIT + 1; mlimit = iend - ORC_LASTLITERALS;
    op = dst; olimit = dst + dstCap;

    if (srcSize < ORC_MFLIMIT + 1) goto last_literals;                          /* lz4.c:981 */

    tbl_put(&tbl, tbl_hash(&tbl, ip), 0);                                       /* lz4.c:984 */
    ip++; fwdH = tbl_hash(&tbl, ip);

    for (;;) {
        const uint8_t* match;
        {   /* find a match: lz4.c:1044-1075 */
            const uint8_t* fwdIp = ip;
            int step = 1, nb = 1 << ORC_SKIP_TRIGGER;
            for (;;) {
                uint32_t h = fwdH, cur = (uint32_t)(fwdIp - base), mi = tbl_get(&tbl, h);
                ip = fwdIp; fwdIp += step; step = nb++ >> ORC_SKIP_TRIGGER;
                if (fwdIp > mfl1) goto last_literals;
                match = base + mi;
                fwdH = tbl_hash(&tbl, fwdIp);
                tbl_put(&tbl, h, cur);
                if (!tbl.is16 && mi + ORC_MAX_DISTANCE < cur) continue;          /* lz4.c:1064 */
                if (rd32(match) == rd32(ip)) break;
            }
        }
        /* catch up: lz4.c:1080 */
        while (ip > anchor && match > src && ip[-1] == match[-1]) { ip--; match--; }

        {   /* literals: lz4.c:1083-1104 */
            unsigned lit = (unsigned)(ip - anchor);
            token = op++;
            if (limited && op + lit + (2 + 1 + ORC_LASTLITERALS) + lit / 255 > olimit) return 0;
            if (lit >= 15) {
                int len = (int)lit - 15;
                *token = 15 << 4;
                for (; len >= 255; len -= 255) *op++ = 255;
                *op++ = (uint8_t)len;
            } else *token = (uint8_t)(lit << 4);
            memcpy(op, anchor, lit); op += lit;
        }
    next_match:
        {   unsigned off = (unsigned)(ip - match), mc;
            op[0] = (uint8_t)off; op[1] = (uint8_t)(off >> 8); op += 2;         /* lz4.c:1133 */
            mc = count_common(ip + ORC_MINMATCH, match + ORC_MINMATCH, mlimit);  /* lz4.c:1153 */
            ip += mc + ORC_MINMATCH;
            if (limited && op + (1 + ORC_LASTLITERALS) + (mc + 240) / 255 > olimit) return 0;
            if (mc >= 15) {                                                     /* lz4.c:1184-1196 */
                *token += 15; mc -= 15;
                while (mc >= 255) { *op++ = 255; mc -= 255; }
                *op++ = (uint8_t)mc;
            } else *token += (uint8_t)mc;
        }
        anchor = ip;
        if (ip >= mfl1) break;                                                  /* lz4.c:1204 */
        tbl_put(&tbl, tbl_hash(&tbl, ip - 2), (uint32_t)(ip - 2 - base));       /* lz4.c:1207 */
        {   /* immediate re-probe: lz4.c:1220-1258 */
            uint32_t h = tbl_hash(&tbl, ip), cur = (uint32_t)(ip - base), mi = tbl_get(&tbl, h);
            match = base + mi;
            tbl_put(&tbl, h, cur);
            if ((tbl.is16 || mi + ORC_MAX_DISTANCE >= cur) && rd32(match) == rd32(ip)) {
                token = op++; *token = 0; goto next_match;
            }
        }
        fwdH = tbl_hash(&tbl, ++ip);                                            /* lz4.c:1262 */
    }

last_literals:
    {   size_t run = (size_t)(iend - anchor);                                   /* lz4.c:1266-1293 */
        if (limited && op + run + 1 + (run + 255 - 15) / 255 > olimit) return 0;
        if (run >= 15) {
            size_t acc = run - 15;
            *op++ = 15 << 4;
            for (; acc >= 255; acc -= 255) *op++ = 255;
            *op++ = (uint8_t)acc;
        } else *op++ = (uint8_t)(run << 4);
        memcpy(op, anchor, run); op += run;
    }
    return (int)(op - dst);
}

/* ---------------------------------------------------------------- decompress */

/* read_variable_length, lz4.c:1903-1928: 255-chain bounded by ilimit.
 * Returns -1 on error; *ipp is left exactly where the reference leaves it. */
static int64_t rd_varlen(const uint8_t* src, int64_t* ipp, int64_t ilimit, int initial_check)
{
    int64_t ip = *ipp, len = 0; unsigned s;
    if (initial_check && ip >= ilimit) return -1;
    do {
        s = src[ip]; ip++; len += s;
        if (ip > ilimit) { *ipp = ip; return -1; }
    } while (s == 255);
    *ipp = ip; return len;
}

/*
 * LZ4_decompress_safe.  The reference decoder has two loops (a "fast" loop while
 * op is at least 64 bytes from the end of dst, lz4.c:1996-2109, then the "safe"
 * loop lz4.c:2114-2329) whose accept/reject rules differ slightly on malformed
 * input (the shortcut paths skip the end-of-block literal rule); both are mirrored
 * as decision logic so return codes match, while the copies themselves are plain
 * byte loops (the reference's wild copies only ever scribble inside dst capacity).
 */
int orc_lz4_decompress_safe(const uint8_t* src, uint8_t* dst, int srcSize, int dstCap)
{
    int64_t ip = 0, op = 0, iend = srcSize, oend = dstCap, len, off, cpy;
    unsigned token;
    int fastloop;

    if (src == NULL || dstCap < 0) return -1;
    if (dstCap == 0) return (srcSize == 1 && src[0] == 0) ? 0 : -1;              /* lz4.c:1978-1982 */
    if (srcSize <= 0) return -1;

    fastloop = (oend - op) >= 64;                                               /* lz4.c:1989 */
    for (;;) {
        int have_match_info = 0;
        token = src[ip++];
        len = token >> 4;
        if (fastloop) {
            if (len == 15) {                                                    /* lz4.c:2003-2015 */
                int64_t add = rd_varlen(src, &ip, iend - 15, 1);
                if (add < 0) goto error;
                len += add;
                cpy = op + len;
                if (cpy > oend - 32 || ip + len > iend - 32) { fastloop = 0; goto safe_literal_copy; }
            } else {
                cpy = op + len;                                                 /* lz4.c:2017-2023 */
                if (ip > iend - 17) { fastloop = 0; goto safe_literal_copy; }
            }
            memmove(dst + op, src + ip, (size_t)len); ip += len; op = cpy;
            off = src[ip] | (src[ip + 1] << 8); ip += 2;                        /* lz4.c:2027 */
            len = token & 15;
            if (len == 15) {
                int64_t add = rd_varlen(src, &ip, iend - ORC_LASTLITERALS + 1, 0);
                if (add < 0) goto error;
                len += add + ORC_MINMATCH;
                if (op - off < 0) goto error;                                   /* lz4.c:2041 */
                if (op + len >= oend - 64) { fastloop = 0; goto safe_match_copy; }
            } else {
                len += ORC_MINMATCH;
                if (op + len >= oend - 64) { fastloop = 0; goto safe_match_copy; }
            }
            if (op - off < 0) goto error;                                       /* lz4.c:2066 */
            goto do_match_copy;
        }

        /* safe loop */
        if (len != 15 && ip < iend - 16 && op <= oend - 32) {                   /* shortcut, lz4.c:2128-2160 */
            memmove(dst + op, src + ip, (size_t)len); op += len; ip += len;
            len = token & 15;
            off = src[ip] | (src[ip + 1] << 8); ip += 2;
            if (len != 15 && off >= 8 && op - off >= 0) {
                len += ORC_MINMATCH;
                goto do_match_copy;
            }
            have_match_info = 1;
            goto copy_match;
        }
        if (len == 15) {                                                        /* lz4.c:2163-2169 */
            int64_t add = rd_varlen(src, &ip, iend - 15, 1);
            if (add < 0) goto error;
            len += add;
        }
        cpy = op + len;
    safe_literal_copy:
        if (cpy > oend - ORC_MFLIMIT || ip + len > iend - (2 + 1 + ORC_LASTLITERALS)) {
            /* must be the last sequence: lz4.c:2175-2213 */
            if (ip + len != iend || cpy > oend) goto error;
            memmove(dst + op, src + ip, (size_t)len);
            ip += len; op += len;
            break;
        }
        memmove(dst + op, src + ip, (size_t)len); ip += len; op = cpy;
        off = src[ip] | (src[ip + 1] << 8); ip += 2;                            /* lz4.c:2229 */
        len = token & 15;
    copy_match:
        (void)have_match_info;
        if (len == 15) {                                                        /* lz4.c:2236-2241 */
            int64_t add = rd_varlen(src, &ip, iend - ORC_LASTLITERALS + 1, 0);
            if (add < 0) goto error;
            len += add;
        }
        len += ORC_MINMATCH;
    safe_match_copy:
        if (op - off < 0) goto error;                                           /* lz4.c:2247 */
        cpy = op + len;
        if (cpy > oend - ORC_MFLIMIT) {                                         /* lz4.c:2313-2317 */
            if (cpy > oend - ORC_LASTLITERALS) goto error;
        }
    do_match_copy:
        if (off == 0) {
            /* offset 0 is not rejected; the reference's small-offset path seeds the
             * copy with LZ4_write32(op,0) so the match expands to zero bytes
             * (lz4.c:479, 2301).  Clamp is unnecessary: cpy<=oend was established. */
            memset(dst + op, 0, (size_t)len);
        } else {
            int64_t i;
            for (i = 0; i < len; i++) dst[op + i] = dst[op + i - off];
        }
        op += len;
    }
    return (int)op;
error:
    return (int)(-ip) - 1;                                                      /* lz4.c:2337 */
}

/* LZ4_decompress_fast: lz4.c:1794-1891; every error is -1, success returns bytes read. */
int orc_lz4_decompress_fast(const uint8_t* src, uint8_t* dst, int originalSize)
{
    int64_t ip = 0, op = 0, oend = originalSize;
    if (originalSize < 0) return -1;
    for (;;) {
        unsigned token = src[ip++];
        int64_t ll = token >> 4;
        if (ll == 15) { unsigned b; do { b = src[ip++]; ll += b; } while (b == 255); }
        if (oend - op < ll) return -1;
        memmove(dst + op, src + ip, (size_t)ll); op += ll; ip += ll;
        if (oend - op < ORC_MFLIMIT) { if (op == oend) break; return -1; }
        {
            int64_t ml = token & 15, off = src[ip] | (src[ip + 1] << 8), u;
            ip += 2;
            if (ml == 15) { unsigned b; do { b = src[ip++]; ml += b; } while (b == 255); }
            ml += ORC_MINMATCH;
            if (oend - op < ml) return -1;
            if (off > op) return -1;
            for (u = 0; u < ml; u++) dst[op + u] = dst[op + u - off];
            op += ml;
            if (oend - op < ORC_LASTLITERALS) return -1;
        }
    }
    return (int)ip;
}
