/*
 * oracle/lz4_java_port_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * lz4-java's pure-Java backends (LZ4Factory.safeInstance(), what BASELINE configs[0]
 * names) restated in plain C — the block compressor of src/build/source_templates/compress.template:16-131 (blocks
 * < 64 KiB + 11: 16-bit table, hash64k) and :133-261 (general: int table pre-filled with the block start, MAX_DISTANCE
 * test), and the two decoders of decompress.template:16-129 with THEIR accept/reject rules, which differ slightly from
 * the C library's (lz4.c) that the JNI backend — and the CUDA path — follow.  Helpers: LZ4Utils.java:43-49 (hash),
 * LZ4SafeUtils.java:36-159 (commonBytes, commonBytesBackward, writeLen, lastLiterals), LZ4Constants.java:24-49.
 *
 * PARITY UNPINNED: there is no JVM in this environment, so these functions cannot be checked against the Java classes
 * themselves.  They are pinned only indirectly — streams they write must decode under the pinned C restatement
 * (lz4_block_oracle.c, itself byte-exact against the reference's C) and they must accept every stream that one writes.
 * Their role is the cross-backend test of LZ4Test.java:305-324 (every compressor against every decompressor): the GPU
 * decoders must read the Java backend's streams, and the Java decoders' rules must accept the GPU compressor's.
 */
#include <stdint.h>
#include <string.h>

enum { MIN_MATCH = 4, COPY_LENGTH = 8, LAST_LITERALS = 5, MF_LIMIT = 12, MIN_LENGTH = 13, ML_BITS = 4, ML_MASK = 15, RUN_MASK = 15,
       HASH_LOG = 12, HASH_LOG_64K = 13, SKIP_STRENGTH = 6, MAX_DISTANCE = 1 << 16, LZ4_64K_LIMIT = (1 << 16) + (MF_LIMIT - 1) };

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }      /* SafeUtils.readInt, little-endian host */
static int jhash(uint32_t v, int log) { return (int)((v * 2654435761u) >> (32 - log)); }  /* i * -1640531535 >>> (32 - log) */

static int common_bytes(const uint8_t* b, int o1, int o2, int limit)                    /* LZ4SafeUtils.java:60-66 */
{ int c = 0; while (o2 < limit && b[o1++] == b[o2++]) ++c; return c; }
static int common_bytes_backward(const uint8_t* b, int o1, int o2, int l1, int l2)      /* LZ4SafeUtils.java:68-74 */
{ int c = 0; while (o1 > l1 && o2 > l2 && b[--o1] == b[--o2]) ++c; return c; }
static int write_len(int len, uint8_t* d, int dOff)                                     /* LZ4SafeUtils.java:151-158 */
{ while (len >= 0xFF) { d[dOff++] = 0xFF; len -= 0xFF; } d[dOff++] = (uint8_t)len; return dOff; }
static int last_literals(const uint8_t* s, int sOff, int runLen, uint8_t* d, int dOff, int destEnd)   /* :133-149 */
{
    if (dOff + runLen + 1 + (runLen + 255 - RUN_MASK) / 255 > destEnd) return -1;
    if (runLen >= RUN_MASK) { d[dOff++] = (uint8_t)(RUN_MASK << ML_BITS); dOff = write_len(runLen - RUN_MASK, d, dOff); }
    else d[dOff++] = (uint8_t)(runLen << ML_BITS);
    memcpy(d + dOff, s + sOff, (size_t)runLen);
    return dOff + runLen;
}

/* LZ4JavaSafeCompressor.compress (compress.template): returns the compressed size, or -1 where Java throws LZ4Exception
 * ("maxDestLen is too small").  srcOff = destOff = 0. */
int orc_java_compress(const uint8_t* src, int srcLen, uint8_t* dest, int maxDestLen)
{
    const int small = srcLen < LZ4_64K_LIMIT;                       /* compress.template:146-148 */
    const int srcEnd = srcLen, srcLimit = srcEnd - LAST_LITERALS, mflimit = srcEnd - MF_LIMIT, destEnd = maxDestLen;
    static __thread int32_t table[1 << HASH_LOG_64K];               /* short[8192] (64k) or int[4096] (general) */
    const int log = small ? HASH_LOG_64K : HASH_LOG;
    int sOff = 0, dOff = 0, anchor = 0;
    if (!small || srcLen >= MIN_LENGTH) {
        memset(table, 0, sizeof table);                              /* new short[] == 0; Arrays.fill(hashTable, anchor = srcOff = 0) */
        ++sOff;
        for (;;) {
            int forwardOff = sOff, ref, step = 1, searchMatchNb = 1 << SKIP_STRENGTH, back;
            for (;;) {                                               /* find a match (:34-50 / :165-181) */
                sOff = forwardOff; forwardOff += step; step = searchMatchNb++ >> SKIP_STRENGTH;
                if (forwardOff > mflimit) goto tail;
                { const int h = jhash(rd32(src + sOff), log); ref = table[h]; table[h] = sOff; }
                back = sOff - ref;
                if ((small || back < MAX_DISTANCE) && rd32(src + ref) == rd32(src + sOff)) break;
            }
            { const int excess = common_bytes_backward(src, ref, sOff, 0, anchor); sOff -= excess; ref -= excess; }
            {
                const int runLen = sOff - anchor;
                int tokenOff = dOff++;
                if (dOff + runLen + (2 + 1 + LAST_LITERALS) + (runLen >> 8) > destEnd) return -1;
                if (runLen >= RUN_MASK) { dest[tokenOff] = (uint8_t)(RUN_MASK << ML_BITS); dOff = write_len(runLen - RUN_MASK, dest, dOff); }
                else dest[tokenOff] = (uint8_t)(runLen << ML_BITS);
                memcpy(dest + dOff, src + anchor, (size_t)runLen);   /* wildArraycopy: same bytes inside the run */
                dOff += runLen;
                for (;;) {
                    int matchLen;
                    dest[dOff] = (uint8_t)(sOff - ref); dest[dOff + 1] = (uint8_t)((sOff - ref) >> 8); dOff += 2;
                    sOff += MIN_MATCH; ref += MIN_MATCH;
                    matchLen = common_bytes(src, ref, sOff, srcLimit);
                    if (dOff + (1 + LAST_LITERALS) + (matchLen >> 8) > destEnd) return -1;
                    sOff += matchLen;
                    if (matchLen >= ML_MASK) { dest[tokenOff] |= ML_MASK; dOff = write_len(matchLen - ML_MASK, dest, dOff); }
                    else dest[tokenOff] |= (uint8_t)matchLen;
                    if (sOff > mflimit) { anchor = sOff; goto tail; }
                    table[jhash(rd32(src + sOff - 2), log)] = sOff - 2;
                    { const int h = jhash(rd32(src + sOff), log); ref = table[h]; table[h] = sOff; }
                    back = sOff - ref;
                    if ((!small && back >= MAX_DISTANCE) || rd32(src + sOff) != rd32(src + ref)) break;
                    tokenOff = dOff++; dest[tokenOff] = 0;
                }
                anchor = sOff++;
            }
        }
    }
tail:
    return last_literals(src, anchor, srcEnd - anchor, dest, dOff, destEnd);
}

/* Shared body of LZ4JavaSafeSafeDecompressor / LZ4JavaSafeFastDecompressor (decompress.template:16-129).
 * safe: srcLen is exact, returns bytes WRITTEN.  fast: `srcLen` is only the readable array length (an index past it is
 * Java's ArrayIndexOutOfBoundsException -> LZ4Exception), returns bytes READ.  -1 wherever Java throws. */
static int java_decompress(const uint8_t* src, int srcLen, uint8_t* dest, int destLen, int safe)
{
    const int srcEnd = srcLen, destEnd = destLen;
    int sOff = 0, dOff = 0;
    if (destLen == 0) {
        if (safe) return (srcLen != 1 || src[0] != 0) ? -1 : 0;                             /* :28-33 */
        return (srcLen < 1 || src[0] != 0) ? -1 : 1;                                        /* :41-46 */
    }
    for (;;) {
        int token, literalLen, literalCopyEnd, matchDec, matchOff, matchLen, matchCopyEnd, i;
        if (sOff >= srcLen) return -1;
        token = src[sOff++];
        literalLen = token >> ML_BITS;
        if (literalLen == RUN_MASK) {
            int len = 0xFF;
            while ((!safe || sOff < srcEnd) && (sOff < srcLen ? 1 : 0) && (len = src[sOff++]) == 0xFF) literalLen += 0xFF;
            if (!safe && len == 0xFF && sOff >= srcLen) return -1;                          /* ran off the array */
            literalLen += len & 0xFF;
        }
        literalCopyEnd = dOff + literalLen;
        if (safe ? (literalCopyEnd > destEnd - COPY_LENGTH || sOff + literalLen > srcEnd - COPY_LENGTH)
                 : (literalCopyEnd > destEnd - COPY_LENGTH)) {
            if (safe) {
                if (literalCopyEnd > destEnd) return -1;
                if (sOff + literalLen != srcEnd) return -1;
            } else {
                if (literalCopyEnd != destEnd) return -1;
                if (sOff + literalLen > srcLen) return -1;
            }
            memcpy(dest + dOff, src + sOff, (size_t)literalLen);
            sOff += literalLen; dOff = literalCopyEnd;
            break;                                                                           /* EOF */
        }
        if (sOff + ((literalLen + 7) & ~7) > srcLen) return -1;                              /* wildArraycopy reads 8 at a time */
        memcpy(dest + dOff, src + sOff, (size_t)literalLen);
        sOff += literalLen; dOff = literalCopyEnd;
        if (sOff + 2 > srcLen) return -1;
        matchDec = src[sOff] | (src[sOff + 1] << 8); sOff += 2;
        matchOff = dOff - matchDec;
        if (matchOff < 0) return -1;                                                         /* :97-99 */
        matchLen = token & ML_MASK;
        if (matchLen == ML_MASK) {
            int len = 0xFF;
            while ((!safe || sOff < srcEnd) && (sOff < srcLen ? 1 : 0) && (len = src[sOff++]) == 0xFF) matchLen += 0xFF;
            if (!safe && len == 0xFF && sOff >= srcLen) return -1;
            matchLen += len & 0xFF;
        }
        matchLen += MIN_MATCH;
        matchCopyEnd = dOff + matchLen;
        if (matchCopyEnd > destEnd - COPY_LENGTH && matchCopyEnd > destEnd) return -1;       /* :113-116 */
        for (i = 0; i < matchLen; i++) dest[dOff + i] = dest[matchOff + i];                  /* byte-serial == both incremental copies */
        dOff = matchCopyEnd;
    }
    return safe ? dOff : sOff;
}
int orc_java_decompress_safe(const uint8_t* src, int srcLen, uint8_t* dest, int maxDestLen) { return java_decompress(src, srcLen, dest, maxDestLen, 1); }
int orc_java_decompress_fast(const uint8_t* src, int srcAvail, uint8_t* dest, int destLen) { return java_decompress(src, srcAvail, dest, destLen, 0); }
