/*
 * oracle/container_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT (see lz4_block_oracle.c header).
 *
 * CPU restatements of lz4-java's own containers (pure Java in the reference, not runnable here — no JDK —
 * so parity is "pinned by format": both directions are checked against these restatements and, for the
 * frame format, against the reference's C lz4frame via oracle/_ref):
 *   orc_lz4block_compress / orc_lz4block_decompress   LZ4BlockOutputStream.java:39-56,203-266;
 *                                                      LZ4BlockInputStream.java:191-264
 *   orc_with_length_compress / _decompress             LZ4CompressorWithLength.java:45-50;
 *                                                      LZ4DecompressorWithLength.java:52-54,125-131
 */
#include <stdint.h>
#include <string.h>
#include <stddef.h>

int orc_lz4_compress_default(const uint8_t*, uint8_t*, int, int);
int orc_lz4_decompress_fast(const uint8_t*, uint8_t*, int);
int orc_lz4_compress_bound(int);
uint32_t orc_xxh32(const void*, size_t, uint32_t);

static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static uint32_t get32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static const char MAGIC[8] = { 'L', 'Z', '4', 'B', 'l', 'o', 'c', 'k' };
#define HDR 21
#define SEED 0x9747b28cU

size_t orc_lz4block_bound(size_t n, int bs) { size_t nb = (n + bs - 1) / bs; return (nb + 1) * HDR + n + nb * 16 + n / 255; }

int64_t orc_lz4block_compress(const uint8_t* src, size_t n, uint8_t* dst, int bs, uint8_t* scratch)
{
    int level = 0; size_t pos = 0, o = 0;
    while ((1 << level) < bs) level++;
    level = level > 10 ? level - 10 : 0;
    while (pos < n) {
        int len = (int)(n - pos < (size_t)bs ? n - pos : (size_t)bs);
        int c = orc_lz4_compress_default(src + pos, scratch, len, orc_lz4_compress_bound(len));
        int raw = c >= len;
        memcpy(dst + o, MAGIC, 8);
        dst[o + 8] = (uint8_t)((raw ? 0x10 : 0x20) | level);
        put32(dst + o + 9, (uint32_t)(raw ? len : c)); put32(dst + o + 13, (uint32_t)len);
        put32(dst + o + 17, orc_xxh32(src + pos, (size_t)len, SEED) & 0x0FFFFFFFU);
        memcpy(dst + o + HDR, raw ? src + pos : scratch, (size_t)(raw ? len : c));
        o += HDR + (size_t)(raw ? len : c); pos += (size_t)len;
    }
    memcpy(dst + o, MAGIC, 8); dst[o + 8] = (uint8_t)(0x10 | level);
    put32(dst + o + 9, 0); put32(dst + o + 13, 0); put32(dst + o + 17, 0);
    return (int64_t)(o + HDR);
}

/* stop: LZ4BlockInputStream's stopOnEmptyBlock (LZ4BlockInputStream.java:60-72,191-233) */
int64_t orc_lz4block_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int stop)
{
    size_t ip = 0, op = 0;
    for (;;) {
        int token, method, level; int32_t clen, olen; uint32_t check;
        if (n - ip < HDR) { if (stop) return -1; break; }
        if (memcmp(src + ip, MAGIC, 8) != 0) return -2;
        token = src[ip + 8]; method = token & 0xF0; level = 10 + (token & 0x0F);
        if (method != 0x10 && method != 0x20) return -2;
        clen = (int32_t)get32(src + ip + 9); olen = (int32_t)get32(src + ip + 13); check = get32(src + ip + 17);
        if (olen > (1 << level) || olen < 0 || clen < 0 || (olen == 0 && clen != 0) || (olen != 0 && clen == 0) ||
            (method == 0x10 && olen != clen)) return -2;
        ip += HDR;
        if (olen == 0) { if (check != 0) return -2; if (stop) break; continue; }
        if (n - ip < (size_t)clen) return -1;
        if (cap - op < (size_t)olen) return -9;
        if (method == 0x10) memcpy(dst + op, src + ip, (size_t)olen);
        else if (orc_lz4_decompress_fast(src + ip, dst + op, olen) != clen) return -2;
        if ((orc_xxh32(dst + op, (size_t)olen, SEED) & 0x0FFFFFFFU) != check) return -2;
        ip += (size_t)clen; op += (size_t)olen;
    }
    return (int64_t)op;
}

int orc_with_length_compress(const uint8_t* src, uint8_t* dst, int n, int cap)
{
    int r;
    if (cap < 4) return 0;
    r = orc_lz4_compress_default(src, dst + 4, n, cap - 4);
    if (r <= 0) return r;
    put32(dst, (uint32_t)n);
    return r + 4;
}
int orc_with_length_decompress(const uint8_t* src, uint8_t* dst, int cap)
{
    int n = (int)get32(src), r;
    if (n < 0 || n > cap) return -1;
    r = orc_lz4_decompress_fast(src + 4, dst, n);
    return r < 0 ? r : r + 4;
}
