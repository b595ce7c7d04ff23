/*
 * oracle/frame_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT (see lz4_block_oracle.c header).
 *
 * CPU restatement of the LZ4 Frame container as lz4-java reads/writes it:
 *   orc_frame_compress    LZ4FrameOutputStream.java:178-251 (header 178-191, writeBlock 199-235,
 *                         end mark 243-251); format: src/lz4/doc/lz4_Frame_format.md
 *   orc_frame_decompress  LZ4FrameInputStream.java:132-321 (nextFrameInfo 132-160, skippable
 *                         162-173, readHeader 180-224, readBlock 258-321)
 * Blocks are independent (FLG bit 5 is mandatory, LZ4FrameOutputStream.java:361-363).
 * Pinned against the reference's LZ4F_compressFrame / LZ4F_decompress (oracle/_ref) in
 * tests/test_oracle_pin.py: frames written here decode there and vice versa.
 */
#include <stdint.h>
#include <string.h>
#include <stddef.h>

int orc_lz4_compress_default(const uint8_t*, uint8_t*, int, int);
int orc_lz4_decompress_safe(const uint8_t*, uint8_t*, int, int);
int orc_lz4_compress_bound(int);
uint32_t orc_xxh32(const void*, size_t, uint32_t);
typedef struct { uint64_t total; uint32_t v[4]; uint8_t mem[16]; uint32_t memsize; uint32_t seed; } orc_xxh32_state;
void orc_xxh32_reset(orc_xxh32_state*, uint32_t);
void orc_xxh32_update(orc_xxh32_state*, const void*, size_t);
uint32_t orc_xxh32_digest(const orc_xxh32_state*);

#define FR_MAGIC      0x184D2204U
#define FR_SKIP_BASE  0x184D2A50U
#define FR_RAW_BIT    0x80000000U

static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static uint32_t get32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

static size_t block_max(int code) { return (size_t)1 << (8 + 2 * code); }   /* 4..7 -> 64K,256K,1M,4M */

size_t orc_frame_bound(size_t n, int bsCode)
{
    size_t bs = block_max(bsCode), nb = (n + bs - 1) / bs;
    return 4 + 2 + 8 + 1 + nb * 8 + n + 4 + 4 + 64;
}

/* flags: bit0 content checksum, bit1 block checksum, bit2 content size.  Returns bytes written, 0 on error. */
size_t orc_frame_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int bsCode, int flags, uint8_t* scratch)
{
    size_t bs, pos = 0, o = 0, hdr;
    uint8_t flg;
    if (bsCode < 4 || bsCode > 7) return 0;
    bs = block_max(bsCode);
    if (cap < orc_frame_bound(n, bsCode)) return 0;
    put32(dst + o, FR_MAGIC); o += 4; hdr = o;
    flg = (uint8_t)((1 << 6) | (1 << 5) | ((flags & 2) ? 1 << 4 : 0) | ((flags & 4) ? 1 << 3 : 0) | ((flags & 1) ? 1 << 2 : 0));
    dst[o++] = flg; dst[o++] = (uint8_t)(bsCode << 4);
    if (flags & 4) { put32(dst + o, (uint32_t)n); put32(dst + o + 4, (uint32_t)((uint64_t)n >> 32)); o += 8; }
    dst[o] = (uint8_t)((orc_xxh32(dst + hdr, o - hdr, 0) >> 8) & 0xFF); o++;
    while (pos < n) {
        size_t len = n - pos < bs ? n - pos : bs;
        int c = orc_lz4_compress_default(src + pos, scratch, (int)len, orc_lz4_compress_bound((int)len));
        const uint8_t* payload; uint32_t word;
        if (c <= 0 || (size_t)c >= len) { payload = src + pos; word = (uint32_t)len | FR_RAW_BIT; c = (int)len; }
        else { payload = scratch; word = (uint32_t)c; }
        put32(dst + o, word); o += 4;
        memcpy(dst + o, payload, (size_t)c); o += (size_t)c;
        if (flags & 2) { put32(dst + o, orc_xxh32(payload, (size_t)c, 0)); o += 4; }
        pos += len;
    }
    put32(dst + o, 0); o += 4;
    if (flags & 1) { put32(dst + o, orc_xxh32(src, n, 0)); o += 4; }
    return o;
}

/*
 * Decodes every frame in [src, src+n) (concatenated and skippable frames included) into dst.
 * Returns total decoded bytes, or a negative code:
 *  -1 premature end, -2 bad magic, -3 descriptor checksum, -4 block too large,
 *  -5 block checksum, -6 block decode error, -7 content checksum, -8 content size, -9 dst too small,
 *  -10 unsupported descriptor (version/reserved bits, dependent blocks).
 */
int64_t orc_frame_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
    size_t ip = 0, op = 0;
    int seen = 0;
    while (ip < n) {
        uint32_t magic; uint8_t flg, bd; size_t hdr, bs, frame_start; uint64_t want = 0;
        orc_xxh32_state st;
        if (n - ip < 4) return -1;
        magic = get32(src + ip); ip += 4;
        if ((magic >> 4) == (FR_SKIP_BASE >> 4)) {
            uint32_t sz; if (n - ip < 4) return -1;
            sz = get32(src + ip); ip += 4;
            if (n - ip < sz) return -1;
            ip += sz; seen = 1; continue;
        }
        if (magic != FR_MAGIC) return -2;
        hdr = ip;
        if (n - ip < 3) return -1;
        flg = src[ip++]; bd = src[ip++];
        if ((flg >> 6) != 1 || (flg & 2) || !(flg & (1 << 5)) || (flg & 1)) return -10;
        if ((bd & 0x8F) || (bd >> 4) < 4) return -10;
        bs = block_max(bd >> 4);
        if (flg & (1 << 3)) { if (n - ip < 9) return -1; want = (uint64_t)get32(src + ip) | ((uint64_t)get32(src + ip + 4) << 32); ip += 8; }
        if (n - ip < 1) return -1;
        if (((orc_xxh32(src + hdr, ip - hdr, 0) >> 8) & 0xFF) != src[ip]) return -3;
        ip++;
        orc_xxh32_reset(&st, 0);
        frame_start = op;
        for (;;) {
            uint32_t word, sz; int raw;
            if (n - ip < 4) return -1;
            word = get32(src + ip); ip += 4;
            raw = (word & FR_RAW_BIT) != 0; sz = word & ~FR_RAW_BIT;
            if (sz == 0) break;
            if (sz > bs) return -4;
            if (n - ip < sz) return -1;
            if (flg & (1 << 4)) {
                if (n - ip < (size_t)sz + 4) return -1;
                if (get32(src + ip + sz) != orc_xxh32(src + ip, sz, 0)) return -5;
            }
            if (raw) {
                if (cap - op < sz) return -9;
                memcpy(dst + op, src + ip, sz);
                if (flg & 4) orc_xxh32_update(&st, dst + op, sz);
                op += sz;
            } else {
                size_t room = cap - op < bs ? cap - op : bs;
                int d = orc_lz4_decompress_safe(src + ip, dst + op, (int)sz, (int)room);
                if (d < 0) return room < bs ? -9 : -6;
                if (flg & 4) orc_xxh32_update(&st, dst + op, (size_t)d);
                op += (size_t)d;
            }
            ip += sz + ((flg & (1 << 4)) ? 4 : 0);
        }
        if (flg & 4) { if (n - ip < 4) return -1; if (get32(src + ip) != orc_xxh32_digest(&st)) return -7; ip += 4; }
        if ((flg & (1 << 3)) && want != (uint64_t)(op - frame_start)) return -8;
        seen = 1;
    }
    return seen ? (int64_t)op : -1;
}
