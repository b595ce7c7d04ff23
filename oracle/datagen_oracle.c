/*
 * oracle/datagen_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT (see lz4_block_oracle.c header).
 *
 * Restatement of the reference's synthetic corpus generator, the one `lz4 -b` uses:
 *   orc_datagen(buf,size,matchProba,litProba,seed) == RDG_genBuffer
 *   /root/reference/src/lz4/programs/datagen.c:155-161 (genBuffer), :61-69 (LCG),
 *   :72-90 (literal distribution), :99-148 (genBlock; matchProba < 1.0 path only).
 * Pinned byte-for-byte against oracle/_ref in tests/test_oracle_pin.py.
 */
#include <stdint.h>
#include <stddef.h>

#define LT_SIZE 8192

static uint32_t lcg(uint32_t* s)
{   /* datagen.c:61-69 */
    uint32_t r = *s * 2654435761U;
    r ^= 2246822519U;
    r = (r << 13) | (r >> 19);
    *s = r;
    return r;
}
static uint32_t rand15(uint32_t* s) { return (lcg(s) >> 3) & 32767; }
static uint32_t randlen(uint32_t* s)
{   /* RDG_RANDLENGTH, datagen.c:100 */
    return ((lcg(s) >> 7) & 7) ? (lcg(s) & 15) : (lcg(s) & 511) + 15;
}

void orc_datagen(uint8_t* buf, size_t size, double matchProba, double litProba, uint32_t seed)
{
    static __thread uint8_t lt[LT_SIZE];
    uint32_t const mp32 = (uint32_t)(32768 * matchProba);
    size_t pos = 0;
    if (litProba == 0.0) litProba = matchProba / 4.5;
    {   /* datagen.c:72-90 */
        uint8_t first = litProba <= 0.0 ? 0 : '(', last = litProba <= 0.0 ? 255 : '}';
        uint8_t ch = litProba <= 0.0 ? 0 : '0';
        uint32_t u = 0;
        while (u < LT_SIZE) {
            uint32_t w = (uint32_t)((double)(LT_SIZE - u) * litProba) + 1;
            uint32_t end = u + w < LT_SIZE ? u + w : LT_SIZE;
            while (u < end) lt[u++] = ch;
            ch++;
            if (ch > last) ch = first;
        }
    }
    if (size == 0) return;
    buf[0] = lt[lcg(&seed) & (LT_SIZE - 1)]; pos = 1;
    while (pos < size) {
        if (rand15(&seed) < mp32) {
            size_t len = (size_t)randlen(&seed) + 4, d, m;
            uint32_t off = rand15(&seed) + 1;
            if (off > pos) off = (uint32_t)pos;
            m = pos - off; d = pos + len; if (d > size) d = size;
            while (pos < d) buf[pos++] = buf[m++];
        } else {
            size_t len = randlen(&seed), d = pos + len;
            if (d > size) d = size;
            while (pos < d) buf[pos++] = lt[lcg(&seed) & (LT_SIZE - 1)];
        }
    }
}
