/*
 * oracle/xxh_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT (see lz4_block_oracle.c header).
 *
 * Plain-C restatement of xxHash 0.6.5 as vendored by the reference
 * (/root/reference/src/lz4/lib/xxhash.c), one-shot and streaming:
 *   orc_xxh32            xxhash.c:392-416 -> 351-389 (stripe loop 373-378), finalize 290-348
 *   orc_xxh64            xxhash.c:855-879 -> 810-852 (stripe loop 832-837), finalize 701-808
 *   orc_xxh32_reset/update/digest   xxhash.c:437-563
 *   orc_xxh64_reset/update/digest   xxhash.c:898-1016
 * Parity pinned against oracle/_ref and python-xxhash in tests/test_oracle_pin.py.
 */
#include <stdint.h>
#include <string.h>
#include <stddef.h>

#define P32_1 2654435761U
#define P32_2 2246822519U
#define P32_3 3266489917U
#define P32_4  668265263U
#define P32_5  374761393U
#define P64_1 11400714785074694791ULL
#define P64_2 14029467366897019727ULL
#define P64_3  1609587929392839161ULL
#define P64_4  9650029242287828579ULL
#define P64_5  2870177450012600261ULL

static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint32_t le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

static uint32_t round32(uint32_t acc, uint32_t w) { return rotl32(acc + w * P32_2, 13) * P32_1; }  /* :269-275 */
static uint64_t round64(uint64_t acc, uint64_t w) { return rotl64(acc + w * P64_2, 31) * P64_1; }  /* :672-678 */
static uint64_t merge64(uint64_t h, uint64_t v) { return (h ^ round64(0, v)) * P64_1 + P64_4; }     /* :680-686 */

static uint32_t finish32(uint32_t h, const uint8_t* p, size_t rem)
{   /* :290-348 written as loops: 4-byte steps, then bytes, then avalanche :278-286 */
    while (rem >= 4) { h = rotl32(h + le32(p) * P32_3, 17) * P32_4; p += 4; rem -= 4; }
    while (rem)      { h = rotl32(h + (*p++) * P32_5, 11) * P32_1; rem--; }
    h ^= h >> 15; h *= P32_2; h ^= h >> 13; h *= P32_3; h ^= h >> 16;
    return h;
}

static uint64_t finish64(uint64_t h, const uint8_t* p, size_t rem)
{   /* :701-808 as loops; avalanche :688-696 */
    while (rem >= 8) { h = rotl64(h ^ round64(0, le64(p)), 27) * P64_1 + P64_4; p += 8; rem -= 8; }
    if (rem >= 4)    { h = rotl64(h ^ ((uint64_t)le32(p) * P64_1), 23) * P64_2 + P64_3; p += 4; rem -= 4; }
    while (rem)      { h = rotl64(h ^ ((*p++) * P64_5), 11) * P64_1; rem--; }
    h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32;
    return h;
}

uint32_t orc_xxh32(const void* buf, size_t len, uint32_t seed)
{
    const uint8_t* p = (const uint8_t*)buf; const uint8_t* end = p + len; uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + P32_1 + P32_2, v2 = seed + P32_2, v3 = seed, v4 = seed - P32_1;
        do {
            v1 = round32(v1, le32(p)); v2 = round32(v2, le32(p + 4));
            v3 = round32(v3, le32(p + 8)); v4 = round32(v4, le32(p + 12)); p += 16;
        } while (p + 16 <= end);
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else h = seed + P32_5;
    h += (uint32_t)len;
    return finish32(h, p, len & 15);
}

uint64_t orc_xxh64(const void* buf, size_t len, uint64_t seed)
{
    const uint8_t* p = (const uint8_t*)buf; const uint8_t* end = p + len; uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P64_1 + P64_2, v2 = seed + P64_2, v3 = seed, v4 = seed - P64_1;
        do {
            v1 = round64(v1, le64(p)); v2 = round64(v2, le64(p + 8));
            v3 = round64(v3, le64(p + 16)); v4 = round64(v4, le64(p + 24)); p += 32;
        } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = merge64(h, v1); h = merge64(h, v2); h = merge64(h, v3); h = merge64(h, v4);
    } else h = seed + P64_5;
    h += (uint64_t)len;
    return finish64(h, p, len & 31);
}

/* ------------------------------------------------------------ streaming */
typedef struct { uint64_t total; uint32_t v[4]; uint8_t mem[16]; uint32_t memsize; uint32_t seed; } orc_xxh32_state;
typedef struct { uint64_t total; uint64_t v[4]; uint8_t mem[32]; uint32_t memsize; uint64_t seed; } orc_xxh64_state;

size_t orc_xxh32_state_size(void) { return sizeof(orc_xxh32_state); }
size_t orc_xxh64_state_size(void) { return sizeof(orc_xxh64_state); }

void orc_xxh32_reset(orc_xxh32_state* s, uint32_t seed)
{
    memset(s, 0, sizeof *s); s->seed = seed;
    s->v[0] = seed + P32_1 + P32_2; s->v[1] = seed + P32_2; s->v[2] = seed; s->v[3] = seed - P32_1;
}
void orc_xxh32_update(orc_xxh32_state* s, const void* in, size_t len)
{
    const uint8_t* p = (const uint8_t*)in; const uint8_t* end = p + len; int k;
    s->total += len;
    if (s->memsize + len < 16) { memcpy(s->mem + s->memsize, p, len); s->memsize += (uint32_t)len; return; }
    if (s->memsize) {
        memcpy(s->mem + s->memsize, p, 16 - s->memsize);
        for (k = 0; k < 4; k++) s->v[k] = round32(s->v[k], le32(s->mem + 4 * k));
        p += 16 - s->memsize; s->memsize = 0;
    }
    while (p + 16 <= end) { for (k = 0; k < 4; k++) s->v[k] = round32(s->v[k], le32(p + 4 * k)); p += 16; }
    if (p < end) { memcpy(s->mem, p, (size_t)(end - p)); s->memsize = (uint32_t)(end - p); }
}
uint32_t orc_xxh32_digest(const orc_xxh32_state* s)
{
    uint32_t h;
    if (s->total >= 16) h = rotl32(s->v[0], 1) + rotl32(s->v[1], 7) + rotl32(s->v[2], 12) + rotl32(s->v[3], 18);
    else h = s->seed + P32_5;       /* == v[2] + P32_5 in the reference (:548) */
    h += (uint32_t)s->total;
    return finish32(h, s->mem, s->memsize);
}

void orc_xxh64_reset(orc_xxh64_state* s, uint64_t seed)
{
    memset(s, 0, sizeof *s); s->seed = seed;
    s->v[0] = seed + P64_1 + P64_2; s->v[1] = seed + P64_2; s->v[2] = seed; s->v[3] = seed - P64_1;
}
void orc_xxh64_update(orc_xxh64_state* s, const void* in, size_t len)
{
    const uint8_t* p = (const uint8_t*)in; const uint8_t* end = p + len; int k;
    s->total += len;
    if (s->memsize + len < 32) { memcpy(s->mem + s->memsize, p, len); s->memsize += (uint32_t)len; return; }
    if (s->memsize) {
        memcpy(s->mem + s->memsize, p, 32 - s->memsize);
        for (k = 0; k < 4; k++) s->v[k] = round64(s->v[k], le64(s->mem + 8 * k));
        p += 32 - s->memsize; s->memsize = 0;
    }
    while (p + 32 <= end) { for (k = 0; k < 4; k++) s->v[k] = round64(s->v[k], le64(p + 8 * k)); p += 32; }
    if (p < end) { memcpy(s->mem, p, (size_t)(end - p)); s->memsize = (uint32_t)(end - p); }
}
uint64_t orc_xxh64_digest(const orc_xxh64_state* s)
{
    uint64_t h; int k;
    if (s->total >= 32) {
        h = rotl64(s->v[0], 1) + rotl64(s->v[1], 7) + rotl64(s->v[2], 12) + rotl64(s->v[3], 18);
        for (k = 0; k < 4; k++) h = merge64(h, s->v[k]);
    } else h = s->seed + P64_5;
    h += s->total;
    return finish64(h, s->mem, s->memsize);
}
