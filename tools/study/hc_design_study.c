/* hc_design_study.c — CPU model for the round-2 redesign of the HC compressor (design tool, not product, not oracle).
 *
 * Question: the current kernel (csrc/lz4hc_compress.cu) searches only the positions its lazy parser visits, one round
 * trip of latency per parse decision, 1 CTA/SM — 1.2 GiB/s.  A throughput-shaped kernel would search EVERY position,
 * 32 consecutive positions per warp against a snapshot of the bucket rings, insert them, and parse afterwards from
 * the per-position (length, distance) arrays.  What does the snapshot cost in ratio, and which cheap intra-chunk
 * candidate source repairs it?  This model answers with exact compressed sizes (LZ4 block size formula) on the
 * reference's own generator, next to LZ4_compress_HC(level 9) from oracle/_ref.
 *
 *   gcc -O2 -o /tmp/hc_study tools/study/hc_design_study.c oracle/datagen_oracle.c -ldl && /tmp/hc_study
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>

void orc_datagen(uint8_t* buf, size_t size, double matchProba, double litProba, uint32_t seed);

typedef struct { int bl, ways, chunk, intra, lazy, backmax; } Cfg;
typedef struct { int ml, dist; } Best;

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hash(uint32_t seq, int bl) { return (seq * 2654435761u) >> (32 - bl); }

static int count(const uint8_t* s, int p, int c, int limit) { int n = 0; while (p + n < limit && s[p + n] == s[c + n]) n++; return n; }

static void consider(const uint8_t* s, int p, int cand, int matchlimit, Best* b)
{
    if (cand < 0 || cand >= p || p - cand > 65535) return;
    if (rd32(s + cand) != rd32(s + p)) return;
    const int ml = count(s, p, cand, matchlimit);
    if (ml >= 4 && (ml > b->ml || (ml == b->ml && p - cand < b->dist))) { b->ml = ml; b->dist = p - cand; }
}

/* per-position best match under the chunked-snapshot search */
static void search_all(const uint8_t* s, int n, const Cfg* c, Best* best)
{
    const int nb = 1 << c->bl, mflimit = n - 12, matchlimit = n - 5;
    int32_t* ring = malloc(sizeof(int32_t) * nb * c->ways);
    uint32_t* head = calloc(nb, sizeof(uint32_t));
    int* last_in_chunk = malloc(sizeof(int) * nb);            /* intra == 1: nearest earlier same-hash position of this chunk */
    for (int i = 0; i < nb * c->ways; i++) ring[i] = -1;
    for (int p = 0; p < n; p++) { best[p].ml = 0; best[p].dist = 0; }
    for (int c0 = 0; c0 <= mflimit; c0 += c->chunk) {
        const int c1 = c0 + c->chunk <= mflimit + 1 ? c0 + c->chunk : mflimit + 1;
        if (c->intra) for (int p = c0; p < c1; p++) last_in_chunk[hash(rd32(s + p), c->bl)] = -1;
        for (int p = c0; p < c1; p++) {
            const uint32_t h = hash(rd32(s + p), c->bl);
            Best b = {0, 0};
            const uint32_t cnt = head[h] < (uint32_t)c->ways ? head[h] : (uint32_t)c->ways;
            for (uint32_t w = 0; w < cnt; w++) consider(s, p, ring[h * c->ways + w], matchlimit, &b);
            if (c->intra == 1) { consider(s, p, last_in_chunk[h], matchlimit, &b); last_in_chunk[h] = p; }
            if (c->intra == 2) for (int q = c0; q < p; q++) if (hash(rd32(s + q), c->bl) == h) consider(s, p, q, matchlimit, &b);  /* all earlier same-hash lanes */
            if (c->intra == 3) { for (int d = 1; d <= 4; d++) consider(s, p, p - d, matchlimit, &b); consider(s, p, last_in_chunk[h], matchlimit, &b); last_in_chunk[h] = p; }
            best[p] = b;
        }
        for (int p = c0; p < c1; p++) {                       /* insert the chunk, in order */
            const uint32_t h = hash(rd32(s + p), c->bl);
            ring[h * c->ways + (head[h]++ % c->ways)] = p;
        }
    }
    free(ring); free(head); free(last_in_chunk);
}

static long seq_size(int lit, int ml) { long z = 1 + lit + 2; if (lit >= 15) z += (lit - 15) / 255 + 1; if (ml - 4 >= 15) z += (ml - 4 - 15) / 255 + 1; return z; }

/* lazy parse over the per-position results (the kernel's rule without its 4-position round boundary) */
static long parse(const uint8_t* s, int n, const Cfg* c, const Best* best, long* nseq)
{
    const int mflimit = n - 12;
    long out = 0; int anchor = 0, ip = 0; *nseq = 0;
    while (ip <= mflimit) {
        if (best[ip].ml < 4) { ip++; continue; }
        int cur = ip;
        for (int k = 0; k < c->lazy && cur + 1 <= mflimit; k++) { if (best[cur + 1].ml > best[cur].ml) cur++; else break; }
        int ms = cur, ml = best[cur].ml; const int dist = best[cur].dist;
        int back = 0;
        while (back < c->backmax && ms - back > anchor && ms - dist - back > 0 && s[ms - back - 1] == s[ms - dist - back - 1]) back++;
        ms -= back; ml += back;
        out += seq_size(ms - anchor, ml); (*nseq)++;
        ip = anchor = ms + ml;
    }
    const int lit = n - anchor;
    out += 1 + lit + (lit >= 15 ? (lit - 15) / 255 + 1 : 0);
    return out;
}

/* Backward dynamic programme over the same per-position results: cost[p] = min(literal: 1 + cost[p+1],
 * match: 3 + ext(l) + cost[p+l]) for l = best[p].ml and (cheaply) a few shorter cuts; then a forward pass emits the
 * chosen sequences and the exact size is counted.  Sequential in p but two or three operations per position —
 * the kind of loop one thread per block does at full multiplicity (cf. algo 4's walk kernel). */
static long parse_dp(const uint8_t* s, int n, const Cfg* c, const Best* best, long* nseq, int cuts)
{
    const int mflimit = n - 12;
    float* cost = malloc(sizeof(float) * (n + 1)); int* take = malloc(sizeof(int) * (n + 1));
    for (int p = n; p > mflimit; p--) { cost[p] = (float)(n - p); take[p] = 0; }
    for (int p = mflimit; p >= 0; p--) {
        float bc = 1.0f + cost[p + 1] + 1.0f / 255; int bt = 0;             /* literal (amortised run-length byte) */
        const int L = best[p].ml;
        if (L >= 4) {
            for (int k = 0; k < cuts; k++) {
                int l = k == 0 ? L : (k == 1 ? L - 1 : (k == 2 ? L - 2 : (k == 3 ? L - 3 : 4 + (L - 4) * (k - 3) / (cuts - 3))));
                if (l < 4) break;
                if (p + l > n - 5) l = n - 5 - p;
                if (l < 4) break;
                const float mc = 3.0f + (l - 4 >= 15 ? (float)((l - 19) / 255 + 1) : 0.0f) + cost[p + l];
                if (mc < bc) { bc = mc; bt = l; }
            }
        }
        cost[p] = bc; take[p] = bt;
    }
    long out = 0; int anchor = 0, ip = 0; *nseq = 0;
    while (ip <= mflimit) {
        if (!take[ip]) { ip++; continue; }
        int ms = ip, ml = take[ip]; const int dist = best[ip].dist;
        int back = 0;
        while (back < c->backmax && ms - back > anchor && ms - dist - back > 0 && s[ms - back - 1] == s[ms - dist - back - 1]) back++;
        ms -= back; ml += back;
        out += seq_size(ms - anchor, ml); (*nseq)++;
        ip = anchor = ms + ml;
    }
    const int lit = n - anchor;
    out += 1 + lit + (lit >= 15 ? (lit - 15) / 255 + 1 : 0);
    free(cost); free(take);
    return out;
}

int main(int argc, char** argv)
{
    const int bs = 262144, nblk = 8;
    const double probas[] = {0.5, 0.8, 0.2};
    void* ref = dlopen("oracle/_ref/liblz4ref.so", RTLD_NOW);
    int (*hc)(const char*, char*, int, int, int) = ref ? (int (*)(const char*, char*, int, int, int))dlsym(ref, "LZ4_compress_HC") : NULL;
    const Cfg cfgs[] = {
        {11, 32, 1, 0, 3, 8},      /* sequential rings = what the current kernel approximates */
        {11, 32, 32, 0, 3, 8},     /* 32-position snapshot, no intra-chunk candidates */
        {11, 32, 32, 1, 3, 8},     /* + nearest earlier same-hash position of the chunk (__match_any_sync) */
        {11, 32, 32, 2, 3, 8},     /* + all earlier same-hash positions of the chunk */
        {11, 32, 32, 3, 3, 8},     /* + nearest same-hash + the four previous positions (runs / short periods) */
        {11, 32, 128, 1, 3, 8},    /* 4 warps x 32 positions per snapshot */
        {11, 32, 128, 3, 3, 8},
        {11, 16, 32, 3, 3, 8},     /* half the ring memory */
        {12, 16, 32, 3, 3, 8},     /* same memory as 11x32, more buckets */
        {12, 32, 32, 3, 3, 8},     /* 256 KiB of rings (global memory / L2 instead of shared) */
        {13, 32, 32, 3, 3, 8},
        {11, 32, 32, 3, 6, 8},     /* deeper lazy */
        {11, 32, 32, 3, 3, 64},    /* longer catch-up */
    };
    uint8_t* buf = malloc((size_t)bs * nblk); char* tmp = malloc(bs + bs / 255 + 64); Best* best = malloc(sizeof(Best) * bs);
    for (int a = 1; a < argc; a++) {                          /* optional: real files, cut into 256 KiB blocks */
        FILE* f = fopen(argv[a], "rb"); if (!f) { perror(argv[a]); continue; }
        const size_t got = fread(buf, 1, (size_t)bs * nblk, f); fclose(f);
        const int nb = (int)((got + bs - 1) / bs);
        long refsz = 0;
        for (int b = 0; b < nb; b++) { const int len = (int)((size_t)(b + 1) * bs <= got ? bs : got - (size_t)b * bs); if (hc) refsz += hc((const char*)buf + (size_t)b * bs, tmp, len, bs + bs / 255 + 64, 9); }
        printf("%s, %zu bytes in %d blocks: LZ4_compress_HC(9) ratio %.4f\n", argv[a], got, nb, refsz ? (double)got / refsz : 0.0);
        for (unsigned ci = 0; ci < sizeof(cfgs) / sizeof(cfgs[0]); ci++) {
            long tot = 0, ns = 0, q;
            for (int b = 0; b < nb; b++) { const int len = (int)((size_t)(b + 1) * bs <= got ? bs : got - (size_t)b * bs); if (len < 13) { tot += len + 1; continue; }
                search_all(buf + (size_t)b * bs, len, &cfgs[ci], best); tot += parse(buf + (size_t)b * bs, len, &cfgs[ci], best, &q); ns += q; }
            printf("  buckets 2^%d ways %2d chunk %3d intra %d lazy %d back %2d : ratio %.4f  (%.1f bytes/sequence)", cfgs[ci].bl, cfgs[ci].ways,
                   cfgs[ci].chunk, cfgs[ci].intra, cfgs[ci].lazy, cfgs[ci].backmax, (double)got / tot, (double)got / (ns ? ns : 1));
            if (cfgs[ci].chunk == 32 && cfgs[ci].intra == 3 && cfgs[ci].lazy == 3 && cfgs[ci].backmax == 8) {
                for (int cuts = 1; cuts <= 8; cuts += 3) {
                    long t2 = 0;
                    for (int b = 0; b < nb; b++) { const int len = (int)((size_t)(b + 1) * bs <= got ? bs : got - (size_t)b * bs); if (len < 13) { t2 += len + 1; continue; }
                        search_all(buf + (size_t)b * bs, len, &cfgs[ci], best); t2 += parse_dp(buf + (size_t)b * bs, len, &cfgs[ci], best, &q, cuts); }
                    printf("  | DP(%d cuts) %.4f", cuts, (double)got / t2);
                }
            }
            printf("\n");
        }
    }
    if (argc > 1) return 0;
    for (unsigned pi = 0; pi < sizeof(probas) / sizeof(probas[0]); pi++) {
        orc_datagen(buf, (size_t)bs * nblk, probas[pi], 0.0, 4);
        long refsz = 0;
        if (hc) for (int b = 0; b < nblk; b++) refsz += hc((const char*)buf + (size_t)b * bs, tmp, bs, bs + bs / 255 + 64, 9);
        printf("RDG P=%.2f, %d x 256 KiB: LZ4_compress_HC(9) ratio %.4f\n", probas[pi], nblk, refsz ? (double)bs * nblk / refsz : 0.0);
        for (unsigned ci = 0; ci < sizeof(cfgs) / sizeof(cfgs[0]); ci++) {
            long tot = 0, ns = 0, q;
            for (int b = 0; b < nblk; b++) { search_all(buf + (size_t)b * bs, bs, &cfgs[ci], best); tot += parse(buf + (size_t)b * bs, bs, &cfgs[ci], best, &q); ns += q; }
            printf("  buckets 2^%d ways %2d chunk %3d intra %d lazy %d back %2d : ratio %.4f  (%.1f bytes/sequence)", cfgs[ci].bl, cfgs[ci].ways,
                   cfgs[ci].chunk, cfgs[ci].intra, cfgs[ci].lazy, cfgs[ci].backmax, (double)bs * nblk / tot, (double)bs * nblk / ns);
            if (cfgs[ci].chunk == 32 && cfgs[ci].intra == 3 && cfgs[ci].lazy == 3 && cfgs[ci].backmax == 8) {
                for (int cuts = 1; cuts <= 8; cuts += 3) {
                    long t2 = 0;
                    for (int b = 0; b < nblk; b++) { search_all(buf + (size_t)b * bs, bs, &cfgs[ci], best); t2 += parse_dp(buf + (size_t)b * bs, bs, &cfgs[ci], best, &q, cuts); }
                    printf("  | DP(%d cuts) %.4f", cuts, (double)bs * nblk / t2);
                }
            }
            printf("\n");
        }
    }
    return 0;
}
