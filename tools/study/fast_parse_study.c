// fast_parse_study.c — CPU model of the fast compressor's lookup phase (128-position sub-rounds: probe all, then insert
// all) and of the greedy parse over its hits.  Development aid: statistics per chunk that decide how the GPU parser is
// shaped (hits, run starts, selected sequences, literal lengths), and the compressed size of parse variants.
//   gcc -O2 -o /tmp/fps tools/study/fast_parse_study.c && /tmp/fps file [block_bytes]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

int main(int argc, char** argv)
{
    FILE* f = fopen(argv[1], "rb"); if (!f) return 1;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* data = malloc(sz + 64); if (fread(data, 1, sz, f) != (size_t)sz) return 1; fclose(f);
    int bs = argc > 2 ? atoi(argv[2]) : 65536;
    int HL = argc > 3 ? atoi(argv[3]) : 13;
    long tot_in = 0, tot_out = 0, tot_hits = 0, tot_runs = 0, tot_seq = 0, tot_lit_gt16 = 0, tot_lit_gt32 = 0, tot_long = 0, tot_pos = 0;
    long lit_hist[8] = {0}; long ml_hist[8] = {0}; long tot_out_noback = 0; long hit_ml_hist[8] = {0};
    long runs_per512_hist[9] = {0};
    uint16_t* dist = malloc(70000 * 2);
    for (long off = 0; off + bs <= sz; off += bs) {
        const uint8_t* s = data + off; int n = bs;
        static uint16_t table[1 << 14]; memset(table, 0, sizeof table);
        int mflimit = n - 12, matchlimit = n - 5;
        memset(dist, 0, 70000 * 2);
        for (int c0 = 0; c0 <= mflimit; c0 += 128) {
            int cand[128];
            int e = c0 + 128 <= mflimit + 1 ? 128 : mflimit + 1 - c0;
            for (int i = 0; i < e; i++) cand[i] = table[(rd32(s + c0 + i) * 2654435761u) >> (32 - HL)];
            for (int i = 0; i < e; i++) table[(rd32(s + c0 + i) * 2654435761u) >> (32 - HL)] = (uint16_t)(c0 + i);
            for (int i = 0; i < e; i++) { int p = c0 + i; if (cand[i] < p && rd32(s + cand[i]) == rd32(s + p)) dist[p] = (uint16_t)(p - cand[i]); }
        }
        // stats: hits, run starts (per 512 chunk: position 0 of a chunk always starts a run)
        for (int p = 0; p <= mflimit; p++) {
            tot_pos++;
            if (dist[p]) { tot_hits++; if ((p & 511) == 0 || dist[p] != dist[p - 1]) tot_runs++; }
        }
        for (int c0 = 0; c0 <= mflimit; c0 += 512) {
            int r = 0;
            for (int p = c0; p < c0 + 512 && p <= mflimit; p++) if (dist[p] && ((p & 511) == 0 || dist[p] != dist[p - 1])) r++;
            runs_per512_hist[r >= 64 ? 8 : r / 8]++;
        }
        // greedy parse over the hits with full measurement (+ catch-up <= 4)
        int ip = 0, op = 0;
        for (int p = 0; p <= mflimit; ) {
            if (!dist[p]) { p++; continue; }
            int d = dist[p], ms = p, mc = p - d, ml = 4;
            while (ms + ml < matchlimit && s[ms + ml] == s[mc + ml]) ml++;
            int back = 0;
            while (back < 4 && ms - back > ip && mc - back > 0 && s[ms - back - 1] == s[mc - back - 1]) back++;
            ms -= back; ml += back;
            int lit = ms - ip;
            op += 1 + (lit >= 15 ? (lit - 15) / 255 + 1 : 0) + lit + 2 + (ml - 4 >= 15 ? (ml - 4 - 15) / 255 + 1 : 0);
            tot_seq++; if (lit > 16) tot_lit_gt16++; if (lit > 32) tot_lit_gt32++; if (ml > 32) tot_long++;
            lit_hist[lit == 0 ? 0 : lit <= 4 ? 1 : lit <= 8 ? 2 : lit <= 16 ? 3 : lit <= 32 ? 4 : lit <= 64 ? 5 : lit <= 128 ? 6 : 7]++;
            ml_hist[ml < 8 ? 0 : ml < 12 ? 1 : ml < 16 ? 2 : ml < 20 ? 3 : ml < 36 ? 4 : ml < 68 ? 5 : ml < 132 ? 6 : 7]++;
            ip = ms + ml; p = ip;
        }
        for (int p = 0; p <= mflimit; p++) if (dist[p]) { int d = dist[p], ml = 4; while (p + ml < matchlimit && s[p + ml] == s[p - d + ml]) ml++;
            hit_ml_hist[ml < 8 ? 0 : ml < 12 ? 1 : ml < 16 ? 2 : ml < 20 ? 3 : ml < 36 ? 4 : ml < 68 ? 5 : ml < 132 ? 6 : 7]++; }
        {   // same parse without catch-up
            int ip2 = 0, op2 = 0;
            for (int p = 0; p <= mflimit; ) {
                if (!dist[p]) { p++; continue; }
                int d = dist[p], ms = p, mc = p - d, ml = 4;
                while (ms + ml < matchlimit && s[ms + ml] == s[mc + ml]) ml++;
                int lit = ms - ip2;
                op2 += 1 + (lit >= 15 ? (lit - 15) / 255 + 1 : 0) + lit + 2 + (ml - 4 >= 15 ? (ml - 4 - 15) / 255 + 1 : 0);
                ip2 = ms + ml; p = ip2;
            }
            int lit = n - ip2;
            op2 += 1 + (lit >= 15 ? (lit - 15) / 255 + 1 : 0) + lit;
            tot_out_noback += op2;
        }
        int lit = n - ip;
        op += 1 + (lit >= 15 ? (lit - 15) / 255 + 1 : 0) + lit;
        tot_in += n; tot_out += op;
    }
    double nb = tot_in / 512.0;
    printf("%s: ratio %.4f  per 512 B: hits %.1f  run starts %.1f  sequences %.1f  (lit>16: %.2f, lit>32: %.2f, ml>32: %.2f)\n", argv[1],
           (double)tot_in / tot_out, tot_hits / nb, tot_runs / nb, tot_seq / nb, tot_lit_gt16 / nb, tot_lit_gt32 / nb, tot_long / nb);
    printf("  literal-length histogram (0,1-4,5-8,9-16,17-32,33-64,65-128,>128):");
    for (int i = 0; i < 8; i++) printf(" %.3f", (double)lit_hist[i] / tot_seq);
    printf("\n  selected match-length histogram (<8,<12,<16,<20,<36,<68,<132,more):");
    for (int i = 0; i < 8; i++) printf(" %.3f", (double)ml_hist[i] / tot_seq);
    printf("\n  all-hit match-length histogram:");
    for (int i = 0; i < 8; i++) printf(" %.3f", (double)hit_ml_hist[i] / tot_hits);
    printf("\n  ratio without catch-up: %.4f", (double)tot_in / tot_out_noback);
    printf("\n  run starts per 512-chunk histogram (0-7,8-15,...,>=64):");
    long t = 0; for (int i = 0; i < 9; i++) t += runs_per512_hist[i];
    for (int i = 0; i < 9; i++) printf(" %.3f", (double)runs_per512_hist[i] / t);
    printf("\n");
    return 0;
}
