/*
 * oracle/cpu_bench.c — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * pthread harness that times a per-block CPU function (either the reference's own
 * C entry points from oracle/_ref/liblz4ref.so or the restatements in this
 * directory) over a batch of independent blocks, the way BASELINE.md §3 specifies:
 * static contiguous partition, T threads, best of `passes` after one warm-up,
 * CLOCK_MONOTONIC around the parallel region.  The function pointers are passed in
 * from Python (ctypes), so the same harness serves `cpu_baseline.kind` "reference"
 * and "port".  The call sites mirror src/jni/net_jpountz_lz4_LZ4JNI.c:75,169,216 and
 * net_jpountz_xxhash_XXHashJNI.c:54,164.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int (*fn_compress)(const char*, char*, int, int);
typedef int (*fn_dec_safe)(const char*, char*, int, int);
typedef int (*fn_dec_fast)(const char*, char*, int);
typedef unsigned (*fn_xxh32)(const void*, size_t, unsigned);
typedef unsigned long long (*fn_xxh64)(const void*, size_t, unsigned long long);

enum { OP_COMPRESS = 0, OP_DEC_SAFE = 1, OP_DEC_FAST = 2, OP_XXH32 = 3, OP_XXH64 = 4 };

typedef struct {
    int op; void* fn;
    const uint8_t* src; const uint64_t* src_off; const int32_t* src_len;
    uint8_t* dst; const uint64_t* dst_off; const int32_t* dst_cap;
    int64_t* result; size_t lo, hi;
    pthread_barrier_t* bar; int passes;
} job_t;

static void run_range(job_t* j)
{
    size_t i;
    for (i = j->lo; i < j->hi; i++) {
        const char* s = (const char*)j->src + j->src_off[i];
        char* d = j->dst ? (char*)j->dst + j->dst_off[i] : NULL;
        switch (j->op) {
        case OP_COMPRESS: j->result[i] = ((fn_compress)j->fn)(s, d, j->src_len[i], j->dst_cap[i]); break;
        case OP_DEC_SAFE: j->result[i] = ((fn_dec_safe)j->fn)(s, d, j->src_len[i], j->dst_cap[i]); break;
        case OP_DEC_FAST: j->result[i] = ((fn_dec_fast)j->fn)(s, d, j->dst_cap[i]); break;
        case OP_XXH32:    j->result[i] = ((fn_xxh32)j->fn)(s, (size_t)j->src_len[i], 0); break;
        case OP_XXH64:    j->result[i] = (int64_t)((fn_xxh64)j->fn)(s, (size_t)j->src_len[i], 0); break;
        }
    }
}

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg; int p;
    for (p = 0; p < j->passes + 1; p++) {      /* pass 0 = warm-up */
        pthread_barrier_wait(j->bar);
        run_range(j);
        pthread_barrier_wait(j->bar);
    }
    return NULL;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* Returns the best pass time in seconds (median in *median_s). */
double orc_cpu_bench(int op, void* fn,
                     const uint8_t* src, const uint64_t* src_off, const int32_t* src_len,
                     uint8_t* dst, const uint64_t* dst_off, const int32_t* dst_cap,
                     int64_t* result, size_t n, int threads, int passes, double* median_s)
{
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)threads);
    double* times = (double*)malloc(sizeof(double) * (size_t)passes);
    pthread_barrier_t bar; int t, p; double best = 1e30;
    pthread_barrier_init(&bar, NULL, (unsigned)threads + 1);
    for (t = 0; t < threads; t++) {
        job_t* j = &jobs[t];
        j->op = op; j->fn = fn; j->src = src; j->src_off = src_off; j->src_len = src_len;
        j->dst = dst; j->dst_off = dst_off; j->dst_cap = dst_cap; j->result = result;
        j->lo = n * (size_t)t / (size_t)threads; j->hi = n * (size_t)(t + 1) / (size_t)threads;
        j->bar = &bar; j->passes = passes;
        pthread_create(&th[t], NULL, worker, j);
    }
    for (p = 0; p < passes + 1; p++) {
        double t0, t1;
        pthread_barrier_wait(&bar); t0 = now_s();
        pthread_barrier_wait(&bar); t1 = now_s();
        if (p > 0) { times[p - 1] = t1 - t0; if (t1 - t0 < best) best = t1 - t0; }
    }
    for (t = 0; t < threads; t++) pthread_join(th[t], NULL);
    if (median_s) {   /* insertion sort, passes is tiny */
        int a, b; for (a = 1; a < passes; a++) for (b = a; b > 0 && times[b] < times[b - 1]; b--) { double x = times[b]; times[b] = times[b - 1]; times[b - 1] = x; }
        *median_s = times[passes / 2];
    }
    pthread_barrier_destroy(&bar); free(th); free(jobs); free(times);
    return best;
}
