/*
 * oracle/batch.c — pthread batch driver over independent units, used ONLY by the CPU-baseline
 * timing legs of bench.py (cpu_baseline / --impl reference) and by tests.
 * TEST INFRASTRUCTURE ONLY (see cj_oracle.h).
 *
 * Mirrors how the reference is driven from many Python threads: every codec call releases the
 * GIL (src/lib.rs:225-288), and units are independent, so a thread pool over units is the
 * reference's own best case on the host cores.
 */
#define _GNU_SOURCE
#include "cj_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <time.h>

typedef struct {
    int codec, dir;
    size_t n;
    const uint8_t* src_base; const uint64_t* src_off; const uint64_t* src_len;
    uint8_t* dst_base; const uint64_t* dst_off; const uint64_t* dst_cap;
    int64_t* out_len;
    size_t next; /* atomic work counter */
} job_t;

static int64_t run_one(int codec, int dir, const uint8_t* s, size_t n, uint8_t* d, size_t cap) {
    if (dir == 0) {
        switch (codec) {
        case CJO_SNAPPY_RAW: return cjo_snappy_raw_decompress(s, n, d, cap);
        case CJO_SNAPPY_FRAMED: return cjo_snappy_frame_decompress(s, n, d, cap);
        case CJO_LZ4_BLOCK: return cjo_lz4_block_decompress(s, n, d, cap);
        case CJO_LZ4_FRAME: return cjo_lz4f_decompress(s, n, d, cap);
        case CJO_ZSTD: return cjo_zstd_decompress(s, n, d, cap);
        }
    } else {
        switch (codec) {
        case CJO_SNAPPY_RAW: return cjo_snappy_raw_compress(s, n, d, cap);
        case CJO_SNAPPY_FRAMED: return cjo_snappy_frame_compress(s, n, d, cap);
        case CJO_LZ4_BLOCK: return cjo_lz4_block_compress(s, n, d, cap, 1);
        case CJO_LZ4_FRAME: return cjo_lz4f_compress(s, n, d, cap, 1 | 2);
        }
    }
    return -(int64_t)CJO_E_UNSUPPORTED;
}

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next, 8, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        size_t e = i + 8 < j->n ? i + 8 : j->n;
        for (; i < e; i++)
            j->out_len[i] = run_one(j->codec, j->dir, j->src_base + j->src_off[i], (size_t)j->src_len[i],
                                    j->dst_base + j->dst_off[i], (size_t)j->dst_cap[i]);
    }
    return NULL;
}

int cjo_batch(int codec, int dir, size_t n, const uint8_t* src_base, const uint64_t* src_off, const uint64_t* src_len,
              uint8_t* dst_base, const uint64_t* dst_off, const uint64_t* dst_cap, int64_t* out_len, int nthreads,
              double* seconds) {
    job_t j = {codec, dir, n, src_base, src_off, src_len, dst_base, dst_off, dst_cap, out_len, 0};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) return -1;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int started = 0;
    for (int t = 0; t < nthreads - 1; t++)
        if (pthread_create(&th[started], NULL, worker, &j) == 0) started++;
    worker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(th);
    return 0;
}

const char* cjo_version(void) { return "cj_oracle 0.1 (snappy raw/framed, lz4 block/frame, zstd decode)"; }

/* ---- dense packing of units (set-up of the benchmark arenas, untimed): dst[do[i]..] = src[so[i] .. so[i]+len[i]) ---- */
#include <string.h>
typedef struct { const uint8_t* src; const uint64_t* so; const uint64_t* len; uint8_t* dst; const uint64_t* dof; size_t n, next; } pjob_t;
static void* pworker(void* arg) {
    pjob_t* j = (pjob_t*)arg;
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next, 64, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        size_t e = i + 64 < j->n ? i + 64 : j->n;
        for (; i < e; i++) memcpy(j->dst + j->dof[i], j->src + j->so[i], (size_t)j->len[i]);
    }
    return NULL;
}
void cjo_pack_units(const uint8_t* src, const uint64_t* so, const uint64_t* len, uint8_t* dst, const uint64_t* dof, size_t n, int nthreads) {
    pjob_t j = {src, so, len, dst, dof, n, 0};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    pthread_t th[64];
    int started = 0;
    for (int t = 1; t < nthreads; t++)
        if (pthread_create(&th[started], NULL, pworker, &j) == 0) started++;
    pworker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}
