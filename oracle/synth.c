/*
 * oracle/synth.c — the synthetic "Silesia-like" block generator of SURVEY.md 8(d), restated in C for the CPU arms.
 * TEST INFRASTRUCTURE ONLY (see cj_oracle.h): bench.py's reference arm and cpu_baseline leg must not map the product
 * library, so they generate their input here.  Same arithmetic as cramjam_b200/csrc/synth.cuh (the device generator);
 * tests/test_oracle_goldens.py checks that both produce identical bytes.
 *
 * Model: a block is a pure function of (seed, global block index).  8 % of the blocks are incompressible, 8 % highly
 * repetitive, the rest alternate short skewed-literal runs with back-references whose offsets follow the measured
 * Silesia classes (~1 % overlapping, ~23 % within 256 B, ~41 % within 4 KiB, ~35 % beyond).
 */
#include "cj_oracle.h"
#include <pthread.h>

static inline uint64_t sm_next(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline uint8_t sym(uint32_t r) {
    uint32_t x = r & 0xff, y = (r >> 8) & 0xff, z = (r >> 16) & 0xff;
    uint32_t v = (x * y * z) >> 16;
    v = (v * 96) >> 8;
    return (uint8_t)(32 + v);
}

void cjo_synth_block(uint8_t* out, size_t len, uint64_t seed, uint64_t index) {
    uint64_t s = seed ^ (index * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull);
    const uint32_t cls = (uint32_t)(sm_next(&s) % 100u);
    size_t pos = 0;
    if (cls < 8) {
        while (pos < len) {
            uint64_t r = sm_next(&s);
            for (int k = 0; k < 8 && pos < len; k++, r >>= 8) out[pos++] = (uint8_t)r;
        }
        return;
    }
    const int rep = cls < 16;
    while (pos < len) {
        const uint64_t r = sm_next(&s);
        uint32_t ll;
        if (rep) ll = 1 + (uint32_t)(r & 3);
        else if (((r >> 8) & 15) == 0) ll = 1 + (uint32_t)((r >> 12) % 48);
        else if (((r >> 8) & 15) < 10) ll = 0;
        else ll = 1 + (uint32_t)((r >> 12) % 7);
        uint64_t lr = 0;
        for (uint32_t i = 0; i < ll && pos < len; i++) {
            if ((i & 1) == 0) lr = sm_next(&s);
            out[pos++] = sym((uint32_t)(lr >> ((i & 1) * 32)));
        }
        if (pos >= len) break;
        if (pos < 8) continue;
        uint32_t ml;
        if (rep) ml = 16 + (uint32_t)((r >> 20) % 120);
        else if (((r >> 20) & 31) == 0) ml = 8 + (uint32_t)((r >> 26) % 160);
        else ml = 5 + (uint32_t)((r >> 26) % 13);
        const uint32_t oc = (uint32_t)((r >> 40) % 100);
        const uint32_t orr = (uint32_t)(r >> 47);
        uint32_t off;
        if (oc < 1) off = 1 + orr % 8;
        else if (oc < 24) off = 16 + orr % 240;
        else if (oc < 65) off = 256 + orr % 3840;
        else off = 4096 + orr % 61440;
        if (off > pos) off = 1 + (off - 1) % (uint32_t)pos;
        for (uint32_t i = 0; i < ml && pos < len; i++, pos++) out[pos] = out[pos - off];
    }
}

typedef struct { uint8_t* out; size_t n, len; uint64_t seed, first; size_t next; } sjob_t;

static void* sworker(void* a) {
    sjob_t* j = (sjob_t*)a;
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next, 16, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        size_t e = i + 16 < j->n ? i + 16 : j->n;
        for (; i < e; i++) cjo_synth_block(j->out + i * j->len, j->len, j->seed, j->first + i);
    }
    return 0;
}

void cjo_synth_blocks(uint8_t* out, size_t n_blocks, size_t block_len, uint64_t seed, uint64_t first_index, int nthreads) {
    sjob_t j = {out, n_blocks, block_len, seed, first_index, 0};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 1; t < nthreads; t++)
        if (pthread_create(&th[started], 0, sworker, &j) == 0) started++;
    sworker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], 0);
}
