/*
 * oracle/checksums.c — CRC-32C, XXH32, XXH64 restated from their public definitions.
 * TEST INFRASTRUCTURE ONLY (see cj_oracle.h).
 *
 * Used by: snappy framing (masked CRC-32C per chunk; reference path src/snappy.rs:22-42 ->
 * snap::read::FrameDecoder / FrameEncoder), LZ4F header + content checksum (XXH32;
 * src/lz4.rs:27-65), zstd content checksum (low 32 bits of XXH64; src/zstd.rs:23-28).
 * Pinned in tests/test_oracle_checksums.py against crc32c("123456789") = 0xE3069283, the
 * checksum words inside the three golden fixtures, and the `xxhash` Python module.
 */
#include "cj_oracle.h"
#include <string.h>

static uint32_t crc_tab[8][256];
static int crc_ready = 0;

static void crc_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
        crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++)
            crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xff];
    __atomic_store_n(&crc_ready, 1, __ATOMIC_RELEASE);
}

uint32_t cjo_crc32c(const void* p, size_t n) {
    if (!__atomic_load_n(&crc_ready, __ATOMIC_ACQUIRE)) crc_init();
    const uint8_t* s = (const uint8_t*)p;
    uint32_t c = 0xFFFFFFFFu;
    while (n >= 8) { /* slice-by-8 */
        uint32_t lo, hi;
        memcpy(&lo, s, 4);
        memcpy(&hi, s + 4, 4);
        lo ^= c;
        c = crc_tab[7][lo & 0xff] ^ crc_tab[6][(lo >> 8) & 0xff] ^ crc_tab[5][(lo >> 16) & 0xff] ^
            crc_tab[4][lo >> 24] ^ crc_tab[3][hi & 0xff] ^ crc_tab[2][(hi >> 8) & 0xff] ^
            crc_tab[1][(hi >> 16) & 0xff] ^ crc_tab[0][hi >> 24];
        s += 8;
        n -= 8;
    }
    while (n--) c = (c >> 8) ^ crc_tab[0][(c ^ *s++) & 0xff];
    return c ^ 0xFFFFFFFFu;
}

uint32_t cjo_crc32c_masked(const void* p, size_t n) {
    uint32_t c = cjo_crc32c(p, n);
    return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

/* ---- XXH32 ---------------------------------------------------------------------------- */
#define P32_1 2654435761u
#define P32_2 2246822519u
#define P32_3 3266489917u
#define P32_4 668265263u
#define P32_5 374761393u
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

uint32_t cjo_xxh32(const void* ptr, size_t n, uint32_t seed) {
    const uint8_t* p = (const uint8_t*)ptr;
    const uint8_t* end = p + n;
    uint32_t h;
    if (n >= 16) {
        uint32_t v1 = seed + P32_1 + P32_2, v2 = seed + P32_2, v3 = seed, v4 = seed - P32_1;
        const uint8_t* lim = end - 16;
        do {
            v1 = rotl32(v1 + rd32(p) * P32_2, 13) * P32_1;
            v2 = rotl32(v2 + rd32(p + 4) * P32_2, 13) * P32_1;
            v3 = rotl32(v3 + rd32(p + 8) * P32_2, 13) * P32_1;
            v4 = rotl32(v4 + rd32(p + 12) * P32_2, 13) * P32_1;
            p += 16;
        } while (p <= lim);
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = seed + P32_5;
    }
    h += (uint32_t)n;
    while (p + 4 <= end) { h = rotl32(h + rd32(p) * P32_3, 17) * P32_4; p += 4; }
    while (p < end) { h = rotl32(h + (*p++) * P32_5, 11) * P32_1; }
    h ^= h >> 15; h *= P32_2; h ^= h >> 13; h *= P32_3; h ^= h >> 16;
    return h;
}

/* ---- XXH64 ---------------------------------------------------------------------------- */
#define P64_1 11400714785074694791ull
#define P64_2 14029467366897019727ull
#define P64_3 1609587929392839161ull
#define P64_4 9650029242287828579ull
#define P64_5 2870177450012600261ull
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t xx64_round(uint64_t acc, uint64_t in) { return rotl64(acc + in * P64_2, 31) * P64_1; }
static inline uint64_t xx64_merge(uint64_t h, uint64_t v) { return (h ^ xx64_round(0, v)) * P64_1 + P64_4; }

uint64_t cjo_xxh64(const void* ptr, size_t n, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)ptr;
    const uint8_t* end = p + n;
    uint64_t h;
    if (n >= 32) {
        uint64_t v1 = seed + P64_1 + P64_2, v2 = seed + P64_2, v3 = seed, v4 = seed - P64_1;
        const uint8_t* lim = end - 32;
        do {
            v1 = xx64_round(v1, rd64(p));
            v2 = xx64_round(v2, rd64(p + 8));
            v3 = xx64_round(v3, rd64(p + 16));
            v4 = xx64_round(v4, rd64(p + 24));
            p += 32;
        } while (p <= lim);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xx64_merge(h, v1); h = xx64_merge(h, v2); h = xx64_merge(h, v3); h = xx64_merge(h, v4);
    } else {
        h = seed + P64_5;
    }
    h += (uint64_t)n;
    while (p + 8 <= end) { h = rotl64(h ^ xx64_round(0, rd64(p)), 27) * P64_1 + P64_4; p += 8; }
    if (p + 4 <= end) { h = rotl64(h ^ ((uint64_t)rd32(p) * P64_1), 23) * P64_2 + P64_3; p += 4; }
    while (p < end) { h = rotl64(h ^ ((*p++) * P64_5), 11) * P64_1; }
    h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32;
    return h;
}
