/*
 * oracle/zstd_dec.c — Zstandard frame decoder, CPU restatement of RFC 8878.
 * TEST INFRASTRUCTURE ONLY (see cj_oracle.h).
 *
 * Reference path being restated: src/zstd.rs:23-28 (decompress) and :67-70 (decompress_into)
 * -> libcramjam::zstd::decompress -> zstd::stream::read::Decoder -> libzstd 1.5.7
 * (zstd-sys 2.0.14+zstd.1.5.7, un-vendored, Cargo.lock:1025-1050).  The streaming decoder
 * consumes every concatenated frame and skips skippable frames; so does this one.
 * Pinned in tests/test_oracle_goldens.py / tests/test_oracle_cross.py against tests/golden/plaintext.txt.zst and against
 * frames produced by the system libzstd.so.1 at levels 1..19 over many shapes of input.
 */
#include "cj_oracle.h"
#include <stdlib.h>
#include <string.h>

#define ERR(code) (-(int64_t)(code))
#define ZSTD_MAGIC 0xFD2FB528u
#define BLOCK_MAX (128 * 1024)

typedef struct { uint8_t sym[512]; uint8_t nbits[512]; uint16_t base[512]; int al; } fse_t;
typedef struct { uint8_t sym[2048]; uint8_t nbits[2048]; int max_bits; int valid; } huf_t;

static inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline int hibit(uint32_t v) { return 31 - __builtin_clz(v); }

/* ---- forward bit reader (FSE table descriptions) -------------------------------------------- */
typedef struct { const uint8_t* p; size_t n; size_t bit; } fbits_t;
static uint32_t fb_read(fbits_t* b, int nb) {
    uint32_t v = 0;
    for (int i = 0; i < nb; i++) {
        size_t byte = b->bit >> 3;
        uint32_t x = byte < b->n ? (b->p[byte] >> (b->bit & 7)) & 1u : 0u;
        v |= x << i;
        b->bit++;
    }
    return v;
}

/* ---- backward bit reader ------------------------------------------------------------------------ */
typedef struct { const uint8_t* p; int64_t pos; } bbits_t; /* pos = bits still unread; may go negative */
static int bb_init(bbits_t* b, const uint8_t* p, size_t n) {
    if (n == 0 || p[n - 1] == 0) return -1;
    b->p = p;
    b->pos = (int64_t)(n - 1) * 8 + hibit(p[n - 1]);
    return 0;
}
/* bits [pos-nb, pos) as an integer, most significant = highest position; zero-filled below 0 */
static uint64_t bb_peek(const bbits_t* b, int nb) {
    uint64_t v = 0;
    for (int i = 0; i < nb; i++) {
        int64_t bit = b->pos - 1 - i;
        uint64_t x = bit >= 0 ? (b->p[bit >> 3] >> (bit & 7)) & 1u : 0u;
        v = (v << 1) | x;
    }
    return v;
}
static uint64_t bb_read(bbits_t* b, int nb) {
    uint64_t v = bb_peek(b, nb);
    b->pos -= nb;
    return v;
}

/* ---- FSE ---------------------------------------------------------------------------------------- */
static int fse_build(fse_t* t, const int16_t* freq, int nsym, int al) {
    const int size = 1 << al;
    uint16_t next[256];
    int high = size;
    t->al = al;
    for (int s = 0; s < nsym; s++)
        if (freq[s] == -1) { t->sym[--high] = (uint8_t)s; next[s] = 1; }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < nsym; s++) {
        if (freq[s] <= 0) continue;
        next[s] = (uint16_t)freq[s];
        for (int i = 0; i < freq[s]; i++) {
            t->sym[pos] = (uint8_t)s;
            do { pos = (pos + step) & mask; } while (pos >= high);
        }
    }
    if (pos != 0) return -1;
    for (int i = 0; i < size; i++) {
        uint16_t ns = next[t->sym[i]]++;
        int nb = al - hibit(ns);
        t->nbits[i] = (uint8_t)nb;
        t->base[i] = (uint16_t)(((uint32_t)ns << nb) - (uint32_t)size);
    }
    return 0;
}

/* Reads an FSE table description; returns bytes consumed or -1. */
static int64_t fse_read_desc(fse_t* t, const uint8_t* p, size_t n, int max_al, int max_sym) {
    if (n == 0) return -1;
    fbits_t b = {p, n, 0};
    int al = 5 + (int)fb_read(&b, 4);
    if (al > max_al) return -1;
    int remaining = 1 << al;
    int16_t freq[256];
    int s = 0;
    while (remaining > 0 && s <= max_sym) {
        int nb = hibit((uint32_t)remaining + 1) + 1;
        uint32_t val = fb_read(&b, nb);
        uint32_t lower = (1u << (nb - 1)) - 1;
        uint32_t thresh = (1u << nb) - 1 - ((uint32_t)remaining + 1);
        if ((val & lower) < thresh) { b.bit--; val &= lower; }
        else if (val > lower) val -= thresh;
        int proba = (int)val - 1;
        remaining -= proba < 0 ? -proba : proba;
        freq[s++] = (int16_t)proba;
        if (proba == 0) {
            uint32_t rep = fb_read(&b, 2);
            for (;;) {
                for (uint32_t i = 0; i < rep && s <= max_sym; i++) freq[s++] = 0;
                if (rep == 3) rep = fb_read(&b, 2); else break;
            }
        }
        if ((b.bit + 7) / 8 > n) return -1;
    }
    if (remaining != 0 || s > max_sym + 1) return -1;
    size_t used = (b.bit + 7) / 8;
    if (used > n) return -1;
    if (fse_build(t, freq, s, al) != 0) return -1;
    return (int64_t)used;
}

static void fse_rle(fse_t* t, uint8_t sym) { t->al = 0; t->sym[0] = sym; t->nbits[0] = 0; t->base[0] = 0; }

/* ---- predefined sequence distributions (RFC 8878 3.1.1.3.2.2) ------------------------------------ */
static const int16_t LL_DEFAULT[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int16_t ML_DEFAULT[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const int16_t OF_DEFAULT[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
static const uint32_t LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
static const uint8_t LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint32_t ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
static const uint8_t ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

/* ---- Huffman ------------------------------------------------------------------------------------ */
static int huf_build(huf_t* h, uint8_t* w, int nw) {
    /* w[0..nw) explicit weights; the last symbol's weight is implied */
    uint32_t sum = 0;
    for (int i = 0; i < nw; i++) {
        if (w[i] > 11) return -1;
        if (w[i]) sum += 1u << (w[i] - 1);
    }
    if (sum == 0) return -1;
    int max_bits = hibit(sum) + 1;
    if (max_bits > 11) return -1;
    uint32_t left = (1u << max_bits) - sum;
    if (left & (left - 1)) return -1; /* must be a power of two */
    w[nw] = (uint8_t)(hibit(left) + 1);
    nw++;
    /* at least two symbols of weight 1 or one... libzstd requires an even count of weight-1 */
    int pos = 0;
    for (int wt = 1; wt <= max_bits; wt++) {
        for (int s = 0; s < nw; s++) {
            if (w[s] != wt) continue;
            int cells = 1 << (wt - 1);
            for (int i = 0; i < cells; i++) { h->sym[pos + i] = (uint8_t)s; h->nbits[pos + i] = (uint8_t)(max_bits + 1 - wt); }
            pos += cells;
        }
    }
    if (pos != (1 << max_bits)) return -1;
    h->max_bits = max_bits;
    h->valid = 1;
    return 0;
}

/* Reads a Huffman tree description; returns bytes consumed or -1. */
static int64_t huf_read_desc(huf_t* h, const uint8_t* p, size_t n) {
    if (n == 0) return -1;
    uint8_t w[257];
    int nw = 0;
    int hb = p[0];
    size_t used;
    if (hb >= 128) {
        nw = hb - 127;
        size_t nbytes = (size_t)(nw + 1) / 2;
        if (1 + nbytes > n) return -1;
        for (int i = 0; i < nw; i++) w[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
        used = 1 + nbytes;
    } else {
        if (hb == 0 || (size_t)1 + (size_t)hb > n) return -1;
        fse_t t;
        int64_t c = fse_read_desc(&t, p + 1, (size_t)hb, 6, 255);
        if (c < 0 || c >= hb) return -1;
        bbits_t b;
        if (bb_init(&b, p + 1 + c, (size_t)hb - (size_t)c) != 0) return -1;
        uint32_t s1 = (uint32_t)bb_read(&b, t.al), s2 = (uint32_t)bb_read(&b, t.al);
        if (b.pos < 0) return -1;
        for (;;) {
            if (nw > 253) return -1;
            w[nw++] = t.sym[s1];
            s1 = t.base[s1] + (uint32_t)bb_read(&b, t.nbits[s1]);
            if (b.pos < 0) { w[nw++] = t.sym[s2]; break; }
            if (nw > 253) return -1;
            w[nw++] = t.sym[s2];
            s2 = t.base[s2] + (uint32_t)bb_read(&b, t.nbits[s2]);
            if (b.pos < 0) { w[nw++] = t.sym[s1]; break; }
        }
        used = 1 + (size_t)hb;
    }
    if (huf_build(h, w, nw) != 0) return -1;
    return (int64_t)used;
}

static int huf_decode_stream(const huf_t* h, const uint8_t* p, size_t n, uint8_t* out, size_t count) {
    bbits_t b;
    if (bb_init(&b, p, n) != 0) return -1;
    for (size_t i = 0; i < count; i++) {
        uint32_t idx = (uint32_t)bb_peek(&b, h->max_bits);
        out[i] = h->sym[idx];
        b.pos -= h->nbits[idx];
        if (b.pos < 0) return -1;
    }
    return b.pos == 0 ? 0 : -1;
}

/* ---- frame state -------------------------------------------------------------------------------- */
typedef struct {
    huf_t huf;
    fse_t ll, of, ml;
    int ll_ok, of_ok, ml_ok;
    uint32_t rep[3];
    uint8_t lit[BLOCK_MAX + 32];
} zctx_t;

static int64_t decode_literals(zctx_t* z, const uint8_t* p, size_t n, size_t* lit_len) {
    if (n < 1) return ERR(CJO_E_TRUNCATED);
    int type = p[0] & 3, sf = (p[0] >> 2) & 3;
    size_t hdr, regen, comp = 0;
    int streams = 1;
    if (type < 2) {
        if (sf == 0 || sf == 2) { hdr = 1; regen = p[0] >> 3; }
        else if (sf == 1) { if (n < 2) return ERR(CJO_E_TRUNCATED); hdr = 2; regen = (p[0] >> 4) | ((size_t)p[1] << 4); }
        else { if (n < 3) return ERR(CJO_E_TRUNCATED); hdr = 3; regen = (p[0] >> 4) | ((size_t)p[1] << 4) | ((size_t)p[2] << 12); }
        if (regen > BLOCK_MAX) return ERR(CJO_E_CORRUPT);
        if (type == 0) {
            if (hdr + regen > n) return ERR(CJO_E_TRUNCATED);
            memcpy(z->lit, p + hdr, regen);
            *lit_len = regen;
            return (int64_t)(hdr + regen);
        }
        if (hdr + 1 > n) return ERR(CJO_E_TRUNCATED);
        memset(z->lit, p[hdr], regen);
        *lit_len = regen;
        return (int64_t)(hdr + 1);
    }
    if (sf == 0 || sf == 1) {
        if (n < 3) return ERR(CJO_E_TRUNCATED);
        uint32_t v = p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
        hdr = 3; regen = (v >> 4) & 0x3ff; comp = (v >> 14) & 0x3ff; streams = sf == 0 ? 1 : 4;
    } else if (sf == 2) {
        if (n < 4) return ERR(CJO_E_TRUNCATED);
        uint32_t v = ld32(p);
        hdr = 4; regen = (v >> 4) & 0x3fff; comp = (v >> 18) & 0x3fff; streams = 4;
    } else {
        if (n < 5) return ERR(CJO_E_TRUNCATED);
        uint64_t v = ld32(p) | ((uint64_t)p[4] << 32);
        hdr = 5; regen = (v >> 4) & 0x3ffff; comp = (v >> 22) & 0x3ffff; streams = 4;
    }
    if (regen > BLOCK_MAX) return ERR(CJO_E_CORRUPT);
    if (hdr + comp > n) return ERR(CJO_E_TRUNCATED);
    const uint8_t* q = p + hdr;
    size_t left = comp;
    if (type == 2) {
        int64_t c = huf_read_desc(&z->huf, q, left);
        if (c < 0) return ERR(CJO_E_CORRUPT);
        q += c;
        left -= (size_t)c;
    } else if (!z->huf.valid) {
        return ERR(CJO_E_CORRUPT); /* treeless without a previous table */
    }
    if (streams == 1) {
        if (huf_decode_stream(&z->huf, q, left, z->lit, regen) != 0) return ERR(CJO_E_CORRUPT);
    } else {
        if (left < 6) return ERR(CJO_E_CORRUPT);
        size_t s1 = q[0] | ((size_t)q[1] << 8), s2 = q[2] | ((size_t)q[3] << 8), s3 = q[4] | ((size_t)q[5] << 8);
        if (6 + s1 + s2 + s3 > left) return ERR(CJO_E_CORRUPT);
        size_t s4 = left - 6 - s1 - s2 - s3;
        size_t per = (regen + 3) / 4;
        if (per * 3 > regen) return ERR(CJO_E_CORRUPT);
        const uint8_t* d = q + 6;
        if (huf_decode_stream(&z->huf, d, s1, z->lit, per) != 0) return ERR(CJO_E_CORRUPT);
        if (huf_decode_stream(&z->huf, d + s1, s2, z->lit + per, per) != 0) return ERR(CJO_E_CORRUPT);
        if (huf_decode_stream(&z->huf, d + s1 + s2, s3, z->lit + 2 * per, per) != 0) return ERR(CJO_E_CORRUPT);
        if (huf_decode_stream(&z->huf, d + s1 + s2 + s3, s4, z->lit + 3 * per, regen - 3 * per) != 0) return ERR(CJO_E_CORRUPT);
    }
    *lit_len = regen;
    return (int64_t)(hdr + comp);
}

/* mode: 0 predefined, 1 RLE, 2 FSE description, 3 repeat.  Returns bytes consumed or -1. */
static int64_t seq_table(fse_t* t, int* ok, int mode, const uint8_t* p, size_t n, const int16_t* def, int def_n, int def_al, int max_al, int max_sym) {
    switch (mode) {
    case 0: if (fse_build(t, def, def_n, def_al) != 0) return -1; *ok = 1; return 0;
    case 1: if (n < 1 || p[0] > max_sym) return -1; fse_rle(t, p[0]); *ok = 1; return 1;
    case 2: { int64_t c = fse_read_desc(t, p, n, max_al, max_sym); if (c < 0) return -1; *ok = 1; return c; }
    default: return *ok ? 0 : -1;
    }
}

static int64_t decode_block(zctx_t* z, const uint8_t* p, size_t n, uint8_t* dst_start, uint8_t* dst, size_t cap) {
    size_t lit_len = 0;
    int64_t c = decode_literals(z, p, n, &lit_len);
    if (c < 0) return c;
    p += c;
    n -= (size_t)c;
    if (n < 1) return ERR(CJO_E_TRUNCATED);
    size_t nseq, h;
    if (p[0] < 128) { nseq = p[0]; h = 1; }
    else if (p[0] < 255) { if (n < 2) return ERR(CJO_E_TRUNCATED); nseq = ((size_t)(p[0] - 128) << 8) + p[1]; h = 2; }
    else { if (n < 3) return ERR(CJO_E_TRUNCATED); nseq = (size_t)p[1] + ((size_t)p[2] << 8) + 0x7F00; h = 3; }
    p += h;
    n -= h;
    size_t op = 0, lp = 0;
    if (nseq == 0) {
        if (n != 0) return ERR(CJO_E_CORRUPT);
    } else {
        if (n < 1) return ERR(CJO_E_TRUNCATED);
        int modes = p[0];
        if (modes & 3) return ERR(CJO_E_CORRUPT);
        p++; n--;
        c = seq_table(&z->ll, &z->ll_ok, (modes >> 6) & 3, p, n, LL_DEFAULT, 36, 6, 9, 35);
        if (c < 0) return ERR(CJO_E_CORRUPT);
        p += c; n -= (size_t)c;
        c = seq_table(&z->of, &z->of_ok, (modes >> 4) & 3, p, n, OF_DEFAULT, 29, 5, 8, 31);
        if (c < 0) return ERR(CJO_E_CORRUPT);
        p += c; n -= (size_t)c;
        c = seq_table(&z->ml, &z->ml_ok, (modes >> 2) & 3, p, n, ML_DEFAULT, 53, 6, 9, 52);
        if (c < 0) return ERR(CJO_E_CORRUPT);
        p += c; n -= (size_t)c;
        bbits_t b;
        if (bb_init(&b, p, n) != 0) return ERR(CJO_E_CORRUPT);
        uint32_t sl = (uint32_t)bb_read(&b, z->ll.al), so = (uint32_t)bb_read(&b, z->of.al), sm = (uint32_t)bb_read(&b, z->ml.al);
        if (b.pos < 0) return ERR(CJO_E_CORRUPT);
        size_t hist = (size_t)(dst - dst_start);
        for (size_t i = 0; i < nseq; i++) {
            int oc = z->of.sym[so], lc = z->ll.sym[sl], mc = z->ml.sym[sm];
            if (oc > 31 || lc > 35 || mc > 52) return ERR(CJO_E_CORRUPT);
            uint64_t ov = ((uint64_t)1 << oc) + bb_read(&b, oc);
            size_t mlen = ML_BASE[mc] + (size_t)bb_read(&b, ML_BITS[mc]);
            size_t llen = LL_BASE[lc] + (size_t)bb_read(&b, LL_BITS[lc]);
            if (b.pos < 0) return ERR(CJO_E_CORRUPT);
            size_t off;
            if (ov > 3) {
                off = (size_t)(ov - 3);
                z->rep[2] = z->rep[1]; z->rep[1] = z->rep[0]; z->rep[0] = (uint32_t)off;
            } else {
                uint32_t idx = (uint32_t)ov - 1 + (llen == 0 ? 1 : 0); /* 0..3 */
                if (idx == 0) {
                    off = z->rep[0];
                } else {
                    uint32_t v = idx < 3 ? z->rep[idx] : z->rep[0] - 1;
                    if (v == 0) return ERR(CJO_E_CORRUPT);
                    if (idx > 1) z->rep[2] = z->rep[1];
                    z->rep[1] = z->rep[0];
                    z->rep[0] = v;
                    off = v;
                }
            }
            if (i + 1 < nseq) {
                sl = z->ll.base[sl] + (uint32_t)bb_read(&b, z->ll.nbits[sl]);
                sm = z->ml.base[sm] + (uint32_t)bb_read(&b, z->ml.nbits[sm]);
                so = z->of.base[so] + (uint32_t)bb_read(&b, z->of.nbits[so]);
                if (b.pos < 0) return ERR(CJO_E_CORRUPT);
            }
            if (llen > lit_len - lp) return ERR(CJO_E_CORRUPT);
            if (llen + mlen > cap - op) return (op + llen + mlen > BLOCK_MAX) ? ERR(CJO_E_CORRUPT) : ERR(CJO_E_DST_SMALL);
            memcpy(dst + op, z->lit + lp, llen);
            lp += llen;
            op += llen;
            if (off > op + hist) return ERR(CJO_E_OFFSET);
            if (off >= mlen) memcpy(dst + op, dst + op - off, mlen);
            else for (size_t k = 0; k < mlen; k++) dst[op + k] = dst[op + k - off];
            op += mlen;
        }
        if (b.pos != 0) return ERR(CJO_E_CORRUPT);
    }
    size_t rest = lit_len - lp;
    if (rest > cap - op) return (op + rest > BLOCK_MAX) ? ERR(CJO_E_CORRUPT) : ERR(CJO_E_DST_SMALL);
    memcpy(dst + op, z->lit + lp, rest);
    op += rest;
    if (op > BLOCK_MAX) return ERR(CJO_E_CORRUPT);
    return (int64_t)op;
}

/* Parses a frame header at p.  Returns header size or negative status. */
static int64_t frame_header(const uint8_t* p, size_t n, uint64_t* fcs, int* has_fcs, uint64_t* window, int* checksum) {
    if (n < 5) return ERR(CJO_E_TRUNCATED);
    if (ld32(p) != ZSTD_MAGIC) return ERR(CJO_E_HEADER);
    uint8_t fhd = p[4];
    int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    if (fhd & 0x08) return ERR(CJO_E_HEADER); /* reserved bit */
    *checksum = (fhd >> 2) & 1;
    size_t pos = 5;
    uint64_t win = 0;
    if (!single) {
        if (n < pos + 1) return ERR(CJO_E_TRUNCATED);
        uint8_t wd = p[pos++];
        int wl = 10 + (wd >> 3);
        if (wl > 31) return ERR(CJO_E_UNSUPPORTED);
        win = ((uint64_t)1 << wl) + (((uint64_t)1 << wl) >> 3) * (wd & 7);
    }
    static const int did_sz[4] = {0, 1, 2, 4};
    if (n < pos + (size_t)did_sz[did]) return ERR(CJO_E_TRUNCATED);
    uint32_t dict = 0;
    for (int i = 0; i < did_sz[did]; i++) dict |= (uint32_t)p[pos + i] << (8 * i);
    pos += (size_t)did_sz[did];
    if (dict != 0) return ERR(CJO_E_UNSUPPORTED);
    int fsz = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (n < pos + (size_t)fsz) return ERR(CJO_E_TRUNCATED);
    uint64_t v = 0;
    for (int i = 0; i < fsz; i++) v |= (uint64_t)p[pos + i] << (8 * i);
    if (fsz == 2) v += 256;
    pos += (size_t)fsz;
    *has_fcs = fsz != 0;
    *fcs = v;
    if (single) win = v;
    *window = win;
    return (int64_t)pos;
}

static int64_t zstd_walk(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int len_only) {
    size_t s = 0, d = 0;
    zctx_t* z = len_only ? NULL : (zctx_t*)malloc(sizeof(zctx_t));
    if (!len_only && !z) return ERR(CJO_E_CORRUPT);
    int64_t rc = 0;
    while (s < n) {
        if (n - s < 4) { rc = ERR(CJO_E_TRUNCATED); goto out; }
        uint32_t magic = ld32(src + s);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (n - s < 8) { rc = ERR(CJO_E_TRUNCATED); goto out; }
            size_t sz = ld32(src + s + 4);
            if (sz > n - s - 8) { rc = ERR(CJO_E_TRUNCATED); goto out; }
            s += 8 + sz;
            continue;
        }
        uint64_t fcs, window;
        int has_fcs, checksum;
        int64_t h = frame_header(src + s, n - s, &fcs, &has_fcs, &window, &checksum);
        if (h < 0) { rc = h; goto out; }
        s += (size_t)h;
        size_t bmax = window < BLOCK_MAX ? (size_t)window : BLOCK_MAX;
        size_t frame_start = d;
        if (!len_only) {
            z->huf.valid = 0;
            z->ll_ok = z->of_ok = z->ml_ok = 0;
            z->rep[0] = 1; z->rep[1] = 4; z->rep[2] = 8;
        }
        if (len_only && !has_fcs) { rc = ERR(CJO_E_UNSUPPORTED); goto out; }
        for (;;) {
            if (n - s < 3) { rc = ERR(CJO_E_TRUNCATED); goto out; }
            uint32_t bh = src[s] | ((uint32_t)src[s + 1] << 8) | ((uint32_t)src[s + 2] << 16);
            s += 3;
            int last = bh & 1, type = (bh >> 1) & 3;
            size_t bsz = bh >> 3;
            if (type == 3) { rc = ERR(CJO_E_CORRUPT); goto out; }
            size_t in_sz = type == 1 ? 1 : bsz;
            if (in_sz > n - s) { rc = ERR(CJO_E_TRUNCATED); goto out; }
            if (!len_only) {
                if (type == 2 ? bsz > BLOCK_MAX : bsz > bmax) { rc = ERR(CJO_E_CORRUPT); goto out; }
                if (type == 0) {
                    if (bsz > cap - d) { rc = ERR(CJO_E_DST_SMALL); goto out; }
                    memcpy(dst + d, src + s, bsz);
                    d += bsz;
                } else if (type == 1) {
                    if (bsz > cap - d) { rc = ERR(CJO_E_DST_SMALL); goto out; }
                    memset(dst + d, src[s], bsz);
                    d += bsz;
                } else {
                    size_t room = cap - d;
                    int64_t r = decode_block(z, src + s, bsz, dst + frame_start, dst + d, room);
                    if (r < 0) { rc = r; goto out; }
                    if ((size_t)r > bmax) { rc = ERR(CJO_E_CORRUPT); goto out; }
                    d += (size_t)r;
                }
            }
            s += in_sz;
            if (last) break;
        }
        if (checksum) {
            if (n - s < 4) { rc = ERR(CJO_E_TRUNCATED); goto out; }
            if (!len_only && (uint32_t)cjo_xxh64(dst + frame_start, d - frame_start, 0) != ld32(src + s)) { rc = ERR(CJO_E_CHECKSUM); goto out; }
            s += 4;
        }
        if (len_only) d += (size_t)fcs;
        else if (has_fcs && fcs != (uint64_t)(d - frame_start)) { rc = ERR(CJO_E_LEN_MISMATCH); goto out; }
    }
    rc = (int64_t)d;
out:
    free(z);
    return rc;
}

int64_t cjo_zstd_decompressed_len(const uint8_t* src, size_t n) { return zstd_walk(src, n, NULL, 0, 1); }

int64_t cjo_zstd_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    static uint8_t dummy;
    return zstd_walk(src, n, dst ? dst : &dummy, cap, 0);
}
