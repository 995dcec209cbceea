/*
 * oracle/snappy.c — Snappy raw block + framed stream, CPU restatement.
 * TEST INFRASTRUCTURE ONLY (see cj_oracle.h).
 *
 * Reference path being restated (the arithmetic itself lives in the un-vendored crate
 * snap 1.1.1, Cargo.lock:744-746):
 *   raw   : src/snappy.rs:52-60   decompress_raw      -> snap::raw::Decoder::decompress_vec
 *           src/snappy.rs:70-78   compress_raw        -> snap::raw::Encoder::compress_vec
 *           src/snappy.rs:93-108  *_raw_into          -> snap::raw::{Encoder,Decoder} slice fns
 *           src/snappy.rs:112-122 compress_raw_max_len / decompress_raw_len
 *   framed: src/snappy.rs:22-42,81-90 compress/decompress(_into) -> snap::read::Frame{En,De}coder
 * Format: google/snappy format_description.txt and framing_format.txt.
 */
#include "cj_oracle.h"
#include <string.h>

#define ERR(code) (-(int64_t)(code))

static inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ld64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

size_t cjo_snappy_max_compressed_len(size_t n) { return 32 + n + n / 6; }

/* uvarint32 preamble.  Returns header length (1..5) or 0 on error; value in *val. */
static int read_preamble(const uint8_t* src, size_t n, uint64_t* val) {
    uint64_t v = 0;
    for (int i = 0; i < 5 && (size_t)i < n; i++) {
        uint8_t b = src[i];
        v |= (uint64_t)(b & 0x7f) << (7 * i);
        if (!(b & 0x80)) { *val = v; return i + 1; }
    }
    return 0;
}

int64_t cjo_snappy_raw_decompressed_len(const uint8_t* src, size_t n) {
    if (n == 0) return 0; /* snap::raw::decompress_len(b"") == Ok(0) */
    uint64_t v;
    if (!read_preamble(src, n, &v)) return ERR(CJO_E_HEADER);
    if (v > 0xFFFFFFFFull) return ERR(CJO_E_TOO_BIG);
    return (int64_t)v;
}

int64_t cjo_snappy_raw_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    if (n == 0) return ERR(CJO_E_EMPTY);
    uint64_t ulen;
    int h = read_preamble(src, n, &ulen);
    if (!h) return ERR(CJO_E_HEADER);
    if (ulen > 0xFFFFFFFFull) return ERR(CJO_E_TOO_BIG);
    if (ulen > cap) return ERR(CJO_E_DST_SMALL);
    size_t s = (size_t)h, d = 0, dn = (size_t)ulen;
    while (s < n) {
        uint8_t tag = src[s++];
        size_t len, off;
        if ((tag & 3) == 0) { /* literal */
            len = (size_t)(tag >> 2) + 1;
            if (len > 60) {
                size_t nb = len - 60;
                if (nb > n - s) return ERR(CJO_E_TRUNCATED);
                uint32_t v = 0;
                for (size_t i = 0; i < nb; i++) v |= (uint32_t)src[s + i] << (8 * i);
                s += nb;
                len = (size_t)v + 1;
            }
            if (len > n - s) return ERR(CJO_E_TRUNCATED);
            if (len > dn - d) return ERR(CJO_E_LEN_MISMATCH);
            memcpy(dst + d, src + s, len);
            s += len;
            d += len;
            continue;
        }
        switch (tag & 3) {
        case 1:
            if (n - s < 1) return ERR(CJO_E_TRUNCATED);
            len = 4 + ((tag >> 2) & 7);
            off = ((size_t)(tag >> 5) << 8) | src[s];
            s += 1;
            break;
        case 2:
            if (n - s < 2) return ERR(CJO_E_TRUNCATED);
            len = 1 + (tag >> 2);
            off = (size_t)src[s] | ((size_t)src[s + 1] << 8);
            s += 2;
            break;
        default:
            if (n - s < 4) return ERR(CJO_E_TRUNCATED);
            len = 1 + (tag >> 2);
            off = ld32(src + s);
            s += 4;
            break;
        }
        if (off == 0 || off > d) return ERR(CJO_E_OFFSET);
        if (len > dn - d) return ERR(CJO_E_LEN_MISMATCH);
        if (off >= len) {
            memcpy(dst + d, dst + d - off, len);
        } else { /* overlapping: byte-serial semantics = pattern replication */
            for (size_t i = 0; i < len; i++) dst[d + i] = dst[d + i - off];
        }
        d += len;
    }
    if (d != dn) return ERR(CJO_E_LEN_MISMATCH);
    return (int64_t)d;
}

/* ---- encoder (Google CompressFragment scheme as carried by snap::raw::Encoder) --------- */
static uint8_t* emit_literal(uint8_t* op, const uint8_t* lit, size_t len) {
    size_t n = len - 1;
    if (n < 60) {
        *op++ = (uint8_t)(n << 2);
    } else if (n < 256) {
        *op++ = 60 << 2; *op++ = (uint8_t)n;
    } else if (n < 65536) {
        *op++ = 61 << 2; *op++ = (uint8_t)n; *op++ = (uint8_t)(n >> 8);
    } else if (n < (1u << 24)) {
        *op++ = 62 << 2; *op++ = (uint8_t)n; *op++ = (uint8_t)(n >> 8); *op++ = (uint8_t)(n >> 16);
    } else {
        *op++ = 63 << 2; *op++ = (uint8_t)n; *op++ = (uint8_t)(n >> 8); *op++ = (uint8_t)(n >> 16); *op++ = (uint8_t)(n >> 24);
    }
    memcpy(op, lit, len);
    return op + len;
}

static uint8_t* emit_copy_upto64(uint8_t* op, size_t off, size_t len) {
    if (len < 12 && off < 2048) {
        *op++ = (uint8_t)(1 | ((len - 4) << 2) | ((off >> 8) << 5));
        *op++ = (uint8_t)off;
    } else {
        *op++ = (uint8_t)(2 | ((len - 1) << 2));
        *op++ = (uint8_t)off;
        *op++ = (uint8_t)(off >> 8);
    }
    return op;
}

static uint8_t* emit_copy(uint8_t* op, size_t off, size_t len) {
    while (len >= 68) { op = emit_copy_upto64(op, off, 64); len -= 64; }
    if (len > 64) { op = emit_copy_upto64(op, off, 60); len -= 60; }
    return emit_copy_upto64(op, off, len);
}

static uint8_t* compress_fragment(const uint8_t* in, size_t n, uint8_t* op, uint16_t* table, int table_bits) {
    const int shift = 32 - table_bits;
    const uint8_t* ip = in;
    const uint8_t* end = in + n;
    const uint8_t* next_emit = ip;
#define HASH(p) ((ld32(p) * 0x1E35A7BDu) >> shift)
    if (n >= 15 + 2) {
        const uint8_t* ip_limit = end - 15;
        ip++;
        uint32_t next_hash = HASH(ip);
        for (;;) {
            uint32_t skip = 32;
            const uint8_t* next_ip = ip;
            const uint8_t* cand;
            do {
                ip = next_ip;
                uint32_t h = next_hash;
                uint32_t step = skip >> 5;
                skip += step;
                next_ip = ip + step;
                if (next_ip > ip_limit) goto remainder;
                next_hash = HASH(next_ip);
                cand = in + table[h];
                table[h] = (uint16_t)(ip - in);
            } while (ld32(ip) != ld32(cand));
            op = emit_literal(op, next_emit, (size_t)(ip - next_emit));
            do {
                const uint8_t* base = ip;
                const uint8_t* a = cand + 4;
                const uint8_t* b = ip + 4;
                while (b + 8 <= end && ld64(a) == ld64(b)) { a += 8; b += 8; }
                while (b < end && *a == *b) { a++; b++; }
                ip = b;
                op = emit_copy(op, (size_t)(base - cand), (size_t)(ip - base));
                next_emit = ip;
                if (ip >= ip_limit) goto remainder;
                table[HASH(ip - 1)] = (uint16_t)(ip - 1 - in);
                uint32_t h = HASH(ip);
                cand = in + table[h];
                table[h] = (uint16_t)(ip - in);
            } while (ld32(ip) == ld32(cand));
            ip++;
            next_hash = HASH(ip);
        }
    }
remainder:
    if (next_emit < end) op = emit_literal(op, next_emit, (size_t)(end - next_emit));
#undef HASH
    return op;
}

int64_t cjo_snappy_raw_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    if (n > 0xFFFFFFFFull) return ERR(CJO_E_TOO_BIG);
    if (cap < cjo_snappy_max_compressed_len(n)) return ERR(CJO_E_DST_SMALL);
    uint8_t* op = dst;
    uint64_t v = n;
    while (v >= 0x80) { *op++ = (uint8_t)(v | 0x80); v >>= 7; }
    *op++ = (uint8_t)v;
    uint16_t table[1 << 14];
    size_t pos = 0;
    while (pos < n) {
        size_t blk = n - pos < 65536 ? n - pos : 65536;
        if (blk < 17) {
            op = emit_literal(op, src + pos, blk);
        } else {
            int bits = 8;
            while (bits < 14 && ((size_t)1 << bits) < blk) bits++;
            memset(table, 0, sizeof(uint16_t) << bits);
            op = compress_fragment(src + pos, blk, op, table, bits);
        }
        pos += blk;
    }
    return (int64_t)(op - dst);
}

/* ---- framing format ---------------------------------------------------------------------- */
static const uint8_t STREAM_ID[10] = {0xff, 0x06, 0x00, 0x00, 's', 'N', 'a', 'P', 'p', 'Y'};

size_t cjo_snappy_frame_max_compressed_len(size_t n) {
    size_t chunks = (n + 65535) / 65536;
    return 10 + chunks * (8 + cjo_snappy_max_compressed_len(65536)) + 16;
}

int64_t cjo_snappy_frame_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    if (cap < 10) return ERR(CJO_E_DST_SMALL);
    uint8_t* op = dst;
    memcpy(op, STREAM_ID, 10);
    op += 10;
    uint8_t tmp[32 + 65536 + 65536 / 6 + 8];
    size_t pos = 0;
    while (pos < n) {
        size_t blk = n - pos < 65536 ? n - pos : 65536;
        uint32_t crc = cjo_crc32c_masked(src + pos, blk);
        int64_t c = cjo_snappy_raw_compress(src + pos, blk, tmp, sizeof tmp);
        if (c < 0) return c;
        /* snap: keep the compressed form only when it saves at least 12.5% */
        int use_comp = (size_t)c < blk - blk / 8;
        size_t body = use_comp ? (size_t)c : blk;
        if ((size_t)(dst + cap - op) < 8 + body) return ERR(CJO_E_DST_SMALL);
        size_t clen = body + 4;
        op[0] = use_comp ? 0x00 : 0x01;
        op[1] = (uint8_t)clen; op[2] = (uint8_t)(clen >> 8); op[3] = (uint8_t)(clen >> 16);
        op[4] = (uint8_t)crc; op[5] = (uint8_t)(crc >> 8); op[6] = (uint8_t)(crc >> 16); op[7] = (uint8_t)(crc >> 24);
        memcpy(op + 8, use_comp ? tmp : src + pos, body);
        op += 8 + body;
        pos += blk;
    }
    return (int64_t)(op - dst);
}

/* Shared walker: when dst == NULL only sums the decompressed lengths. */
static int64_t frame_walk(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    size_t s = 0, d = 0;
    int seen_id = 0;
    while (s < n) {
        if (n - s < 4) return ERR(CJO_E_TRUNCATED);
        uint8_t type = src[s];
        size_t len = (size_t)src[s + 1] | ((size_t)src[s + 2] << 8) | ((size_t)src[s + 3] << 16);
        s += 4;
        if (len > n - s) return ERR(CJO_E_TRUNCATED);
        if (!seen_id && type != 0xff) return ERR(CJO_E_HEADER); /* stream must open with the identifier */
        if (type == 0xff) {
            if (len != 6 || memcmp(src + s, "sNaPpY", 6) != 0) return ERR(CJO_E_HEADER);
            seen_id = 1;
        } else if (type == 0x00 || type == 0x01) {
            if (len < 4) return ERR(CJO_E_CORRUPT);
            uint32_t want = ld32(src + s);
            const uint8_t* body = src + s + 4;
            size_t blen = len - 4;
            size_t ulen;
            if (type == 0x00) {
                int64_t u = cjo_snappy_raw_decompressed_len(body, blen);
                if (u < 0) return u;
                if (blen == 0) return ERR(CJO_E_EMPTY);
                ulen = (size_t)u;
            } else {
                ulen = blen;
            }
            if (ulen > 65536) return ERR(CJO_E_CORRUPT);
            if (dst) {
                if (ulen > cap - d) return ERR(CJO_E_DST_SMALL);
                if (type == 0x00) {
                    int64_t r = cjo_snappy_raw_decompress(body, blen, dst + d, ulen);
                    if (r < 0) return r;
                } else {
                    memcpy(dst + d, body, blen);
                }
                if (cjo_crc32c_masked(dst + d, ulen) != want) return ERR(CJO_E_CHECKSUM);
            }
            d += ulen;
        } else if (type >= 0x80) {
            /* 0x80-0xfd reserved skippable, 0xfe padding: skip */
        } else {
            return ERR(CJO_E_CORRUPT); /* 0x02-0x7f reserved unskippable */
        }
        s += len;
    }
    return (int64_t)d;
}

int64_t cjo_snappy_frame_decompressed_len(const uint8_t* src, size_t n) { return frame_walk(src, n, NULL, 0); }

int64_t cjo_snappy_frame_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    static uint8_t dummy;
    return frame_walk(src, n, dst ? dst : &dummy, cap);
}
