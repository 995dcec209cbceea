// lz_decode6.cuh — generation-6 batch decode of Snappy raw / LZ4 blocks: the part shared by the device kernels
// (lz_decode6.cu) and the host-side emulation the CPU tests run (tests/emu/g6_emu.cpp).
//
// Reference entry points: snap::raw::Decoder::decompress behind cramjam.snappy.decompress_raw / decompress_raw_into
// (src/snappy.rs:52-60,102-108) and LZ4_decompress_safe behind lz4::block::decompress_into (src/lz4.rs:78-95,140-173).
//
// Design.  The thread-per-block kernel (generation 4) keeps 65 536 output windows live at once, so its
// back-reference fetches come from DRAM (4.75x the algorithmic traffic).  Here a block's 64 KiB output window lives in
// shared memory and a whole CTA decodes the block:
//
//   * a WALK kernel (one thread per block, the only inherently serial part of an LZ77 byte format) validates the
//     stream and leaves a CHECKPOINT for every 128 output bytes: where in the compressed stream, and in which state,
//     a decoder that starts at that output position has to continue;
//   * the EXEC kernel (one persistent CTA per SM) brings a block's compressed bytes and checkpoints into shared memory
//     with bulk copies (cp.async.bulk + mbarrier), and every lane decodes one 128-byte RANGE of the output from its
//     checkpoint, straight into the window.  Literals come from the staged input, back-references from the window;
//   * a back-reference may point at bytes another lane has not produced yet.  A bitmap (one bit per output byte, each
//     32-bit word written by exactly one lane) says which bytes are final; a piece whose source is not final is skipped
//     and the lane goes on with the rest of its range (tools/sim_ranges.py: in-order lanes serialise to ~5 useful
//     lanes per block, out-of-order lanes keep ~50 % of 256 busy), coming back to the skipped pieces on its next pass;
//   * the finished window leaves with one bulk store.  Back-references never leave the SM.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define G6_HD __host__ __device__ __forceinline__
#define G6_HDM __host__ __device__ __forceinline__
#else
#define G6_HD static inline
#define G6_HDM inline
#endif

namespace cj {
namespace g6 {

constexpr uint32_t R = 128;            // output bytes per range (one checkpoint each)
constexpr uint32_t MAXU = 65536;       // largest block the window holds
constexpr uint32_t NR = MAXU / R;      // ranges per full block
constexpr uint32_t CK_BYTES = 8;       // bytes per checkpoint
// what the decoder that starts at a checkpoint does first
enum : uint32_t { ST_TOKEN = 0, ST_LIT = 1, ST_COPY = 2, ST_MATCH = 3 };
// checkpoint word 0: input position [0,17) | LZ4 match-length nibble [24,28) | state [30,32)
// checkpoint word 1: remaining bytes of the element that covers the boundary [0,16) | its offset (copies) [16,32)
//   ST_TOKEN  an element (Snappy) / a sequence token (LZ4) starts at the input position
//   ST_LIT    `rem` literal bytes start at the input position (LZ4: then the match part with the nibble follows)
//   ST_COPY   `rem` bytes of a copy with the given offset remain; the next element / token is at the input position
//   ST_MATCH  LZ4 only: the match part (offset, length extension) of a sequence starts at the input position
G6_HD uint32_t ck_w0(uint32_t ip, uint32_t st, uint32_t nib) { return ip | (nib << 24) | (st << 30); }
G6_HD uint32_t ck_w1(uint32_t rem, uint32_t off) { return rem | (off << 16); }

G6_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {   // (hi:lo >> sh) & 0xffffffff, sh in [0,32)
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
G6_HD uint32_t low_mask(uint32_t n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }   // n in [0,32]

// Unaligned little-endian reads from a 8-byte aligned byte array with at least 8 readable bytes behind every position
// that is asked for (shared memory on the device: two aligned loads and a funnel shift).
G6_HD uint32_t rd32u(const uint8_t* base, uint32_t pos) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (pos & ~3u));
    return funnel_r(w[0], w[1], (pos & 3u) * 8u);
}
G6_HD uint64_t rd64u(const uint8_t* base, uint32_t pos) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (pos & ~3u));
    const uint32_t s = (pos & 3u) * 8u;
    const uint32_t a = w[0], b = w[1], c = w[2];
    return (uint64_t)funnel_r(a, b, s) | ((uint64_t)funnel_r(b, c, s) << 32);
}

// Writes the low n (1..8) bytes of v at byte position o of a window whose 8-byte words around o belong to the caller
// alone: read-modify-write of the one or two aligned words (pieces are written out of order, so no byte outside
// [o, o+n) may change).
G6_HD void put_bytes(uint8_t* win, uint32_t o, uint64_t v, uint32_t n) {
    uint64_t* p = reinterpret_cast<uint64_t*>(win + (o & ~7u));
    const uint32_t sh = (o & 7u) * 8u;
    const uint64_t m = n >= 8 ? ~0ull : ((1ull << (8u * n)) - 1ull);
    v &= m;
    p[0] = (p[0] & ~(m << sh)) | (v << sh);
    if ((o & 7u) + n > 8u) p[1] = (p[1] & ~(m >> (64u - sh))) | (v >> (64u - sh));
}

// What a lane needs of the block it works on (all pointers: shared memory on the device).
struct Block {
    uint8_t* win;           // output window, byte p of the block at win[p]; 8-byte aligned, 16 bytes of slack behind ulen
    const uint8_t* in;      // compressed input, byte i at in[i]; 8-byte aligned, 16 bytes of slack behind n
    uint32_t* bits;         // bit p of word p / 32: output byte p is final.  One slack word behind the last
    const uint32_t* ck;     // checkpoints, two words per range
    uint32_t n, ulen;
};

struct Range {
    uint32_t ip;            // input position of the next element / token to decode
    uint32_t o, hi, lo;     // output cursor, end and start of the range
    uint32_t rem;           // bytes of the current element still to produce inside the range
    uint32_t src;           // literal: input position of the next literal byte; copy: the offset
    uint32_t st;            // ST_LIT / ST_COPY: kind of the current element;  with rem == 0: ST_TOKEN / ST_MATCH = what to decode next
    uint32_t nib;           // LZ4: match-length nibble of the sequence whose literals are being produced
    uint32_t e_ip, e_o, e_st;   // where the current element started (input position, output position, state | nibble << 8)
    uint32_t p_ip, p_o, p_st;   // the first element of this pass that had to be skipped: the next pass resumes there
    bool pending;           // a piece was skipped in this pass
    bool from_ck;           // ... and it was the element the checkpoint starts in: the next pass resumes at the checkpoint
};

enum { STEP_BUSY = 0, STEP_DONE = 1 };

template <int CODEC>
G6_HD void range_start(Range& r, const Block& b, uint32_t k) {
    const uint32_t w0 = b.ck[2 * k], w1 = b.ck[2 * k + 1];
    r.lo = k * R;
    r.hi = r.lo + R < b.ulen ? r.lo + R : b.ulen;
    r.o = r.lo;
    r.ip = w0 & 0x1FFFFu;
    r.nib = (w0 >> 24) & 15u;
    r.st = w0 >> 30;
    r.rem = w1 & 0xFFFFu;
    r.src = w1 >> 16;
    r.pending = false;
    r.from_ck = false;
    r.e_ip = 0xFFFFFFFFu;            // the element the checkpoint starts in (if any) is resumed through the checkpoint
    r.e_o = r.lo;
    r.e_st = 0;
    r.p_ip = r.p_o = r.p_st = 0;
    if (r.st == ST_LIT) {            // the literal bytes sit at ip, the next element behind them
        r.src = r.ip;
        r.ip += r.rem;
    }
    if (r.rem > r.hi - r.o) r.rem = r.hi - r.o;
    if (r.st == ST_LIT && r.rem == 0) r.st = CODEC == 0 ? ST_TOKEN : ST_MATCH;   // (not produced by the walk; kept total)
    if (r.st == ST_COPY && r.rem == 0) r.st = ST_TOKEN;
}

// Decodes the next element (Snappy) at r.ip.  The walk kernel has validated the stream: no checks here.
G6_HD void snappy_next(Range& r, const Block& b) {
    const uint32_t t = rd32u(b.in, r.ip);
    const uint32_t tag = t & 0xFFu, type = tag & 3u;
    uint32_t len;
    if (type == 0) {
        len = (tag >> 2) + 1;
        uint32_t hdr = 1;
        if (len > 60) {
            const uint32_t nb = len - 60;
            uint32_t v = (t >> 8) & (nb >= 3 ? 0xFFFFFFu : low_mask(8 * nb));
            if (nb == 4) v |= (uint32_t)b.in[r.ip + 4] << 24;
            len = v + 1;
            hdr = 1 + nb;
        }
        r.src = r.ip + hdr;
        r.ip += hdr + len;
        r.st = ST_LIT;
    } else if (type == 1) {
        len = 4 + ((tag >> 2) & 7u);
        r.src = ((tag >> 5) << 8) | ((t >> 8) & 0xFFu);
        r.ip += 2;
        r.st = ST_COPY;
    } else {
        len = (tag >> 2) + 1;
        r.src = (t >> 8) & 0xFFFFu;
        r.ip += 3;
        r.st = ST_COPY;
    }
    r.rem = len < r.hi - r.o ? len : r.hi - r.o;
}

// LZ4: a sequence is a literal run (token, length extension, bytes) and a match part (offset, length extension).
G6_HD void lz4_next(Range& r, const Block& b) {
    if (r.st == ST_TOKEN) {
        const uint32_t token = b.in[r.ip];
        uint32_t p = r.ip + 1, ll = token >> 4;
        r.nib = token & 15u;
        if (ll == 15) {
            uint32_t x;
            do { x = b.in[p++]; ll += x; } while (x == 255);
        }
        r.src = p;
        r.ip = p + ll;
        r.st = ST_MATCH;
        if (ll) {
            r.st = ST_LIT;
            r.rem = ll < r.hi - r.o ? ll : r.hi - r.o;
            return;
        }
    }
    // match part at ip (never reached at the end of the block: the range is complete before)
    const uint32_t off = (uint32_t)b.in[r.ip] | ((uint32_t)b.in[r.ip + 1] << 8);
    uint32_t p = r.ip + 2, ml = r.nib + 4;
    if (r.nib == 15) {
        uint32_t x;
        do { x = b.in[p++]; ml += x; } while (x == 255);
    }
    r.ip = p;
    r.src = off;
    r.st = ST_COPY;
    r.rem = ml < r.hi - r.o ? ml : r.hi - r.o;
}

#if defined(__CUDA_ARCH__)
#define G6_FENCE() __threadfence_block()
#define G6_VLOAD(p) (*reinterpret_cast<const volatile uint32_t*>(p))
#define G6_VSTORE(p, v) (*reinterpret_cast<volatile uint32_t*>(p) = (v))
#else
#define G6_FENCE() ((void)0)
#define G6_VLOAD(p) (*(p))
#define G6_VSTORE(p, v) (*(p) = (v))
#endif

// One iteration of a lane: the next element is decoded if the current one is finished, and at most one piece (<= 8 bytes,
// inside one 32-byte span of the output) is moved.
template <int CODEC>
G6_HD int range_step(Range& r, const Block& b, uint32_t k) {
    if (r.rem == 0) {
        if (r.o >= r.hi) {
            if (!r.pending) return STEP_DONE;
            // next pass: from the first element that was skipped (pieces that are final by now are passed over)
            r.pending = false;
            if (r.from_ck) {
                range_start<CODEC>(r, b, k);
                return STEP_BUSY;
            }
            r.ip = r.p_ip;
            r.o = r.p_o;
            r.st = r.p_st & 0xFFu;
            r.nib = r.p_st >> 8;
        }
        r.e_ip = r.ip;
        r.e_o = r.o;
        if (CODEC == 0) {
            r.e_st = ST_TOKEN;
            snappy_next(r, b);
        } else {
            if (r.st == ST_LIT) r.st = ST_MATCH;        // literals of the sequence are out: its match part is next
            else if (r.st == ST_COPY) r.st = ST_TOKEN;
            r.e_st = r.st | (r.nib << 8);
            lz4_next(r, b);
        }
        // an element that lies inside one bitmap word and is final already is passed over whole
        const uint32_t q = r.o & 31u;
        if (q + r.rem <= 32u) {
            const uint32_t m = low_mask(r.rem) << q;
            if ((G6_VLOAD(&b.bits[r.o >> 5]) & m) == m) {
                r.o += r.rem;
                r.rem = 0;
                return STEP_BUSY;
            }
        }
    }
    const uint32_t q = r.o & 31u;
    uint32_t n = r.rem < 8u ? r.rem : 8u;
    if (n > 32u - q) n = 32u - q;
    if (r.st == ST_COPY && r.src < n) n = r.src;        // overlapping copy: a piece never reads what it writes
    const uint32_t own = G6_VLOAD(&b.bits[r.o >> 5]);
    if ((own >> q) & 1u) {                              // this piece is final already (an earlier pass)
        r.o += n;
        r.rem -= n;
        if (r.st == ST_LIT) r.src += n;
        return STEP_BUSY;
    }
    uint64_t v;
    if (r.st == ST_COPY) {
        const uint32_t s = r.o - r.src;
        const uint32_t ready = funnel_r(G6_VLOAD(&b.bits[s >> 5]), G6_VLOAD(&b.bits[(s >> 5) + 1]), s & 31u);
        const uint32_t need = low_mask(n);
        if ((ready & need) != need) {                   // the source is not final: leave the rest of the element for the next pass
            if (!r.pending) {
                r.pending = true;
                r.from_ck = r.e_o < r.lo || (r.e_o == r.lo && r.e_ip == 0xFFFFFFFFu);
                r.p_ip = r.e_ip;
                r.p_o = r.e_o;
                r.p_st = r.e_st;
            }
            r.o += r.rem;
            r.rem = 0;
            return STEP_BUSY;
        }
        v = rd64u(b.win, s);
    } else {
        v = rd64u(b.in, r.src);
        r.src += n;
    }
    put_bytes(b.win, r.o, v, n);
    G6_FENCE();                                          // the bytes before the bits that announce them
    G6_VSTORE(&b.bits[r.o >> 5], own | (low_mask(n) << q));
    r.o += n;
    r.rem -= n;
    return STEP_BUSY;
}

// ---- the walk (host restatement; the device kernel in lz_decode6.cu follows the same rules with a ring-buffered input) ----
// Validates one block and writes its checkpoints.  Returns the block's uncompressed length, or 0 when the block is not
// for this path (malformed, 4-byte-offset copies, larger than the window, ...): those go to the generation-2 kernel,
// which owns all error reporting.
struct WalkOut {
    uint32_t* ck;       // 2 * NR words
    G6_HDM void emit(uint32_t k, uint32_t w0, uint32_t w1) { ck[2 * k] = w0; ck[2 * k + 1] = w1; }
};

// boundaries strictly inside (o, o + len): the element covers them
G6_HD void emit_inside(WalkOut& w, uint32_t o, uint32_t len, uint32_t st, uint32_t ip_lit_or_next, uint32_t off, uint32_t nib) {
    for (uint32_t B = (o / R + 1) * R; B < o + len; B += R) {
        const uint32_t done = B - o;
        w.emit(B / R, ck_w0(st == ST_LIT ? ip_lit_or_next + done : ip_lit_or_next, st, nib), ck_w1(len - done, off));
    }
}

static inline uint32_t snappy_walk_host(const uint8_t* src, uint32_t n, uint64_t dcap, uint32_t max_in, WalkOut w) {
    if (n < 1 || n > max_in) return 0;
    uint64_t v = 0;
    uint32_t ip = 0;
    bool done = false;
    for (int i = 0; i < 5 && ip < n; i++) {
        const uint32_t x = src[ip++];
        v |= (uint64_t)(x & 0x7f) << (7 * i);
        if (!(x & 0x80)) { done = true; break; }
    }
    if (!done || v < 1 || v > dcap || v > MAXU) return 0;
    const uint32_t ulen = (uint32_t)v;
    uint32_t o = 0;
    while (ip < n) {
        if (o >= ulen) return 0;
        if (o % R == 0) w.emit(o / R, ck_w0(ip, ST_TOKEN, 0), 0);
        const uint32_t tag = src[ip], type = tag & 3;
        if (type == 0) {
            uint32_t len = (tag >> 2) + 1, hdr = 1;
            if (len > 60) {
                const uint32_t nb = len - 60;
                if (nb > n - ip - 1) return 0;
                uint32_t x = 0;
                for (uint32_t i = 0; i < nb; i++) x |= (uint32_t)src[ip + 1 + i] << (8 * i);
                if (x >= MAXU) return 0;
                len = x + 1;
                hdr = 1 + nb;
            }
            if (hdr > n - ip || len > n - ip - hdr || len > ulen - o) return 0;
            emit_inside(w, o, len, ST_LIT, ip + hdr, 0, 0);
            ip += hdr + len;
            o += len;
        } else if (type == 3) {
            return 0;
        } else {
            const uint32_t adv = type == 1 ? 2 : 3;
            if (adv > n - ip) return 0;
            const uint32_t len = type == 1 ? 4 + ((tag >> 2) & 7) : (tag >> 2) + 1;
            const uint32_t off = type == 1 ? ((tag >> 5) << 8) | src[ip + 1] : (uint32_t)src[ip + 1] | ((uint32_t)src[ip + 2] << 8);
            if (off == 0 || off > o || len > ulen - o) return 0;
            ip += adv;
            emit_inside(w, o, len, ST_COPY, ip, off, 0);
            o += len;
        }
    }
    return o == ulen ? ulen : 0;
}

// LZ4 block: exactly the acceptance rules of LZ4_decompress_safe as oracle/lz4.c restates them (MFLIMIT 12, LASTLITERALS 5,
// where the length extensions must stop), and the decoded size must equal the capacity (the batch API passes the known
// size); a block that is rejected here — including a legal one that decodes to less than the capacity — goes to generation 2.
static inline uint32_t lz4_walk_host(const uint8_t* src, uint32_t n, uint64_t dcap, uint32_t max_in, WalkOut w) {
    if (n < 1 || n > max_in || dcap < 1 || dcap > MAXU) return 0;
    const uint32_t cap = (uint32_t)dcap;
    uint32_t ip = 0, o = 0;
    for (;;) {
        if (ip >= n) return 0;
        if (o % R == 0 && o < cap) w.emit(o / R, ck_w0(ip, ST_TOKEN, 0), 0);
        const uint32_t token = src[ip++];
        uint64_t ll = token >> 4;
        if (ll == 15) {
            if (n < 15 || ip >= n - 15) return 0;
            uint32_t x;
            do {
                x = src[ip++];
                ll += x;
                if (ip > n - 15) return 0;
            } while (x == 255);
        }
        const uint32_t nib = token & 15;
        if ((uint64_t)o + ll + 12 > cap || (uint64_t)ip + ll + 8 > n) {     // the final, literal-only sequence
            if ((uint64_t)ip + ll != n || (uint64_t)o + ll != cap) return 0;
            if (ll) emit_inside(w, o, (uint32_t)ll, ST_LIT, ip, 0, nib);
            return cap;
        }
        if (ll) emit_inside(w, o, (uint32_t)ll, ST_LIT, ip, 0, nib);
        ip += (uint32_t)ll;
        o += (uint32_t)ll;
        if (o % R == 0 && ll != 0) w.emit(o / R, ck_w0(ip, ST_MATCH, nib), 0);
        const uint32_t off = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8);
        ip += 2;
        uint64_t ml = nib;
        if (nib == 15) {
            uint32_t x;
            do {
                x = src[ip++];
                ml += x;
                if (ip > n - 4) return 0;
            } while (x == 255);
        }
        ml += 4;
        if (off == 0 || off > o || (uint64_t)o + ml + 5 > cap) return 0;
        emit_inside(w, o, (uint32_t)ml, ST_COPY, ip, off, 0);
        o += (uint32_t)ml;
    }
}

}  // namespace g6
}  // namespace cj
