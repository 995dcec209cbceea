// zstd_huf.cuh — Huffman literal stage of the Zstandard encoder (RFC 8878 3.1.1.3.1, 4.2): code lengths from a
// histogram (limited to 11 bits), the tree description, canonical codes, and the backward bitstream of one literal stream.
// Serial routines (one lane each on the device), written host + device so that tests/test_zstd_huf_host.py can build
// literal-only frames on the CPU and have libzstd decode them before any GPU time is spent.
//
// Reference entry point: cramjam.zstd.compress -> libcramjam::zstd::compress -> libzstd (src/zstd.rs:37-64).  Compressed
// bytes are not pinned by the reference: the stage only has to be format-valid and round-trip exact.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZH_HD __host__ __device__
#else
#define ZH_HD
#endif

namespace cj {
namespace zh {

constexpr int MAXBITS = 11;      // Huffman_Tree max depth for literals (RFC 8878 4.2.1)
constexpr int MAXSYM_DIRECT = 128;   // the direct (4 bits per weight) tree description holds at most 128 weights

// Code lengths for the symbols with count > 0 (nbits[s] = 0 for the others).  Returns the longest length used (<= MAXBITS),
// 0 when fewer than two symbols occur.  The code is complete (Kraft sum exactly 1), as the format requires.
ZH_HD inline int build_lengths(const uint32_t* count, uint8_t* nbits) {
    uint16_t order[256];     // used symbols, ascending by count
    int n = 0;
    for (int s = 0; s < 256; s++) {
        nbits[s] = 0;
        if (count[s]) order[n++] = (uint16_t)s;
    }
    if (n < 2) return 0;
    for (int i = 1; i < n; i++) {   // insertion sort: the alphabet is small and nearly any order is fine
        const uint16_t x = order[i];
        const uint32_t cx = count[x];
        int j = i - 1;
        while (j >= 0 && (count[order[j]] > cx)) { order[j + 1] = order[j]; j--; }
        order[j + 1] = x;
    }
    // two-queue Huffman: leaves in `order`, internal nodes appended in creation order (their weights are non-decreasing)
    uint32_t iw[256];        // internal node weights
    uint16_t parent[512];    // [0, n): leaves by rank, [n, 2n-1): internal nodes
    int li = 0, ii = 0, ni = 0;
    while (ni < n - 1) {
        uint32_t w = 0;
        for (int k = 0; k < 2; k++) {
            const bool leaf = li < n && (ii >= ni || count[order[li]] <= iw[ii]);
            if (leaf) { w += count[order[li]]; parent[li++] = (uint16_t)(n + ni); }
            else { w += iw[ii]; parent[n + ii++] = (uint16_t)(n + ni); }
        }
        iw[ni++] = w;
    }
    // depths: the root is the last internal node; internal nodes point forward, so one backward pass resolves them
    uint8_t depth[512];
    depth[n + ni - 1] = 0;
    for (int k = n + ni - 2; k >= 0; k--) depth[k] = (uint8_t)(depth[parent[k]] + 1);
    int blc[64];
    for (int i = 0; i < 64; i++) blc[i] = 0;
    int maxd = 0;
    for (int r = 0; r < n; r++) {
        int d = depth[r] > 63 ? 63 : depth[r];
        blc[d]++;
        if (d > maxd) maxd = d;
    }
    // length limiting (the classic adjustment of JPEG Annex K.2 / deflate encoders): move leaves up until nothing is deeper than MAXBITS
    for (int i = maxd; i > MAXBITS; i--) {
        while (blc[i] > 0) {
            int j = i - 2;
            while (blc[j] == 0) j--;
            blc[i] -= 2;
            blc[i - 1] += 1;
            blc[j + 1] += 2;
            blc[j] -= 1;
        }
    }
    if (maxd > MAXBITS) maxd = MAXBITS;
    // hand the lengths out: rarest symbols get the longest codes
    int r = 0;
    for (int len = maxd; len >= 1; len--)
        for (int k = 0; k < blc[len]; k++) nbits[order[r++]] = (uint8_t)len;
    while (maxd > 1 && blc[maxd] == 0) maxd--;
    return maxd;
}

// Canonical code values exactly as libzstd's HUF_buildCTable assigns them (the decoder's table order: longest codes first,
// symbols of one length in natural order).  code[s] = value | nbits << 12.
ZH_HD inline void assign_codes(const uint8_t* nbits, int maxbits, uint16_t* code) {
    uint16_t per[MAXBITS + 2], val[MAXBITS + 2];
    for (int i = 0; i <= MAXBITS + 1; i++) { per[i] = 0; val[i] = 0; }
    for (int s = 0; s < 256; s++) per[nbits[s]]++;
    uint16_t mn = 0;
    for (int nb = maxbits; nb > 0; nb--) {
        val[nb] = mn;
        mn = (uint16_t)((mn + per[nb]) >> 1);
    }
    for (int s = 0; s < 256; s++) code[s] = nbits[s] ? (uint16_t)(val[nbits[s]]++ | (nbits[s] << 12)) : 0;
}

// Direct tree description: header byte 127 + number of weights, then 4-bit weights, two per byte, for symbols
// 0 .. last-1 (the last present symbol's weight is implied).  Returns the bytes written, 0 if the alphabet does not fit.
ZH_HD inline int write_tree_direct(const uint8_t* nbits, int maxbits, uint8_t* out) {
    int last = 255;
    while (last > 0 && nbits[last] == 0) last--;
    if (last > MAXSYM_DIRECT || last < 1) return 0;
    out[0] = (uint8_t)(127 + last);
    for (int s = 0; s < last; s += 2) {
        const int w0 = nbits[s] ? maxbits + 1 - nbits[s] : 0;
        const int w1 = (s + 1 < last && nbits[s + 1]) ? maxbits + 1 - nbits[s + 1] : 0;
        out[1 + s / 2] = (uint8_t)((w0 << 4) | w1);
    }
    return 1 + (last + 1) / 2;
}

// One Huffman stream: symbols are coded from the last to the first into an LSB-first bit container, closed by a 1 bit
// (the decoder reads the stream backwards from that bit).  Returns the bytes written.
ZH_HD inline uint32_t encode_stream(const uint8_t* lit, uint32_t n, const uint16_t* code, uint8_t* out) {
    uint64_t acc = 0;
    uint32_t nb = 0, pos = 0;
    for (uint32_t i = n; i-- > 0;) {
        const uint32_t c = code[lit[i]];
        acc |= (uint64_t)(c & 0xFFFu) << nb;
        nb += c >> 12;
        if (nb >= 32) {
            out[pos] = (uint8_t)acc; out[pos + 1] = (uint8_t)(acc >> 8); out[pos + 2] = (uint8_t)(acc >> 16); out[pos + 3] = (uint8_t)(acc >> 24);
            pos += 4;
            acc >>= 32;
            nb -= 32;
        }
    }
    acc |= 1ull << nb;
    nb += 1;
    while (nb > 0) {
        out[pos++] = (uint8_t)acc;
        acc >>= 8;
        nb = nb > 8 ? nb - 8 : 0;
    }
    return pos;
}

// Size of one stream in bytes without writing it (bit lengths summed from the histogram of the segment is not available per
// segment, so this walks the symbols): used by the host test only.
ZH_HD inline uint32_t stream_bits(const uint8_t* lit, uint32_t n, const uint16_t* code) {
    uint32_t bits = 1;
    for (uint32_t i = 0; i < n; i++) bits += code[lit[i]] >> 12;
    return bits;
}

// Literals_Section_Header of a Compressed_Literals_Block (type 2).  Returns the header length; `four` tells whether the
// payload is four streams behind a 6-byte jump table.
ZH_HD inline int header_len(uint32_t regen) { return regen < 1024 ? 3 : (regen < 16384 ? 4 : 5); }
ZH_HD inline void write_header(uint8_t* out, uint32_t regen, uint32_t comp, bool four) {
    if (regen < 1024) {   // size format 00 (one stream) / 01 (four streams), 10 bits each
        const uint32_t v = 2u | ((four ? 1u : 0u) << 2) | (regen << 4) | (comp << 14);
        out[0] = (uint8_t)v; out[1] = (uint8_t)(v >> 8); out[2] = (uint8_t)(v >> 16);
    } else if (regen < 16384) {   // size format 10: four streams, 14 bits each
        const uint32_t v = 2u | (2u << 2) | (regen << 4) | (comp << 18);
        out[0] = (uint8_t)v; out[1] = (uint8_t)(v >> 8); out[2] = (uint8_t)(v >> 16); out[3] = (uint8_t)(v >> 24);
    } else {   // size format 11: four streams, 18 bits each
        const uint64_t v = 2ull | (3ull << 2) | ((uint64_t)regen << 4) | ((uint64_t)comp << 22);
        out[0] = (uint8_t)v; out[1] = (uint8_t)(v >> 8); out[2] = (uint8_t)(v >> 16); out[3] = (uint8_t)(v >> 24); out[4] = (uint8_t)(v >> 32);
    }
}

}  // namespace zh
}  // namespace cj
