// zstd_decode.cu — Zstandard frame decode kernel (sm_100a), RFC 8878.
//
// Replaces libcramjam::zstd::decompress -> zstd::stream::read::Decoder -> libzstd
// (reference src/zstd.rs:23-28,67-70).  One warp per frame stream (blocks of a frame depend on each
// other through the window, the repeat-offset history and the repeat/treeless tables, so the frame is
// the parallel unit; concatenated and skippable frames are walked in the same loop).
//
// Per compressed block: the Huffman table (<= 2048 x u16) and the three FSE tables (512 / 256 / 512 x
// u32) are built in shared memory; the four Huffman literal streams are decoded by four lanes at
// once into a per-warp literal buffer in global memory; the sequence bitstream — three interleaved
// FSE states read backwards, the strictly serial part of the format — is decoded warp-uniformly and
// every sequence is executed by the whole warp through the same shared-memory output ring the LZ4 /
// Snappy kernels use (literal run and match copied 32 bytes per instruction, far matches re-read from
// drained output through L2).  The content checksum (XXH64) is verified on the device.
// Acceptance rules and status codes follow oracle/zstd_dec.c.
#include "internal.h"
#include "lz_decode.cuh"

namespace cj {

#ifndef CJ_ZS_WARPS
#define CJ_ZS_WARPS 10
#endif
#ifndef CJ_ZS_CTAS
#define CJ_ZS_CTAS 2
#endif
constexpr int ZS_WARPS = CJ_ZS_WARPS;   // 10 warps x 11.1 KiB: two CTAs (20 warps) fit an SM
constexpr uint32_t ZS_BLOCK_MAX = 128 * 1024;
constexpr size_t ZS_LIT_STRIDE = ZS_BLOCK_MAX + 64;  // per-warp literal buffer in global scratch

// per-warp shared memory layout.  The Huffman table (needed while a block's literals are decoded) and the three FSE tables (needed
// while its sequences are decoded) share one region: the kernel is bound by the latency of its serial bit chains, so resident warps
// are what it needs (15 -> 20 per SM).  What a later block may ask to reuse (treeless literals, Repeat_Mode tables) is kept as its
// description — Huffman weights, normalized counts — and rebuilt when the region held the other kind in between.
constexpr int ZS_OFF_U = ORING;                         // union: 2048 x u16 Huffman table | LL 512 + OF 256 + ML 512 x u32 FSE tables
constexpr int ZS_OFF_HUF = ZS_OFF_U;
constexpr int ZS_OFF_LL = ZS_OFF_U;
constexpr int ZS_OFF_OF = ZS_OFF_LL + 512 * 4;
constexpr int ZS_OFF_ML = ZS_OFF_OF + 256 * 4;
constexpr int ZS_OFF_TMP = ZS_OFF_ML + 512 * 4;         // 256 x i16 frequencies + 256 x u16 next-state counters + 272 weights + 64 x u32 weight-FSE table
constexpr int ZS_OFF_SAVE = ZS_OFF_TMP + 512 + 512 + 272 + 256;   // saved descriptions: 256 Huffman weights | LL 36 + OF 32 + ML 53 (+7 pad) x i16 counts
constexpr int ZS_SMEM_WARP = ZS_OFF_SAVE + 256 + (36 + 32 + 60) * 2;
static_assert(ZS_SMEM_WARP % 16 == 0, "per-warp shared memory must stay 16-byte aligned");

struct FseTab {
    uint32_t* t;  // entry = sym | nbits << 8 | base << 16
    int al;
    bool ok;      // a table has been set up in this frame (Repeat_Mode may refer to it)
    bool live;    // ... and it is in shared memory right now (the Huffman table has not been built over it since)
    int kind;     // how it was set up: 0 predefined, 1 RLE (sym), 2 described (counts saved)
    int sym, nsym;
    int16_t* saved;   // normalized counts of a described table
};

__device__ __forceinline__ int hibit(uint32_t v) { return 31 - __clz(v); }

// ---- bit readers over global memory (warp-uniform or per-lane; plain byte loads, L1 resident) ----
struct FwdBits {
    const uint8_t* p;
    uint32_t n, bit;
    __device__ __forceinline__ uint32_t read(int nb) {
        uint32_t v = 0;
        for (int i = 0; i < nb; i++) {
            const uint32_t byte = bit >> 3;
            const uint32_t x = byte < n ? (__ldg(p + byte) >> (bit & 7)) & 1u : 0u;
            v |= x << i;
            bit++;
        }
        return v;
    }
};

// Backward bit reader.  A 64-bit window of the stream is kept in registers (two aligned 8-byte loads + a funnel
// shift per refill, every ~40-56 consumed bits); a read is a shift and a mask.  Only aligned words that hold at least
// one byte of the stream are loaded (CJ_DEVICE callers need not pad their units), the surplus bits are never used.
struct BackBits {
    const uint8_t* p;
    const uint8_t* e;   // one past the stream's last byte
    int32_t pos;   // bits still unread (may go negative = over-read)
    int32_t wbit;  // stream bit index of win's bit 0 (multiple of 8)
    uint64_t win;
    __device__ __forceinline__ bool init(const uint8_t* p_, uint32_t n_) {
        p = p_;
        e = p_ + n_;
        win = 0;
        pos = 0;
        wbit = 0;
        if (n_ == 0) return false;
        const uint32_t last = __ldg(p_ + n_ - 1);
        if (last == 0) return false;
        pos = (int32_t)(n_ - 1) * 8 + hibit(last);
        wbit = pos;  // forces a refill on the first read
        return true;
    }
    __device__ __forceinline__ void refill() {  // window = the 8 bytes ending at the byte that holds bit pos-1
        int32_t wb = ((pos - 1) >> 3) - 7;
        if (wb < 0) wb = 0;
        const uint8_t* a = p + wb;
        const uint64_t* a8 = reinterpret_cast<const uint64_t*>((uintptr_t)a & ~(uintptr_t)7);
        const uint32_t sh = (uint32_t)((uintptr_t)a & 7) * 8;
        const uint64_t lo = __ldg(a8), hi = (sh && reinterpret_cast<const uint8_t*>(a8 + 1) < e) ? __ldg(a8 + 1) : 0ull;
        win = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
        wbit = wb * 8;
    }
    // bits [pos-nb, pos) as an integer (most significant = highest position), zero-filled below bit 0; nb <= 32
    __device__ __forceinline__ uint32_t peek(int nb) {
        const int32_t lo = pos - nb;
        if (lo < wbit) {
            if (lo < 0) {  // reading past the beginning: zero-fill the missing low bits
                if (pos <= 0) return 0;
                if (wbit != 0) refill();
                const uint64_t v = win & ((1ull << pos) - 1);
                return (uint32_t)(v << (uint32_t)(-lo));
            }
            refill();
        }
        // 32-bit funnel shift over the two window halves instead of a 64-bit shift + mask
        const uint32_t sft = (uint32_t)(lo - wbit);
        const uint32_t w0 = (uint32_t)win, w1 = (uint32_t)(win >> 32);
        const uint32_t v = (sft & 32) ? (w1 >> (sft & 31)) : __funnelshift_r(w0, w1, sft);
        return v & __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)nb);   // low nb bits (nb <= 32; the clamped funnel shift builds the mask in one instruction, bfe is emulated)
    }
    __device__ __forceinline__ uint32_t read(int nb) {
        const uint32_t v = peek(nb);
        pos -= nb;
        return v;
    }
    // The next `total` bits (total <= 56 <= pos) as the top of a 64-bit value {hi, lo}: one window check for all the fields of a
    // sequence, which are then peeled off the top with take().  Consumes the bits.
    __device__ __forceinline__ void top(uint32_t total, uint32_t& hi, uint32_t& lo) {
        if ((int32_t)(pos - (int32_t)total) < wbit) refill();   // afterwards the window holds at least the 57 bits below pos
        const uint32_t s = 64u - (uint32_t)(pos - wbit);       // window bit 63 - s is stream bit pos - 1; 0 <= s <= 63
        const uint32_t w0 = (uint32_t)win, w1 = (uint32_t)(win >> 32);
        const bool big = (s & 32u) != 0;
        hi = big ? (w0 << (s & 31u)) : __funnelshift_l(w0, w1, s);
        lo = big ? 0u : (w0 << (s & 31u));
        pos -= (int32_t)total;
    }
    static __device__ __forceinline__ uint32_t take(uint32_t& hi, uint32_t& lo, uint32_t nb) {   // nb <= 31
        const uint32_t v = __funnelshift_rc(hi, 0u, 32u - nb);   // hi >> (32 - nb), 0 for nb == 0
        hi = __funnelshift_l(lo, hi, nb);
        lo <<= nb;
        return v;
    }
};

// ---- FSE ----
// freq[] (int16, nsym entries) in shared memory -> decode table.  Executed by lane 0.
__device__ bool fse_build_lane0(uint32_t* tab, const int16_t* freq, uint16_t* next, int nsym, int al) {
    const int size = 1 << al;
    int high = size;
    for (int s = 0; s < nsym; s++)
        if (freq[s] == -1) { tab[--high] = (uint32_t)s; next[s] = 1; }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < nsym; s++) {
        if (freq[s] <= 0) continue;
        next[s] = (uint16_t)freq[s];
        for (int i = 0; i < freq[s]; i++) {
            tab[pos] = (uint32_t)s;
            do { pos = (pos + step) & mask; } while (pos >= high);
        }
    }
    if (pos != 0) return false;
    for (int i = 0; i < size; i++) {
        const uint32_t sym = tab[i] & 0xff;
        const uint32_t ns = next[sym]++;
        const int nb = al - hibit(ns);
        tab[i] = sym | ((uint32_t)nb << 8) | ((((ns << nb) - (uint32_t)size) & 0xffff) << 16);
    }
    return true;
}

// Reads an FSE table description at p (lane 0); returns bytes consumed or -1.
__device__ int fse_read_desc_lane0(uint32_t* tab, int* al_out, const uint8_t* p, uint32_t n, int max_al, int max_sym, int16_t* freq, uint16_t* next, int* nsym_out = nullptr) {
    if (n == 0) return -1;
    FwdBits b{p, n, 0};
    const int al = 5 + (int)b.read(4);
    if (al > max_al) return -1;
    int remaining = 1 << al, s = 0;
    while (remaining > 0 && s <= max_sym) {
        const int nb = hibit((uint32_t)remaining + 1) + 1;
        uint32_t val = b.read(nb);
        const uint32_t lower = (1u << (nb - 1)) - 1;
        const uint32_t thresh = (1u << nb) - 1 - ((uint32_t)remaining + 1);
        if ((val & lower) < thresh) { b.bit--; val &= lower; }
        else if (val > lower) val -= thresh;
        const int proba = (int)val - 1;
        remaining -= proba < 0 ? -proba : proba;
        freq[s++] = (int16_t)proba;
        if (proba == 0) {
            uint32_t rep = b.read(2);
            for (;;) {
                for (uint32_t i = 0; i < rep && s <= max_sym; i++) freq[s++] = 0;
                if (rep == 3) rep = b.read(2); else break;
            }
        }
        if ((b.bit + 7) / 8 > n) return -1;
    }
    if (remaining != 0 || s > max_sym + 1) return -1;
    const uint32_t used = (b.bit + 7) / 8;
    if (used > n) return -1;
    if (!fse_build_lane0(tab, freq, next, s, al)) return -1;
    *al_out = al;
    if (nsym_out) *nsym_out = s;
    return (int)used;
}

__constant__ int16_t ZS_LL_DEFAULT[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
__constant__ int16_t ZS_ML_DEFAULT[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
__constant__ int16_t ZS_OF_DEFAULT[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
// baseline value | number of extra bits << 24, per literal-length / match-length code (RFC 8878 3.1.1.3.2.1.1): one constant load per code
__constant__ uint32_t ZS_LL_PK[36] = {0x0, 0x1, 0x2, 0x3, 0x4, 0x5, 0x6, 0x7, 0x8, 0x9, 0xa, 0xb, 0xc, 0xd, 0xe, 0xf, 0x1000010, 0x1000012, 0x1000014, 0x1000016, 0x2000018, 0x200001c, 0x3000020, 0x3000028, 0x4000030, 0x6000040, 0x7000080, 0x8000100, 0x9000200, 0xa000400, 0xb000800, 0xc001000, 0xd002000, 0xe004000, 0xf008000, 0x10010000};
__constant__ uint32_t ZS_ML_PK[53] = {0x3, 0x4, 0x5, 0x6, 0x7, 0x8, 0x9, 0xa, 0xb, 0xc, 0xd, 0xe, 0xf, 0x10, 0x11, 0x12, 0x13, 0x14, 0x15, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x1b, 0x1c, 0x1d, 0x1e, 0x1f, 0x20, 0x21, 0x22, 0x1000023, 0x1000025, 0x1000027, 0x1000029, 0x200002b, 0x200002f, 0x3000033, 0x300003b, 0x4000043, 0x4000053, 0x5000063, 0x7000083, 0x8000103, 0x9000203, 0xa000403, 0xb000803, 0xc001003, 0xd002003, 0xe004003, 0xf008003, 0x10010003};

// Sets up one sequence table according to its mode.  Returns bytes consumed or -1.  Warp-uniform result.
__device__ int seq_table(FseTab& t, int mode, const uint8_t* p, uint32_t n, const int16_t* def, int def_n, int def_al, int max_al, int max_sym,
                         int16_t* freq, uint16_t* next, int lane) {
    int consumed = -1, al = t.al, sym = t.sym, nsym = t.nsym;
    if (mode == 3) {   // Repeat_Mode: the previous table — still there, or rebuilt from what was saved of it
        if (!t.ok) return -1;
        if (t.live) return 0;
        int good = 0;
        if (lane == 0) {
            if (t.kind == 1) { t.t[0] = (uint32_t)t.sym; good = 1; }
            else {
                const int16_t* from = t.kind == 0 ? def : t.saved;
                const int cnt = t.kind == 0 ? def_n : t.nsym;
                for (int i = 0; i < cnt; i++) freq[i] = from[i];
                good = fse_build_lane0(t.t, freq, next, cnt, t.al) ? 1 : 0;
            }
        }
        __syncwarp();
        good = __shfl_sync(FULL, good, 0);
        if (!good) return -1;
        t.live = true;
        return 0;
    }
    if (lane == 0) {
        if (mode == 0) {
            for (int i = 0; i < def_n; i++) freq[i] = def[i];
            if (fse_build_lane0(t.t, freq, next, def_n, def_al)) { consumed = 0; al = def_al; }
        } else if (mode == 1) {
            if (n >= 1 && __ldg(p) <= max_sym) { sym = __ldg(p); t.t[0] = (uint32_t)sym; consumed = 1; al = 0; }
        } else {
            consumed = fse_read_desc_lane0(t.t, &al, p, n, max_al, max_sym, freq, next, &nsym);
            if (consumed >= 0) for (int i = 0; i < nsym; i++) t.saved[i] = freq[i];   // freq[] still holds the counts as read
        }
    }
    __syncwarp();
    consumed = __shfl_sync(FULL, consumed, 0);
    al = __shfl_sync(FULL, al, 0);
    sym = __shfl_sync(FULL, sym, 0);
    nsym = __shfl_sync(FULL, nsym, 0);
    if (consumed >= 0) { t.al = al; t.ok = true; t.live = true; t.kind = mode; t.sym = sym; t.nsym = nsym; }
    return consumed;
}

// ---- Huffman ----
// w[0..nw) explicit weights in shared memory (room for one more); builds huf[] (lane 0).  Returns max_bits or -1.
__device__ int huf_build_lane0(uint16_t* huf, uint8_t* w, int nw) {
    uint32_t sum = 0;
    for (int i = 0; i < nw; i++) {
        if (w[i] > 11) return -1;
        if (w[i]) sum += 1u << (w[i] - 1);
    }
    if (sum == 0) return -1;
    const int max_bits = hibit(sum) + 1;
    if (max_bits > 11) return -1;
    const uint32_t left = (1u << max_bits) - sum;
    if (left & (left - 1)) return -1;
    w[nw] = (uint8_t)(hibit(left) + 1);
    nw++;
    int pos = 0;
    for (int wt = 1; wt <= max_bits; wt++) {
        for (int s = 0; s < nw; s++) {
            if (w[s] != wt) continue;
            const int cells = 1 << (wt - 1);
            const uint16_t e = (uint16_t)(s | ((max_bits + 1 - wt) << 8));
            for (int i = 0; i < cells; i++) huf[pos + i] = e;
            pos += cells;
        }
    }
    return pos == (1 << max_bits) ? max_bits : -1;
}

// Reads the Huffman tree description (lane 0).  Returns bytes consumed or -1; *max_bits set.
__device__ int huf_read_desc_lane0(uint16_t* huf, int* max_bits, const uint8_t* p, uint32_t n, uint32_t* fse_scratch, int16_t* freq, uint16_t* next, uint8_t* w, uint8_t* w_saved, int* nw_saved) {
    if (n == 0) return -1;
    int nw = 0;
    const int hb = __ldg(p);
    uint32_t used;
    if (hb >= 128) {
        nw = hb - 127;
        const uint32_t nbytes = (uint32_t)(nw + 1) / 2;
        if (1 + nbytes > n) return -1;
        for (int i = 0; i < nw; i++) {
            const uint32_t b = __ldg(p + 1 + i / 2);
            w[i] = (i & 1) ? (b & 15) : (b >> 4);
        }
        used = 1 + nbytes;
    } else {
        if (hb == 0 || 1u + (uint32_t)hb > n) return -1;
        int al = 0;
        const int c = fse_read_desc_lane0(fse_scratch, &al, p + 1, (uint32_t)hb, 6, 255, freq, next);
        if (c < 0 || c >= hb) return -1;
        BackBits b;
        if (!b.init(p + 1 + c, (uint32_t)(hb - c))) return -1;
        uint32_t s1 = b.read(al), s2 = b.read(al);
        if (b.pos < 0) return -1;
        for (;;) {
            if (nw > 253) return -1;
            uint32_t e = fse_scratch[s1];
            w[nw++] = (uint8_t)e;
            s1 = (e >> 16) + b.read((e >> 8) & 0xff);
            if (b.pos < 0) { w[nw++] = (uint8_t)fse_scratch[s2]; break; }
            if (nw > 253) return -1;
            e = fse_scratch[s2];
            w[nw++] = (uint8_t)e;
            s2 = (e >> 16) + b.read((e >> 8) & 0xff);
            if (b.pos < 0) { w[nw++] = (uint8_t)fse_scratch[s1]; break; }
        }
        used = 1 + (uint32_t)hb;
    }
    for (int i = 0; i < nw; i++) w_saved[i] = w[i];   // the explicit weights: what a treeless block rebuilds the table from
    *nw_saved = nw;
    const int mb = huf_build_lane0(huf, w, nw);
    if (mb < 0) return -1;
    *max_bits = mb;
    return (int)used;
}

// One Huffman stream, executed by a single lane.  Returns true on success.
__device__ bool huf_decode_stream(const uint16_t* huf, int max_bits, const uint8_t* p, uint32_t n, uint8_t* out, uint32_t count) {
    BackBits b;
    if (!b.init(p, n)) return false;
    for (uint32_t i = 0; i < count; i++) {
        const uint32_t e = huf[b.peek(max_bits)];
        out[i] = (uint8_t)e;
        b.pos -= (int32_t)(e >> 8);
        if (b.pos < 0) return false;
    }
    return b.pos == 0;
}

struct ZState {
    uint16_t* huf;
    int huf_bits;
    bool huf_ok;     // a Huffman table has been described in this frame (treeless literals may refer to it)
    bool huf_live;   // ... and it is in shared memory right now (no FSE table has been built over it since)
    uint8_t* huf_w;  // its explicit weights, saved
    int huf_nw;
    FseTab ll, of, ml;
    uint32_t rep0, rep1, rep2;
    int16_t* freq;
    uint16_t* next;
    uint8_t* w;
    uint32_t* wfse;  // FSE table of the Huffman weights (accuracy <= 6)
    uint8_t* lit;    // global literal buffer of this warp
};

// Literals section -> (lit_ptr, lit_len); returns bytes consumed or -(status).
__device__ int zs_literals(ZState& z, const uint8_t* p, uint32_t n, const uint8_t** lit_ptr, uint32_t* lit_len, int lane) {
    if (n < 1) return -CJ_ST_TRUNCATED;
    const uint32_t b0 = __ldg(p);
    const int type = b0 & 3, sf = (b0 >> 2) & 3;
    uint32_t hdr, regen, comp = 0;
    int streams = 1;
    if (type < 2) {
        if (sf == 0 || sf == 2) { hdr = 1; regen = b0 >> 3; }
        else if (sf == 1) { if (n < 2) return -CJ_ST_TRUNCATED; hdr = 2; regen = (b0 >> 4) | (__ldg(p + 1) << 4); }
        else { if (n < 3) return -CJ_ST_TRUNCATED; hdr = 3; regen = (b0 >> 4) | (__ldg(p + 1) << 4) | (__ldg(p + 2) << 12); }
        if (regen > ZS_BLOCK_MAX) return -CJ_ST_CORRUPT;
        if (type == 0) {
            if (hdr + regen > n) return -CJ_ST_TRUNCATED;
            *lit_ptr = p + hdr;  // raw literals are used in place
            *lit_len = regen;
            return (int)(hdr + regen);
        }
        if (hdr + 1 > n) return -CJ_ST_TRUNCATED;
        const uint8_t v = __ldg(p + hdr);
        for (uint32_t i = lane; i < regen; i += 32) z.lit[i] = v;
        __syncwarp();
        *lit_ptr = z.lit;
        *lit_len = regen;
        return (int)(hdr + 1);
    }
    if (sf == 0 || sf == 1) {
        if (n < 3) return -CJ_ST_TRUNCATED;
        const uint32_t v = b0 | (__ldg(p + 1) << 8) | (__ldg(p + 2) << 16);
        hdr = 3; regen = (v >> 4) & 0x3ff; comp = (v >> 14) & 0x3ff; streams = sf == 0 ? 1 : 4;
    } else if (sf == 2) {
        if (n < 4) return -CJ_ST_TRUNCATED;
        const uint32_t v = b0 | (__ldg(p + 1) << 8) | (__ldg(p + 2) << 16) | (__ldg(p + 3) << 24);
        hdr = 4; regen = (v >> 4) & 0x3fff; comp = (v >> 18) & 0x3fff; streams = 4;
    } else {
        if (n < 5) return -CJ_ST_TRUNCATED;
        const uint64_t v = (uint64_t)(b0 | (__ldg(p + 1) << 8) | (__ldg(p + 2) << 16) | (__ldg(p + 3) << 24)) | ((uint64_t)__ldg(p + 4) << 32);
        hdr = 5; regen = (uint32_t)((v >> 4) & 0x3ffff); comp = (uint32_t)((v >> 22) & 0x3ffff); streams = 4;
    }
    if (regen > ZS_BLOCK_MAX) return -CJ_ST_CORRUPT;
    if (hdr + comp > n) return -CJ_ST_TRUNCATED;
    const uint8_t* q = p + hdr;
    uint32_t left = comp;
    if (type == 2) {
        int c = -1, mb = 0, nw = 0;
        z.ll.live = z.of.live = z.ml.live = false;   // the Huffman table is built over the FSE tables
        z.huf_live = false;
        if (lane == 0) c = huf_read_desc_lane0(z.huf, &mb, q, left, z.wfse, z.freq, z.next, z.w, z.huf_w, &nw);
        __syncwarp();
        c = __shfl_sync(FULL, c, 0);
        mb = __shfl_sync(FULL, mb, 0);
        nw = __shfl_sync(FULL, nw, 0);
        if (c < 0) { z.huf_ok = false; return -CJ_ST_CORRUPT; }
        z.huf_bits = mb;
        z.huf_nw = nw;
        z.huf_ok = true;
        z.huf_live = true;
        q += c;
        left -= (uint32_t)c;
    } else if (!z.huf_ok) {
        return -CJ_ST_CORRUPT;  // treeless without a previous table
    } else if (!z.huf_live) {   // treeless: the previous table, rebuilt from its saved weights
        int mb = -1;
        z.ll.live = z.of.live = z.ml.live = false;
        if (lane == 0) {
            for (int i = 0; i < z.huf_nw; i++) z.w[i] = z.huf_w[i];
            mb = huf_build_lane0(z.huf, z.w, z.huf_nw);
        }
        __syncwarp();
        mb = __shfl_sync(FULL, mb, 0);
        if (mb < 0) return -CJ_ST_CORRUPT;
        z.huf_bits = mb;
        z.huf_live = true;
    }
    bool ok = true;
    if (streams == 1) {
        if (lane == 0) ok = huf_decode_stream(z.huf, z.huf_bits, q, left, z.lit, regen);
    } else {
        if (left < 6) return -CJ_ST_CORRUPT;
        const uint32_t s1 = __ldg(q) | (__ldg(q + 1) << 8), s2 = __ldg(q + 2) | (__ldg(q + 3) << 8), s3 = __ldg(q + 4) | (__ldg(q + 5) << 8);
        if (6 + s1 + s2 + s3 > left) return -CJ_ST_CORRUPT;
        const uint32_t s4 = left - 6 - s1 - s2 - s3;
        const uint32_t per = (regen + 3) / 4;
        if (per * 3 > regen) return -CJ_ST_CORRUPT;
        if (lane < 4) {
            const uint32_t soff = lane == 0 ? 0 : (lane == 1 ? s1 : (lane == 2 ? s1 + s2 : s1 + s2 + s3));
            const uint32_t slen = lane == 0 ? s1 : (lane == 1 ? s2 : (lane == 2 ? s3 : s4));
            const uint32_t cnt = lane < 3 ? per : regen - 3 * per;
            ok = huf_decode_stream(z.huf, z.huf_bits, q + 6 + soff, slen, z.lit + lane * per, cnt);
        }
    }
    __syncwarp();
    if (__ballot_sync(FULL, !ok)) return -CJ_ST_CORRUPT;
    *lit_ptr = z.lit;
    *lit_len = regen;
    return (int)(hdr + comp);
}

// One compressed block.  Returns CJ_OK or a status; appends to out.
__device__ int32_t zs_block(ZState& z, const uint8_t* p, uint32_t n, OutRing& out, uint32_t cap, int lane) {
    const uint8_t* lit = nullptr;
    uint32_t lit_len = 0;
    int c = zs_literals(z, p, n, &lit, &lit_len, lane);
    if (c < 0) return -c;
    p += c;
    n -= (uint32_t)c;
    if (n < 1) return CJ_ST_TRUNCATED;
    uint32_t nseq, h;
    const uint32_t b0 = __ldg(p);
    if (b0 < 128) { nseq = b0; h = 1; }
    else if (b0 < 255) { if (n < 2) return CJ_ST_TRUNCATED; nseq = ((b0 - 128) << 8) + __ldg(p + 1); h = 2; }
    else { if (n < 3) return CJ_ST_TRUNCATED; nseq = __ldg(p + 1) + (__ldg(p + 2) << 8) + 0x7F00; h = 3; }
    p += h;
    n -= h;
    const uint32_t block_start = out.op;
    uint32_t lp = 0;
    if (nseq == 0) {
        if (n != 0) return CJ_ST_CORRUPT;
    } else {
        if (n < 1) return CJ_ST_TRUNCATED;
        const uint32_t modes = __ldg(p);
        if (modes & 3) return CJ_ST_CORRUPT;
        p++; n--;
        z.huf_live = false;   // the FSE tables are built over the Huffman table (its literals are decoded already)
        c = seq_table(z.ll, (modes >> 6) & 3, p, n, ZS_LL_DEFAULT, 36, 6, 9, 35, z.freq, z.next, lane);
        if (c < 0) return CJ_ST_CORRUPT;
        p += c; n -= (uint32_t)c;
        c = seq_table(z.of, (modes >> 4) & 3, p, n, ZS_OF_DEFAULT, 29, 5, 8, 31, z.freq, z.next, lane);
        if (c < 0) return CJ_ST_CORRUPT;
        p += c; n -= (uint32_t)c;
        c = seq_table(z.ml, (modes >> 2) & 3, p, n, ZS_ML_DEFAULT, 53, 6, 9, 52, z.freq, z.next, lane);
        if (c < 0) return CJ_ST_CORRUPT;
        p += c; n -= (uint32_t)c;
        BackBits b;
        if (!b.init(p, n)) return CJ_ST_CORRUPT;
        uint32_t sl = b.read(z.ll.al), so = b.read(z.of.al), sm = b.read(z.ml.al);
        if (b.pos < 0) return CJ_ST_CORRUPT;
        for (uint32_t i0 = 0; i0 < nseq; i0 += 32) {
            // ---- strictly serial part: up to 32 sequences decoded warp-uniformly; lane k keeps the k-th ----
            uint32_t cnt = min(32u, nseq - i0);
            int32_t derr = CJ_OK;  // a decode error stops the batch; the sequences before it still run first (their errors come first)
            uint32_t LL = 0, ML = 0, OFF = 0, LP = 0;
            for (uint32_t k = 0; k < cnt; k++) {
                // table symbols are range-checked when the tables are built, so codes index the constant tables safely
                const uint32_t el = z.ll.t[sl], eo = z.of.t[so], em = z.ml.t[sm];
                const uint32_t lc = el & 0xff, oc = eo & 0xff, mc = em & 0xff;
                const uint32_t mpk = ZS_ML_PK[mc], lpk = ZS_LL_PK[lc];
                const uint32_t mlb = mpk >> 24, llb = lpk >> 24;
                const bool more = i0 + k + 1 < nseq;   // state updates follow: LL, ML, OF bits in that order (<= 9 + 9 + 8)
                const uint32_t nl = (el >> 8) & 0xff, nm = (em >> 8) & 0xff, no = (eo >> 8) & 0xff;
                const uint32_t total = oc + mlb + llb + (more ? nl + nm + no : 0u);
                uint32_t ovx, mlx, llx;
                const bool one = total <= 56u && (int32_t)total <= b.pos;   // all fields of the sequence from one window (nearly always)
                uint32_t fh = 0, fl = 0;
                if (one) {
                    b.top(total, fh, fl);
                    ovx = BackBits::take(fh, fl, oc);        // offset extra bits (<= 31)
                    mlx = BackBits::take(fh, fl, mlb);       // match-length extra bits (<= 16)
                    llx = BackBits::take(fh, fl, llb);       // literal-length extra bits (<= 16)
                } else {
                    ovx = b.read((int)oc);
                    const uint32_t t = b.read((int)(mlb + llb));
                    mlx = t >> llb;
                    llx = t & ((1u << llb) - 1);
                }
                const uint32_t mlen = (mpk & 0xFFFFFFu) + mlx;
                const uint32_t llen = (lpk & 0xFFFFFFu) + llx;
                bool bad = false;
                uint32_t off;
                if (oc >= 2) {
                    off = (1u << oc) + ovx - 3;
                    z.rep2 = z.rep1; z.rep1 = z.rep0; z.rep0 = off;
                } else {
                    const uint32_t idx = (1u << oc) + ovx - 1 + (llen == 0 ? 1 : 0);
                    if (idx == 0) {
                        off = z.rep0;
                    } else {
                        const uint32_t v = idx == 1 ? z.rep1 : (idx == 2 ? z.rep2 : z.rep0 - 1);
                        bad = v == 0;
                        if (idx > 1) z.rep2 = z.rep1;
                        z.rep1 = z.rep0;
                        z.rep0 = v;
                        off = v;
                    }
                }
                if (more) {
                    if (one) {
                        sl = (el >> 16) + BackBits::take(fh, fl, nl);
                        sm = (em >> 16) + BackBits::take(fh, fl, nm);
                        so = (eo >> 16) + BackBits::take(fh, fl, no);
                    } else {
                        const uint32_t u = b.read((int)(nl + nm + no));
                        sl = (el >> 16) + (u >> (nm + no));
                        sm = (em >> 16) + ((u >> no) & ((1u << nm) - 1));
                        so = (eo >> 16) + (u & ((1u << no) - 1));
                    }
                }
                // every decode-side failure of a sequence is the same status, so one test per sequence is enough
                if (bad || b.pos < 0 || llen > lit_len - lp) { derr = CJ_ST_CORRUPT; cnt = k; break; }
                if ((uint32_t)lane == k) { LL = llen; ML = mlen; OFF = off; LP = lp; }
                lp += llen;
            }
            // ---- lane-parallel execution of those sequences (literals from the literal buffer, matches through the ring) ----
            uint32_t left = cnt;
            while (left) {
                out.make_room();
                const uint32_t k = exec_lanes<RULE_LEN, true>(out, left, LL, lit, LP, 0xFFFFFFFFu, ML, OFF, cap, lane);
                if (k == left) break;
                // element k did not pass the batch checks (too long for a batch, or invalid): serial path, exact status
                const uint32_t jl = __shfl_sync(FULL, LL, k), jm = __shfl_sync(FULL, ML, k), jo = __shfl_sync(FULL, OFF, k), jp = __shfl_sync(FULL, LP, k);
                if ((uint64_t)jl + jm > (uint64_t)(cap - out.op))
                    return ((uint64_t)(out.op - block_start) + jl + jm > ZS_BLOCK_MAX) ? CJ_ST_CORRUPT : CJ_ST_DST_SMALL;
                if (jl) out.put_literals_coherent(lit + jp, jl);
                if (jo > out.op - out.base) return CJ_ST_OFFSET;
                out.put_match(jo, jm);
                const uint32_t sh = k + 1;  // drop the executed prefix: lane j takes over lane j+sh
                LL = __shfl_down_sync(FULL, LL, sh); ML = __shfl_down_sync(FULL, ML, sh);
                OFF = __shfl_down_sync(FULL, OFF, sh); LP = __shfl_down_sync(FULL, LP, sh);
                left -= sh;
            }
            if (derr) return derr;
            if (out.op - block_start > ZS_BLOCK_MAX) return CJ_ST_CORRUPT;
        }
        if (b.pos != 0) return CJ_ST_CORRUPT;
    }
    const uint32_t rest = lit_len - lp;
    if (rest > cap - out.op) return ((uint64_t)(out.op - block_start) + rest > ZS_BLOCK_MAX) ? CJ_ST_CORRUPT : CJ_ST_DST_SMALL;
    if (rest) out.put_literals_coherent(lit + lp, rest);
    if (out.op - block_start > ZS_BLOCK_MAX) return CJ_ST_CORRUPT;
    return CJ_OK;
}

// XXH64 of a global range (one lane).
__device__ uint64_t xxh64_global(const uint8_t* p, uint64_t n) {
    const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull, P4 = 9650029242287828579ull,
                   P5 = 2870177450012600261ull;
    auto rotl = [](uint64_t x, int r) { return (x << r) | (x >> (64 - r)); };
    auto rd64 = [&](uint64_t q) { uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)__ldcg(p + q + i) << (8 * i); return v; };
    auto rd32 = [&](uint64_t q) { uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)__ldcg(p + q + i) << (8 * i); return v; };
    auto round = [&](uint64_t acc, uint64_t in) { return rotl(acc + in * P2, 31) * P1; };
    auto merge = [&](uint64_t h, uint64_t v) { return (h ^ round(0, v)) * P1 + P4; };
    uint64_t q = 0, h;
    const bool al8 = ((uintptr_t)p & 7) == 0;
    if (n >= 32) {
        uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0ull - P1;
        do {
            if (al8) {
                const uint64_t* w = reinterpret_cast<const uint64_t*>(p + q);
                v1 = round(v1, __ldcg(w)); v2 = round(v2, __ldcg(w + 1)); v3 = round(v3, __ldcg(w + 2)); v4 = round(v4, __ldcg(w + 3));
            } else {
                v1 = round(v1, rd64(q)); v2 = round(v2, rd64(q + 8)); v3 = round(v3, rd64(q + 16)); v4 = round(v4, rd64(q + 24));
            }
            q += 32;
        } while (q + 32 <= n);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        h = merge(h, v1); h = merge(h, v2); h = merge(h, v3); h = merge(h, v4);
    } else {
        h = P5;
    }
    h += n;
    while (q + 8 <= n) { h = rotl(h ^ round(0, rd64(q)), 27) * P1 + P4; q += 8; }
    if (q + 4 <= n) { h = rotl(h ^ ((uint64_t)rd32(q) * P1), 23) * P2 + P3; q += 4; }
    while (q < n) { h = rotl(h ^ ((uint64_t)__ldcg(p + q) * P5), 11) * P1; q++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

__device__ int32_t zstd_decode_stream(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap, uint8_t* smem_warp, uint8_t* lit_buf,
                                      int lane, uint32_t* produced) {
    OutRing out;
    out.init(smem_warp, dst, lane);
    ZState z;
    z.huf = reinterpret_cast<uint16_t*>(smem_warp + ZS_OFF_HUF);
    z.ll.t = reinterpret_cast<uint32_t*>(smem_warp + ZS_OFF_LL);
    z.of.t = reinterpret_cast<uint32_t*>(smem_warp + ZS_OFF_OF);
    z.ml.t = reinterpret_cast<uint32_t*>(smem_warp + ZS_OFF_ML);
    z.freq = reinterpret_cast<int16_t*>(smem_warp + ZS_OFF_TMP);
    z.next = reinterpret_cast<uint16_t*>(smem_warp + ZS_OFF_TMP + 512);
    z.w = smem_warp + ZS_OFF_TMP + 1024;
    z.wfse = reinterpret_cast<uint32_t*>(smem_warp + ZS_OFF_TMP + 1024 + 272);
    z.huf_w = smem_warp + ZS_OFF_SAVE;
    z.ll.saved = reinterpret_cast<int16_t*>(smem_warp + ZS_OFF_SAVE + 256);
    z.of.saved = z.ll.saved + 36;
    z.ml.saved = z.of.saved + 32;
    z.lit = lit_buf;
    uint32_t ip = 0;
    int32_t st = CJ_OK;
    while (ip < n && st == CJ_OK) {
        if (n - ip < 4) { st = CJ_ST_TRUNCATED; break; }
        const uint32_t magic = rd32g(src + ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (n - ip < 8) { st = CJ_ST_TRUNCATED; break; }
            const uint32_t sz = rd32g(src + ip + 4);
            if (sz > n - ip - 8) { st = CJ_ST_TRUNCATED; break; }
            ip += 8 + sz;
            continue;
        }
        // ---- frame header ----
        if (n - ip < 5) { st = CJ_ST_TRUNCATED; break; }
        if (magic != 0xFD2FB528u) { st = CJ_ST_HEADER; break; }
        const uint32_t fhd = __ldg(src + ip + 4);
        const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
        if (fhd & 0x08) { st = CJ_ST_HEADER; break; }
        const bool checksum = (fhd >> 2) & 1;
        uint32_t pos = ip + 5;
        uint64_t window = 0;
        if (!single) {
            if (n < pos + 1) { st = CJ_ST_TRUNCATED; break; }
            const uint32_t wd = __ldg(src + pos++);
            const int wl = 10 + (wd >> 3);
            if (wl > 31) { st = CJ_ST_UNSUPPORTED; break; }
            window = (1ull << wl) + ((1ull << wl) >> 3) * (wd & 7);
        }
        const uint32_t did_sz = did == 0 ? 0 : (did == 1 ? 1 : (did == 2 ? 2 : 4));
        if (n < pos + did_sz) { st = CJ_ST_TRUNCATED; break; }
        uint32_t dict = 0;
        for (uint32_t i = 0; i < did_sz; i++) dict |= __ldg(src + pos + i) << (8 * i);
        pos += did_sz;
        if (dict != 0) { st = CJ_ST_UNSUPPORTED; break; }
        const uint32_t fsz = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
        if (n < pos + fsz) { st = CJ_ST_TRUNCATED; break; }
        uint64_t fcs = 0;
        for (uint32_t i = 0; i < fsz; i++) fcs |= (uint64_t)__ldg(src + pos + i) << (8 * i);
        if (fsz == 2) fcs += 256;
        pos += fsz;
        if (single) window = fcs;
        ip = pos;
        const uint32_t bmax = window < ZS_BLOCK_MAX ? (uint32_t)window : ZS_BLOCK_MAX;
        const uint32_t frame_start = out.op;
        out.base = frame_start;
        z.huf_ok = z.huf_live = false;
        z.huf_nw = 0;
        z.ll.ok = z.of.ok = z.ml.ok = false;
        z.ll.live = z.of.live = z.ml.live = false;
        z.ll.kind = z.of.kind = z.ml.kind = 0;
        z.ll.sym = z.of.sym = z.ml.sym = 0;
        z.ll.nsym = z.of.nsym = z.ml.nsym = 0;
        z.ll.al = z.of.al = z.ml.al = 0;
        z.rep0 = 1; z.rep1 = 4; z.rep2 = 8;
        // ---- blocks ----
        for (;;) {
            if (n - ip < 3) { st = CJ_ST_TRUNCATED; break; }
            const uint32_t bh = __ldg(src + ip) | (__ldg(src + ip + 1) << 8) | (__ldg(src + ip + 2) << 16);
            ip += 3;
            const bool last = bh & 1;
            const uint32_t type = (bh >> 1) & 3, bsz = bh >> 3;
            if (type == 3) { st = CJ_ST_CORRUPT; break; }
            const uint32_t in_sz = type == 1 ? 1 : bsz;
            if (in_sz > n - ip) { st = CJ_ST_TRUNCATED; break; }
            if (type == 2 ? bsz > ZS_BLOCK_MAX : bsz > bmax) { st = CJ_ST_CORRUPT; break; }
            if (type == 0) {
                if (bsz > cap - out.op) { st = CJ_ST_DST_SMALL; break; }
                if (bsz) out.put_literals(src + ip, bsz);
            } else if (type == 1) {
                if (bsz > cap - out.op) { st = CJ_ST_DST_SMALL; break; }
                if (bsz) {  // one literal byte, then an offset-1 match
                    out.put_literals(src + ip, 1);
                    if (bsz > 1) out.put_match(1, bsz - 1);
                }
            } else {
                const uint32_t before = out.op;
                st = zs_block(z, src + ip, bsz, out, cap, lane);
                if (st != CJ_OK) break;
                if (out.op - before > bmax) { st = CJ_ST_CORRUPT; break; }
            }
            ip += in_sz;
            if (last) break;
        }
        if (st != CJ_OK) break;
        if (checksum) {
            if (n - ip < 4) { st = CJ_ST_TRUNCATED; break; }
            out.flush_to(out.op, true);
            uint32_t h = 0;
            if (lane == 0) h = (uint32_t)xxh64_global(dst + frame_start, out.op - frame_start);
            h = __shfl_sync(FULL, h, 0);
            if (h != rd32g(src + ip)) { st = CJ_ST_CHECKSUM; break; }
            ip += 4;
        }
        if (fsz != 0 && fcs != (uint64_t)(out.op - frame_start)) { st = CJ_ST_LEN_MISMATCH; break; }
    }
    out.finish();
    *produced = out.op;
    return st;
}

__global__ void __launch_bounds__(ZS_WARPS * 32) zstd_decode_kernel(Batch b, unsigned* __restrict__ counter, uint8_t* __restrict__ lit_scratch) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* smem_warp = smem + (size_t)warp * ZS_SMEM_WARP;
    uint8_t* lit_buf = lit_scratch + ((size_t)blockIdx.x * ZS_WARPS + warp) * ZS_LIT_STRIDE;
    for (;;) {
        const uint32_t u = next_unit(counter, lane);
        if (u >= b.n) break;
        const uint64_t slen = b.src_len[u], dcap = b.dst_cap[u];
        uint32_t produced = 0;
        int32_t st;
        if (slen > MAX_UNIT) st = CJ_ST_TOO_BIG;
        else st = zstd_decode_stream(b.src_base + b.src_off[u], (uint32_t)slen, b.dst_base + b.dst_off[u], dcap > MAX_UNIT ? MAX_UNIT : (uint32_t)dcap,
                                     smem_warp, lit_buf, lane, &produced);
        if (lane == 0) {
            b.dst_len[u] = st == CJ_OK ? produced : 0;
            b.status[u] = st;
        }
        __syncwarp();
    }
}

int zstd_grid(int sm_count, uint32_t n) {
    int grid = sm_count * CJ_ZS_CTAS;  // CTAs x warps x 11.1 KiB of shared memory per SM
    const int need = (int)((n + ZS_WARPS - 1) / ZS_WARPS);
    if (grid > need) grid = need;
    return grid < 1 ? 1 : grid;
}

size_t zstd_scratch_bytes(int sm_count, uint32_t n) { return (size_t)zstd_grid(sm_count, n) * ZS_WARPS * ZS_LIT_STRIDE; }

cudaError_t launch_zstd_decode(const Batch& b, unsigned* counter, uint8_t* lit_scratch, int sm_count, cudaStream_t stream) {
    const size_t smem = (size_t)ZS_SMEM_WARP * ZS_WARPS;
    static cj_per_device_flag attr_flag;
    int& attr_done = attr_flag.here();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(zstd_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = 1;
    }
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    zstd_decode_kernel<<<zstd_grid(sm_count, b.n), ZS_WARPS * 32, smem, stream>>>(b, counter, lit_scratch);
    return cudaGetLastError();
}

}  // namespace cj

// ---- host-side header walk: decompressed size / bound (no payload byte is interpreted) ----
int cj_zstd_walk_host(const uint8_t* s, size_t n, size_t* out, bool* exact, std::vector<cj_frame_info>* frames) {
    size_t p = 0, tot = 0;
    *exact = true;
    while (p < n) {
        if (n - p < 4) return CJ_ST_TRUNCATED;
        uint32_t magic; memcpy(&magic, s + p, 4);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (n - p < 8) return CJ_ST_TRUNCATED;
            uint32_t sz; memcpy(&sz, s + p + 4, 4);
            if (sz > n - p - 8) return CJ_ST_TRUNCATED;
            if (frames) frames->push_back({p, (size_t)8 + sz, 0, true});
            p += 8 + sz;
            continue;
        }
        if (magic != 0xFD2FB528u) return CJ_ST_HEADER;
        if (n - p < 5) return CJ_ST_TRUNCATED;
        const uint8_t fhd = s[p + 4];
        const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
        size_t q = p + 5;
        uint64_t window = 0;
        if (!single) {
            if (n < q + 1) return CJ_ST_TRUNCATED;
            const uint8_t wd = s[q++];
            const int wl = 10 + (wd >> 3);
            if (wl > 31) return CJ_ST_UNSUPPORTED;
            window = (1ull << wl) + ((1ull << wl) >> 3) * (wd & 7);
        }
        static const int did_sz[4] = {0, 1, 2, 4};
        q += did_sz[did];
        const int fsz = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
        if (n < q + fsz) return CJ_ST_TRUNCATED;
        uint64_t fcs = 0;
        for (int i = 0; i < fsz; i++) fcs |= (uint64_t)s[q + i] << (8 * i);
        if (fsz == 2) fcs += 256;
        q += fsz;
        if (single) window = fcs;
        const size_t bmax = window < 128 * 1024 ? (size_t)window : 128 * 1024;
        size_t frame_tot = 0;
        for (;;) {
            if (n - q < 3) return CJ_ST_TRUNCATED;
            const uint32_t bh = s[q] | ((uint32_t)s[q + 1] << 8) | ((uint32_t)s[q + 2] << 16);
            q += 3;
            const uint32_t type = (bh >> 1) & 3, bsz = bh >> 3;
            if (type == 3) return CJ_ST_CORRUPT;
            const size_t in_sz = type == 1 ? 1 : bsz;
            if (in_sz > n - q) return CJ_ST_TRUNCATED;
            if ((type == 2 ? bmax : (size_t)bsz) > SIZE_MAX / 2 - frame_tot) return CJ_ST_TOO_BIG;
            frame_tot += type == 2 ? bmax : bsz;
            q += in_sz;
            if (bh & 1) break;
        }
        if (fhd & 4) { if (n - q < 4) return CJ_ST_TRUNCATED; q += 4; }
        // untrusted 64-bit content sizes: a sum that wraps must not come out as a small bound
        const uint64_t add = fsz ? fcs : (uint64_t)frame_tot;
        if (add > (uint64_t)SIZE_MAX - tot) return CJ_ST_TOO_BIG;
        tot += (size_t)add;
        if (!fsz) *exact = false;
        if (frames) frames->push_back({p, q - p, (size_t)add, fsz != 0});
        p = q;
    }
    *out = tot;
    return CJ_OK;
}
