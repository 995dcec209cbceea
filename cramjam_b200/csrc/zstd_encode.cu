// zstd_encode.cu — Zstandard frame encoder (sm_100a): cramjam.zstd.compress / compress_into
// (reference src/zstd.rs:37-64 -> libcramjam::zstd::compress -> libzstd).
//
// SURVEY.md 8f ranks the zstd encoder as a "next" row (north_star names zstd *decode*); this is a first
// real encoder so the API produces compressed frames rather than stored ones: LZ77 parsing by the same
// warp-parallel greedy hash match finder as the LZ4 / Snappy encoders (lz_match.cuh), sequences coded
// with FSE over the PREDEFINED literal-length / offset / match-length distributions (RFC 8878 3.1.1.3.2.2),
// literals Huffman-coded per block (zstd_huf.cuh: histogram by the warp, code lengths + tree description by one lane,
// the four streams by four lanes) when that shrinks them, raw otherwise.  `level`: 1-2 (and negative levels) skip the
// Huffman stage (fastest), everything else — the reference default 0 -> 3 included — runs it.  The match finder is the
// greedy single-probe one of the block encoders, so the ratio stays below libzstd's at the same level.
// Frames are single-segment with the pledged content size
// (reference src/zstd.rs:45,61), blocks of <= 128 KiB; a block that does not shrink is emitted as a Raw_Block.
// One warp per frame; the FSE state chain is serial per block and runs warp-uniformly.
#include "internal.h"
#include "lz_match.cuh"
#include "zstd_huf.cuh"

namespace cj {

constexpr int ZE_WARPS = 11;   // two CTAs of 11 warps (22 warps x 9.5 KiB) fit an SM; five 4-warp CTAs hold 20 (1 KB of shared memory is reserved per CTA)
constexpr uint32_t ZE_BLOCK = 128 * 1024;
constexpr uint32_t ZE_MAXSEQ = ZE_BLOCK / 4 + 8;                       // every match is >= 4 bytes
constexpr size_t ZE_HUF_TMP = 256 + 4 * ((size_t)ZE_BLOCK / 4 * 3 / 2 + 64);          // tree description + four streams at their worst case (11 bits per symbol)
constexpr size_t ZE_SCRATCH = (size_t)ZE_MAXSEQ * 12 + ZE_BLOCK + 64 + ZE_HUF_TMP + 64;  // per warp: sequences (ll, ml, off) + literal bytes + Huffman scratch
constexpr int ZE_HUF_SMEM = 256 * 4 + 256 * 2;   // per warp: literal histogram + code table (value | nbits << 12)

struct ZEncTables {  // FSE compression tables of the predefined distributions + length->code maps
    uint16_t ll_state[64], of_state[32], ml_state[64];
    int32_t ll_dnb[36], ll_dfs[36], of_dnb[32], of_dfs[32], ml_dnb[53], ml_dfs[53];
    uint32_t ll_base[36], ml_base[53];
    uint8_t ll_bits[36], ml_bits[53];
    uint8_t ll_code[64], ml_code[128];
};
__constant__ ZEncTables g_ze;

struct BitW {  // forward, LSB-first bit writer into global memory
    uint8_t* p;
    uint32_t pos;  // bytes written
    uint64_t acc;
    uint32_t nb;
    int lane;
    __device__ __forceinline__ void add(uint32_t v, uint32_t n) {
        acc |= (uint64_t)(v & (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1))) << nb;
        nb += n;
    }
    // v has no bits at or above n (n <= 32, nb + n <= 64): the caller packed several small fields into it with 32-bit operations, so
    // the 64-bit shift is paid once per group instead of once per field
    __device__ __forceinline__ void add_packed(uint32_t v, uint32_t n) {
        const uint32_t s = nb & 31u;
        const uint32_t lo = v << s, hi = s ? v >> (32u - s) : 0u;
        acc |= nb < 32u ? ((uint64_t)hi << 32) | lo : (uint64_t)lo << 32;
        nb += n;
    }
    __device__ __forceinline__ void flush() {  // keeps < 8 bits pending; call at least every 56 added bits
        const uint32_t bytes = nb >> 3;
        if ((uint32_t)lane < bytes) p[pos + lane] = (uint8_t)(acc >> (8 * lane));
        pos += bytes;
        acc = bytes >= 8 ? 0 : acc >> (8 * bytes);
        nb &= 7;
    }
};

__device__ __forceinline__ uint32_t ze_hibit(uint32_t v) { return 31 - __clz(v); }
__device__ __forceinline__ uint32_t ze_ll_code(uint32_t ll) { return ll < 64 ? g_ze.ll_code[ll] : ze_hibit(ll) + 19; }
__device__ __forceinline__ uint32_t ze_ml_code(uint32_t mlb) { return mlb < 128 ? g_ze.ml_code[mlb] : ze_hibit(mlb) + 36; }

struct FseC {
    uint32_t state;
    __device__ __forceinline__ void init(const uint16_t* st, const int32_t* dnb, const int32_t* dfs, uint32_t sym) {
        const uint32_t nbo = (uint32_t)(dnb[sym] + (1 << 15)) >> 16;
        const uint32_t value = (nbo << 16) - (uint32_t)dnb[sym];
        state = st[(int32_t)(value >> nbo) + dfs[sym]];
    }
    __device__ __forceinline__ void encode(BitW& w, const uint16_t* st, const int32_t* dnb, const int32_t* dfs, uint32_t sym) {
        const uint32_t nbo = (state + (uint32_t)dnb[sym]) >> 16;
        w.add(state, nbo);
        state = st[(int32_t)(state >> nbo) + dfs[sym]];
    }
    // the same, but the state bits go behind the `n` bits already packed in `bits` instead of into the writer
    __device__ __forceinline__ void encode_into(uint32_t& bits, uint32_t& n, const uint16_t* st, const int32_t* dnb, const int32_t* dfs, uint32_t sym) {
        const uint32_t nbo = (state + (uint32_t)dnb[sym]) >> 16;   // <= 9
        bits |= (state & ((1u << nbo) - 1u)) << n;
        n += nbo;
        state = st[(int32_t)(state >> nbo) + dfs[sym]];
    }
};

// Collects the sequences and literal bytes of a block (find_matches emitter; one match at a time).
struct ZSeqEmitter {
    const uint8_t* __restrict__ src;
    uint32_t* seq;
    uint8_t* lits;
    uint32_t nseq, nlit;
    int lane;
    __device__ __forceinline__ bool window(const uint8_t*, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t) { return false; }
    __device__ __forceinline__ void serial(uint32_t lit_at, uint32_t ll, uint32_t off, uint32_t ml) {
        for (uint32_t i = lane; i < ll; i += 32) lits[nlit + i] = __ldg(src + lit_at + i);
        if (lane == 0) { seq[3 * nseq] = ll; seq[3 * nseq + 1] = ml; seq[3 * nseq + 2] = off; }
        nlit += ll;
        nseq++;
    }
};

// Encodes one block src[b0, b1) into out (room for at least (b1-b0) + 16 bytes).  Returns the block content size
// written at out, or 0 if the block should be stored raw.
// Huffman-coded literals section (Compressed_Literals_Block) of `nlit` literal bytes at out; returns its size or 0 when the
// literals should be stored raw instead (too few, alphabet beyond the direct tree description, or no gain).
__device__ uint32_t ze_huf_literals(const uint8_t* lits, uint32_t nlit, uint8_t* tmp, uint32_t* hist, uint16_t* code, uint8_t* out, int lane) {
    if (nlit < 256) return 0;
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < nlit; i += 32) atomicAdd(&hist[lits[i]], 1u);
    __syncwarp();
    // ---- code lengths, canonical codes and the tree description: one lane (serial by nature, ~100 symbols) ----
    uint32_t tree_len = 0, est = 0;
    if (lane == 0) {
        uint8_t nbits[256];
        const int mb = zh::build_lengths(hist, nbits);
        if (mb) {
            tree_len = (uint32_t)zh::write_tree_direct(nbits, mb, tmp);   // parked at the start of the scratch, copied below
            if (tree_len) {
                zh::assign_codes(nbits, mb, code);
                uint64_t bits = 0;
                for (int sym = 0; sym < 256; sym++) bits += (uint64_t)hist[sym] * nbits[sym];
                est = (uint32_t)(bits / 8) + 4 + 6 + tree_len;
            }
        }
    }
    tree_len = __shfl_sync(FULL, tree_len, 0);
    est = __shfl_sync(FULL, est, 0);
    if (!tree_len || est + (nlit >> 6) + 8 >= nlit) return 0;
    __syncwarp();
    // ---- the streams: four lanes, one stream each (a stream is a serial bit chain), into the scratch behind the tree ----
    const bool four = nlit >= 1024;
    const uint32_t seg = four ? (nlit + 3) / 4 : nlit;
    const uint32_t slot = seg + seg / 2 + 64;   // a stream's worst case: 11 bits per symbol (ZE_HUF_TMP holds four of them)
    uint8_t* streams = tmp + 256;
    uint32_t sz = 0;
    if (lane < (four ? 4 : 1)) {
        const uint32_t a = lane * seg, b = (four && lane < 3) ? a + seg : nlit;
        sz = zh::encode_stream(lits + a, b - a, code, streams + (size_t)lane * slot);
    }
    __syncwarp();
    const uint32_t s0 = __shfl_sync(FULL, sz, 0), s1 = __shfl_sync(FULL, sz, 1), s2 = __shfl_sync(FULL, sz, 2), s3 = __shfl_sync(FULL, sz, 3);
    const uint32_t comp = tree_len + (four ? 6 + s0 + s1 + s2 + s3 : s0);
    if ((nlit < 1024 && comp >= 1024) || (nlit < 16384 && comp >= 16384) || comp + 8 >= nlit || (four && (s0 | s1 | s2) > 0xFFFFu)) return 0;
    const uint32_t hl = (uint32_t)zh::header_len(nlit);
    if (lane == 0) {
        zh::write_header(out, nlit, comp, four);
        if (four) {
            uint8_t* j = out + hl + tree_len;
            j[0] = (uint8_t)s0; j[1] = (uint8_t)(s0 >> 8); j[2] = (uint8_t)s1; j[3] = (uint8_t)(s1 >> 8); j[4] = (uint8_t)s2; j[5] = (uint8_t)(s2 >> 8);
        }
    }
    for (uint32_t i = lane; i < tree_len; i += 32) out[hl + i] = tmp[i];
    uint32_t op = hl + tree_len + (four ? 6 : 0);
    const uint32_t sizes[4] = {s0, s1, s2, s3};
    for (int k = 0; k < (four ? 4 : 1); k++) {
        const uint8_t* sp = streams + (size_t)k * slot;
        for (uint32_t i = lane; i < sizes[k]; i += 32) out[op + i] = sp[i];
        op += sizes[k];
    }
    __syncwarp();
    return op;
}

__device__ uint32_t ze_block(const uint8_t* __restrict__ src, uint32_t b0, uint32_t b1, enc_slot_t* table, uint32_t* seq, uint8_t* lits, uint8_t* out,
                             uint32_t* hist, uint16_t* code, bool huffman, int lane) {
    const uint32_t bsz = b1 - b0;
    uint32_t nseq = 0, nlit = 0;
    uint32_t anchor = b0;
    if (bsz >= 8) {
        ZSeqEmitter em{src, seq, lits, nseq, nlit, lane};
        anchor = find_matches(src, b0, b1 - 3, b1, table, lane, em);
        nseq = em.nseq;
        nlit = em.nlit;
    }
    if (nseq == 0) return 0;
    const uint32_t tail = b1 - anchor;
    for (uint32_t i = lane; i < tail; i += 32) lits[nlit + i] = __ldg(src + anchor + i);
    nlit += tail;
    __syncwarp();
    // ---- literals section: Compressed_Literals_Block (Huffman) when it pays, else Raw_Literals_Block ----
    uint32_t op = huffman ? ze_huf_literals(lits, nlit, lits + ZE_BLOCK + 64, hist, code, out, lane) : 0;
    if (op == 0) {
        if (nlit < 32) { if (lane == 0) out[0] = (uint8_t)(nlit << 3); op = 1; }
        else if (nlit < 4096) { if (lane == 0) { out[0] = (uint8_t)((nlit << 4) | (1 << 2)); out[1] = (uint8_t)(nlit >> 4); } op = 2; }
        else { if (lane == 0) { out[0] = (uint8_t)((nlit << 4) | (3 << 2)); out[1] = (uint8_t)(nlit >> 4); out[2] = (uint8_t)(nlit >> 12); } op = 3; }
        if (op + nlit + 16 > bsz) return 0;
        for (uint32_t i = lane; i < nlit; i += 32) out[op + i] = lits[i];
        op += nlit;
    }
    if (op + 16 > bsz) return 0;
    // ---- sequences section: count, modes (all predefined), FSE bitstream written from the last sequence to the first ----
    if (nseq < 128) { if (lane == 0) out[op] = (uint8_t)nseq; op += 1; }
    else if (nseq < 0x7F00) { if (lane == 0) { out[op] = (uint8_t)((nseq >> 8) + 128); out[op + 1] = (uint8_t)nseq; } op += 2; }
    else { if (lane == 0) { out[op] = 255; out[op + 1] = (uint8_t)(nseq - 0x7F00); out[op + 2] = (uint8_t)((nseq - 0x7F00) >> 8); } op += 3; }
    if (lane == 0) out[op] = 0;
    op += 1;
    BitW w{out + op, 0, 0, 0, lane};
    FseC sl, so, sm;
    {
        const uint32_t ll = seq[3 * (nseq - 1)], mlb = seq[3 * (nseq - 1) + 1] - 3, ofv = seq[3 * (nseq - 1) + 2] + 3;
        const uint32_t lc = ze_ll_code(ll), mc = ze_ml_code(mlb), oc = ze_hibit(ofv);
        sm.init(g_ze.ml_state, g_ze.ml_dnb, g_ze.ml_dfs, mc);
        so.init(g_ze.of_state, g_ze.of_dnb, g_ze.of_dfs, oc);
        sl.init(g_ze.ll_state, g_ze.ll_dnb, g_ze.ll_dfs, lc);
        w.add(ll, g_ze.ll_bits[lc]);
        w.add(mlb, g_ze.ml_bits[mc]);
        w.flush();
        w.add(ofv, oc);
        w.flush();
    }
    // sequences are coded last to first; 32 of them are fetched at a time (lane j holds sequence top - j) and handed round by shuffle,
    // so the serial loop does not wait for global memory
    for (int64_t top = (int64_t)nseq - 2; top >= 0; top -= 32) {
      const uint32_t cntb = top + 1 < 32 ? (uint32_t)(top + 1) : 32u;
      uint32_t my_ll = 0, my_ml = 3, my_of = 0;
      if ((uint32_t)lane < cntb) {
          const uint32_t idx = (uint32_t)top - (uint32_t)lane;
          my_ll = seq[3 * idx]; my_ml = seq[3 * idx + 1]; my_of = seq[3 * idx + 2];
      }
      for (uint32_t j = 0; j < cntb; j++) {
        const uint32_t ll = __shfl_sync(FULL, my_ll, j), mlb = __shfl_sync(FULL, my_ml, j) - 3, ofv = __shfl_sync(FULL, my_of, j) + 3;
        const uint32_t lc = ze_ll_code(ll), mc = ze_ml_code(mlb), oc = ze_hibit(ofv);
        // the three state updates (<= 8 + 9 + 9 bits) are packed with 32-bit operations and added once; so are the two length
        // fields (<= 16 + 16); then the offset (<= 17): at most 7 + 26 + 32 = 65 bits could be pending, so the writer is flushed
        // after the states only when the lengths would not fit (rare) and once at the end of the sequence
        uint32_t sb = 0, sn = 0;
        so.encode_into(sb, sn, g_ze.of_state, g_ze.of_dnb, g_ze.of_dfs, oc);
        sm.encode_into(sb, sn, g_ze.ml_state, g_ze.ml_dnb, g_ze.ml_dfs, mc);
        sl.encode_into(sb, sn, g_ze.ll_state, g_ze.ll_dnb, g_ze.ll_dfs, lc);
        w.add_packed(sb, sn);
        const uint32_t lbits = g_ze.ll_bits[lc], mbits = g_ze.ml_bits[mc];
        if (w.nb + lbits + mbits + oc > 64u) w.flush();               // afterwards < 8 pending: 7 + 32 + 17 <= 56
        const uint32_t lm = (ll & ((1u << lbits) - 1u)) | ((lbits < 32u ? (mlb & ((1u << mbits) - 1u)) << lbits : 0u));
        if (lbits + mbits <= 32u) w.add_packed(lm, lbits + mbits);
        else { w.add(ll, lbits); w.add(mlb, mbits); }
        w.add(ofv, oc);                                               // <= 17
        w.flush();
        if (op + w.pos + 16 > bsz) return 0;  // not shrinking: store the block raw instead
      }
    }
    w.add(sm.state, 6);
    w.add(so.state, 5);
    w.add(sl.state, 6);
    w.flush();
    w.add(1, 1);  // end mark
    w.nb = (w.nb + 7) & ~7u;
    w.flush();
    __syncwarp();
    op += w.pos;
    return op + 3 < bsz ? op : 0;
}

__device__ int32_t ze_frame(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint64_t cap, enc_slot_t* table, uint8_t* scratch, uint32_t* hist,
                            uint16_t* code, bool huffman, int lane, uint32_t* produced) {
    *produced = 0;
    const uint64_t bound = (uint64_t)n + 3ull * (n / ZE_BLOCK + 1) + 18;
    if (cap < bound) return CJ_ST_DST_SMALL;
    uint32_t* seq = reinterpret_cast<uint32_t*>(scratch);
    uint8_t* lits = scratch + (size_t)ZE_MAXSEQ * 12;
    uint32_t op = 0;
    // frame header: magic, single-segment descriptor, Frame_Content_Size
    uint32_t hl;
    if (lane == 0) {
        dst[0] = 0x28; dst[1] = 0xB5; dst[2] = 0x2F; dst[3] = 0xFD;
        if (n < 256) { dst[4] = 0x20; dst[5] = (uint8_t)n; }
        else if (n < 65536 + 256) { dst[4] = 0x60; const uint32_t v = n - 256; dst[5] = (uint8_t)v; dst[6] = (uint8_t)(v >> 8); }
        else { dst[4] = 0xA0; dst[5] = (uint8_t)n; dst[6] = (uint8_t)(n >> 8); dst[7] = (uint8_t)(n >> 16); dst[8] = (uint8_t)(n >> 24); }
    }
    hl = n < 256 ? 6 : (n < 65536 + 256 ? 7 : 9);
    op = hl;
    match_table_reset(table, lane);
    uint32_t b0 = 0;
    do {
        const uint32_t b1 = min(n, b0 + ZE_BLOCK);
        const bool last = b1 >= n;
        const uint32_t bsz = b1 - b0;
        uint32_t csz = bsz >= 64 ? ze_block(src, b0, b1, table, seq, lits, dst + op + 3, hist, code, huffman, lane) : 0;
        __syncwarp();
        uint32_t bh;
        if (csz) {
            bh = (last ? 1u : 0u) | (2u << 1) | (csz << 3);
        } else {  // Raw_Block
            for (uint32_t i = lane; i < bsz; i += 32) dst[op + 3 + i] = __ldg(src + b0 + i);
            csz = bsz;
            bh = (last ? 1u : 0u) | (bsz << 3);
        }
        if (lane == 0) { dst[op] = (uint8_t)bh; dst[op + 1] = (uint8_t)(bh >> 8); dst[op + 2] = (uint8_t)(bh >> 16); }
        op += 3 + csz;
        b0 = b1;
    } while (b0 < n);
    __syncwarp();
    *produced = op;
    return CJ_OK;
}

__global__ void __launch_bounds__(ZE_WARPS * 32) zstd_encode_kernel(Batch b, unsigned* __restrict__ counter, uint8_t* __restrict__ scratch, int huffman) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    enc_slot_t* table = reinterpret_cast<enc_slot_t*>(smem) + (size_t)warp * ENC_HSIZE;
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + ENC_TABLE_BYTES * ZE_WARPS + (size_t)warp * ZE_HUF_SMEM);
    uint16_t* code = reinterpret_cast<uint16_t*>(hist + 256);
    uint8_t* my_scratch = scratch + ((size_t)blockIdx.x * ZE_WARPS + warp) * ZE_SCRATCH;
    for (;;) {
        const uint32_t u = next_unit(counter, lane);
        if (u >= b.n) break;
        const uint64_t slen = b.src_len[u];
        uint32_t produced = 0;
        int32_t st;
        if (slen > MAX_UNIT) st = CJ_ST_TOO_BIG;
        else st = ze_frame(b.src_base + b.src_off[u], (uint32_t)slen, b.dst_base + b.dst_off[u], b.dst_cap[u], table, my_scratch, hist, code, huffman != 0, lane, &produced);
        __syncwarp();
        if (lane == 0) {
            b.dst_len[u] = st == CJ_OK ? produced : 0;
            b.status[u] = st;
        }
    }
}

// ---- host: FSE compression tables of the predefined distributions (built once per process) ----
namespace {
const int16_t LL_DEF[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
const int16_t ML_DEF[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
const int16_t OF_DEF[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
const uint32_t LL_BASE_H[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
const uint8_t LL_BITS_H[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
const uint32_t ML_BASE_H[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
const uint8_t ML_BITS_H[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

int hibit_h(uint32_t v) { int r = 0; while (v >>= 1) r++; return r; }

void build_ctable(const int16_t* norm, int nsym, int tlog, uint16_t* state_tab, int32_t* dnb, int32_t* dfs) {
    const int size = 1 << tlog, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    std::vector<int> cumul(nsym + 2, 0);
    std::vector<uint8_t> sym(size, 0);
    int high = size - 1;
    for (int u = 1; u <= nsym; u++) {
        if (norm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1; sym[high--] = (uint8_t)(u - 1); }
        else cumul[u] = cumul[u - 1] + norm[u - 1];
    }
    int pos = 0;
    for (int s = 0; s < nsym; s++)
        for (int i = 0; i < norm[s]; i++) {
            sym[pos] = (uint8_t)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    for (int u = 0; u < size; u++) { const int s = sym[u]; state_tab[cumul[s]++] = (uint16_t)(size + u); }
    int total = 0;
    for (int s = 0; s < nsym; s++) {
        if (norm[s] == 0) { dnb[s] = ((tlog + 1) << 16) - (1 << tlog); dfs[s] = 0; }
        else if (norm[s] == -1 || norm[s] == 1) { dnb[s] = (tlog << 16) - (1 << tlog); dfs[s] = total - 1; total++; }
        else {
            const int max_bits_out = tlog - hibit_h((uint32_t)norm[s] - 1);
            const int min_state_plus = norm[s] << max_bits_out;
            dnb[s] = (max_bits_out << 16) - min_state_plus;
            dfs[s] = total - norm[s];
            total += norm[s];
        }
    }
}
}  // namespace

static cudaError_t upload_tables() {
    ZEncTables t;
    memset(&t, 0, sizeof t);
    build_ctable(LL_DEF, 36, 6, t.ll_state, t.ll_dnb, t.ll_dfs);
    build_ctable(OF_DEF, 29, 5, t.of_state, t.of_dnb, t.of_dfs);
    build_ctable(ML_DEF, 53, 6, t.ml_state, t.ml_dnb, t.ml_dfs);
    // offset codes 29..31 are not in the predefined table; offsets here never exceed 65535 + 3 (code <= 16)
    memcpy(t.ll_base, LL_BASE_H, sizeof t.ll_base);
    memcpy(t.ml_base, ML_BASE_H, sizeof t.ml_base);
    memcpy(t.ll_bits, LL_BITS_H, sizeof t.ll_bits);
    memcpy(t.ml_bits, ML_BITS_H, sizeof t.ml_bits);
    for (uint32_t ll = 0; ll < 64; ll++) { int c = 35; while (LL_BASE_H[c] > ll) c--; t.ll_code[ll] = (uint8_t)c; }
    for (uint32_t mlb = 0; mlb < 128; mlb++) { int c = 52; while (ML_BASE_H[c] > mlb + 3) c--; t.ml_code[mlb] = (uint8_t)c; }
    return cudaMemcpyToSymbol(g_ze, &t, sizeof t);
}

int zstd_enc_grid(int sm_count, uint32_t n) {
    int grid = sm_count * 2;  // 9.5 KiB of shared memory per warp (match table + Huffman histogram / codes): 22 warps per SM
    const int need = (int)((n + ZE_WARPS - 1) / ZE_WARPS);
    if (grid > need) grid = need;
    return grid < 1 ? 1 : grid;
}

size_t zstd_enc_scratch_bytes(int sm_count, uint32_t n) { return (size_t)zstd_enc_grid(sm_count, n) * ZE_WARPS * ZE_SCRATCH; }

cudaError_t launch_zstd_encode(const Batch& b, unsigned* counter, uint8_t* scratch, int sm_count, int level, cudaStream_t stream) {
    const size_t smem = (ENC_TABLE_BYTES + ZE_HUF_SMEM) * ZE_WARPS;
    // libzstd's levels trade speed for ratio; here: levels 1-2 and the negative ("fast") levels store literals raw, every
    // other level (0 = the library default 3 included) adds the Huffman literal stage
    const int huffman = (level == 1 || level == 2 || level < 0) ? 0 : 1;
    static cj_per_device_flag ready_flag;   // constant tables and the attribute belong to the device the call is made on
    int& ready = ready_flag.here();
    if (!ready) {
        cudaError_t e = upload_tables();
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(zstd_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ready = 1;
    }
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    zstd_encode_kernel<<<zstd_enc_grid(sm_count, b.n), ZE_WARPS * 32, smem, stream>>>(b, counter, scratch, huffman);
    return cudaGetLastError();
}

}  // namespace cj
