// lz_encode.cu — LZ4 block and Snappy raw block encode kernels (sm_100a).
//
// Replaces, for a whole batch of independent blocks per launch:
//   LZ4   : lz4::block::compress_into -> LZ4_compress_default / _fast     (reference src/lz4.rs:113-131,191-216)
//   Snappy: snap::raw::Encoder::compress                                  (reference src/snappy.rs:73-78,94-99)
// The reference pins compressed bytes only for one 14-byte input (tests/test_variants.py:329-334,
// which is all-literal and reproduced here); everything else is "format-valid + round-trip exact",
// so the match finder is designed for the GPU, not transliterated.
//
// Execution model: persistent grid, one warp per block, atomic work queue.  Per warp a 4096-entry
// hash table of 16-bit position slots lives in shared memory (8 KiB; lz_match.cuh).  The warp hashes 32 consecutive positions
// per step (one per lane), looks all of them up, then inserts them — when several lanes share a
// hash bucket the highest position wins, chosen with __match_any_sync so the result never depends
// on store ordering (the encoder is deterministic; tests/test_variants.py:281 needs
// compress_block == compress_block_into).  Every lane verifies its candidate with one 4-byte
// compare; the matches of a step are then taken greedily in position order: the warp extends each
// one 32 bytes per ballot, emits the pending literal run (lane-parallel byte copy) and the copy
// element(s), and skips the lanes the match covered.
#include "lz_match.cuh"

namespace cj {

#ifndef CJ_ENC_WARPS
#define CJ_ENC_WARPS 14   // warps per CTA: 2 CTAs x 14 warps x 8 KiB tables fill an SM (28 warps); 4-warp CTAs stop at 6 CTAs = 24 warps (1 KB of shared memory is reserved per CTA)
#endif
constexpr int ENC_WARPS = CJ_ENC_WARPS;

struct EncOut {
    uint8_t* dst;
    uint32_t op;
    int lane;
    __device__ __forceinline__ void bytes_from(const uint8_t* __restrict__ s, uint32_t len) {
        for (uint32_t i = lane; i < len; i += 32) dst[op + i] = __ldg(s + i);
        op += len;
    }
    // up to 4 header bytes packed little-endian in v
    __device__ __forceinline__ void put_packed(uint32_t v, uint32_t nbytes) {
        if ((uint32_t)lane < nbytes) dst[op + lane] = (uint8_t)(v >> (8 * lane));
        op += nbytes;
    }
    __device__ __forceinline__ void fill(uint8_t v, uint32_t count) {
        for (uint32_t i = lane; i < count; i += 32) dst[op + i] = v;
        op += count;
    }
};

// ---- Snappy element emission -----------------------------------------------------------------
__device__ __forceinline__ void snappy_emit_literal(EncOut& o, const uint8_t* __restrict__ s, uint32_t len) {
    if (len == 0) return;
    const uint32_t n1 = len - 1;
    if (n1 < 60) o.put_packed(n1 << 2, 1);
    else if (n1 < 256) o.put_packed((60u << 2) | (n1 << 8), 2);
    else if (n1 < 65536) o.put_packed((61u << 2) | (n1 << 8), 3);
    else if (n1 < (1u << 24)) o.put_packed((62u << 2) | (n1 << 8), 4);
    else { o.put_packed(63u << 2, 1); o.put_packed(n1, 4); }
    o.bytes_from(s, len);
}

__device__ __forceinline__ void snappy_emit_copy_upto64(EncOut& o, uint32_t off, uint32_t len) {
    if (len < 12 && off < 2048) o.put_packed(1u | ((len - 4) << 2) | ((off >> 8) << 5) | ((off & 0xff) << 8), 2);
    else o.put_packed(2u | ((len - 1) << 2) | (off << 8), 3);
}

__device__ __forceinline__ void snappy_emit_copy(EncOut& o, uint32_t off, uint32_t len) {
    if (len >= 68) {  // many 64-byte copy-2 elements: lanes write them side by side
        const uint32_t cnt = (len - 4) / 64;  // keeps a tail of >= 4 bytes
        const uint32_t packed = 2u | (63u << 2) | (off << 8);
        for (uint32_t i = o.lane; i < cnt * 3; i += 32) o.dst[o.op + i] = (uint8_t)(packed >> (8 * (i % 3)));
        o.op += cnt * 3;
        len -= cnt * 64;
    }
    if (len > 64) { snappy_emit_copy_upto64(o, off, 60); len -= 60; }
    snappy_emit_copy_upto64(o, off, len);
}

// ---- LZ4 sequence emission --------------------------------------------------------------------
// token + literal-length extension + literals (+ offset + match-length extension when mlen != 0)
__device__ __forceinline__ void lz4_emit_sequence(EncOut& o, const uint8_t* __restrict__ lit, uint32_t ll, uint32_t off, uint32_t mlen) {
    const uint32_t mcode = mlen ? mlen - 4 : 0;
    const uint32_t token = (min(ll, 15u) << 4) | min(mcode, 15u);
    o.put_packed(token, 1);
    if (ll >= 15) {
        const uint32_t r = ll - 15;
        o.fill(255, r / 255);
        o.put_packed(r % 255, 1);
    }
    o.bytes_from(lit, ll);
    if (mlen) {
        o.put_packed(off, 2);
        if (mcode >= 15) {
            const uint32_t r = mcode - 15;
            o.fill(255, r / 255);
            o.put_packed(r % 255, 1);
        }
    }
}

// ---- emitter handed to find_matches -------------------------------------------------------------
// window(): all matches of one 32-position step at once.  Every chosen lane sizes its own element pair
// (literal run + copy / LZ4 sequence), a warp prefix sum places the pairs, the lane writes its own header
// bytes, and every literal position of the step stores its byte (already in the lane's register) where its
// run lands — no per-match serial work, no reload of literal bytes.  Steps holding anything that needs more
// than one header byte per length (or several copy elements) fall back to serial(), one match at a time.
template <int CODEC>
struct BlockEmitter {
    EncOut& o;
    const uint8_t* __restrict__ src;

    __device__ __forceinline__ void serial(uint32_t lit_at, uint32_t ll, uint32_t off, uint32_t ml) {
        if (CODEC == CJ_SNAPPY_RAW) {
            snappy_emit_literal(o, src + lit_at, ll);
            snappy_emit_copy(o, off, ml);
        } else {
            lz4_emit_sequence(o, src + lit_at, ll, off, ml);
        }
    }

    __device__ __forceinline__ bool window(const uint8_t* __restrict__, uint32_t p, uint32_t v, uint32_t anchor, uint32_t sel, uint32_t mlen,
                                           uint32_t off) {
        const int lane = o.lane;
        const bool me = (sel >> lane) & 1;
        const uint32_t below = sel & ((1u << lane) - 1);
        const uint32_t pe = __shfl_sync(FULL, p + lane + mlen, below ? 31 - __clz(below) : 0);
        const uint32_t prev_end = below ? pe : anchor;  // where this lane's literal run starts
        const uint32_t ll = me ? p + lane - prev_end : 0u;
        uint32_t lhdr, tail;  // bytes before / after the literals
        bool simple;
        if (CODEC == CJ_SNAPPY_RAW) {
            simple = ll <= 60 && mlen <= 64;
            lhdr = ll ? 1 : 0;
            tail = (mlen < 12 && off < 2048) ? 2 : 3;
        } else {
            simple = ll < 15 + 255 && mlen - 4 < 15 + 255;
            lhdr = ll >= 15 ? 2 : 1;
            tail = mlen - 4 >= 15 ? 3 : 2;
        }
        if (__any_sync(FULL, me && !simple)) return false;
        const uint32_t size = me ? lhdr + ll + tail : 0u;
        uint32_t incl = size;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        uint8_t* dst = o.dst;
        const uint32_t ob = o.op + incl - size;  // where this lane's pair starts
        const uint32_t lb = ob + lhdr;           // ... and its literals
        if (me) {
            uint8_t* c = dst + lb + ll;
            if (CODEC == CJ_SNAPPY_RAW) {
                if (ll) dst[ob] = (uint8_t)((ll - 1) << 2);
                if (tail == 2) {
                    c[0] = (uint8_t)(1u | ((mlen - 4) << 2) | ((off >> 8) << 5));
                    c[1] = (uint8_t)off;
                } else {
                    c[0] = (uint8_t)(2u | ((mlen - 1) << 2));
                    c[1] = (uint8_t)off;
                    c[2] = (uint8_t)(off >> 8);
                }
            } else {
                const uint32_t mc = mlen - 4;
                dst[ob] = (uint8_t)((min(ll, 15u) << 4) | min(mc, 15u));
                if (ll >= 15) dst[ob + 1] = (uint8_t)(ll - 15);
                c[0] = (uint8_t)off;
                c[1] = (uint8_t)(off >> 8);
                if (mc >= 15) c[2] = (uint8_t)(mc - 15);
            }
        }
        // literal positions of this step: not inside a match, and a chosen match follows
        const uint32_t above = sel >> lane;
        const int k = above ? lane + __ffs(above) - 1 : 0;
        const uint32_t k_lb = __shfl_sync(FULL, lb, k), k_pe = __shfl_sync(FULL, prev_end, k);
        const uint32_t pos = p + lane;
        if (above && !me && pos >= k_pe) dst[k_lb + (pos - k_pe)] = (uint8_t)v;
        if (anchor < p) {  // the first run began in an earlier step: those bytes come from memory
            const uint32_t f_lb = __shfl_sync(FULL, lb, __ffs(sel) - 1);
            for (uint32_t i = lane; i < p - anchor; i += 32) dst[f_lb + i] = __ldg(src + anchor + i);
        }
        o.op += total;
        return true;
    }
};

template <int CODEC, int HBITS, int WAYS>
__device__ int32_t encode_block(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint64_t cap, enc_slot_t* table, int lane, uint32_t* produced) {
    *produced = 0;
    const uint64_t bound = CODEC == CJ_SNAPPY_RAW ? 32ull + n + n / 6 : (uint64_t)n + n / 255 + 16;
    if (CODEC == CJ_LZ4_BLOCK && n > 0x7E000000u) return CJ_ST_TOO_BIG;
    if (cap < bound) return CJ_ST_DST_SMALL;
    EncOut o;
    o.dst = dst;
    o.op = 0;
    o.lane = lane;
    if (CODEC == CJ_SNAPPY_RAW) {  // uvarint32 preamble
        uint32_t v = n, k = 0, packed = 0;
        uint8_t b4 = 0;
        for (;;) {
            uint32_t b = v & 0x7f;
            v >>= 7;
            if (v) b |= 0x80;
            if (k < 4) packed |= b << (8 * k); else b4 = (uint8_t)b;
            k++;
            if (!v) break;
        }
        o.put_packed(packed, min(k, 4u));
        if (k == 5) o.put_packed(b4, 1);
    }
    // positions that may start a match, and where a match must end
    const uint32_t start_limit = CODEC == CJ_SNAPPY_RAW ? (n >= 4 ? n - 4 + 1 : 0) : (n >= 13 ? n - 12 + 1 : 0);  // exclusive
    const uint32_t match_limit = CODEC == CJ_SNAPPY_RAW ? n : (n >= 5 ? n - 5 : 0);
    uint32_t anchor = 0;
    if (start_limit > 0) {
        match_table_reset<HBITS + WAYS - 1>(table, lane);
        BlockEmitter<CODEC> em{o, src};
        anchor = find_matches<BlockEmitter<CODEC>, HBITS, WAYS>(src, 0, start_limit, match_limit, table, lane, em);
    }
    if (CODEC == CJ_SNAPPY_RAW) snappy_emit_literal(o, src + anchor, n - anchor);
    else lz4_emit_sequence(o, src + anchor, n - anchor, 0, 0);
    *produced = o.op;
    return CJ_OK;
}

template <int CODEC, int HBITS, int WAYS>
__global__ void __launch_bounds__(ENC_WARPS * 32) lz_encode_kernel(Batch b, unsigned* __restrict__ counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    enc_slot_t* table = reinterpret_cast<enc_slot_t*>(smem) + (((size_t)warp * WAYS) << HBITS);
    for (;;) {
        const uint32_t u = next_unit(counter, lane);
        if (u >= b.n) break;
        const uint64_t slen = b.src_len[u];
        uint32_t produced = 0;
        int32_t st;
        if (slen > MAX_UNIT) st = CJ_ST_TOO_BIG;
        else st = encode_block<CODEC, HBITS, WAYS>(b.src_base + b.src_off[u], (uint32_t)slen, b.dst_base + b.dst_off[u], b.dst_cap[u], table, lane, &produced);
        __syncwarp();
        if (lane == 0) {
            b.dst_len[u] = st == CJ_OK ? produced : 0;
            b.status[u] = st;
        }
    }
}

template <int CODEC, int HBITS, int WAYS = 1>
static cudaError_t launch_enc(const Batch& b, unsigned* counter, int sm_count, cudaStream_t stream, bool reset_counter) {
    const size_t smem = (sizeof(enc_slot_t) << HBITS) * WAYS * ENC_WARPS;
    auto k = lz_encode_kernel<CODEC, HBITS, WAYS>;
    static cj_per_device_flag ctas_flag;
    int& ctas_per_sm = ctas_flag.here();
    if (!ctas_per_sm) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, ENC_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        ctas_per_sm = occ < 1 ? 1 : occ;
        if (const char* v = getenv("CJ_ENC_CTAS")) { const int f = atoi(v); if (f >= 1 && f <= occ) ctas_per_sm = f; }   // experiments: resident CTAs per SM
    }
    int grid = sm_count * ctas_per_sm;  // persistent: every resident CTA slot, units handed out by the atomic queue
    const int need = (int)((b.n + ENC_WARPS - 1) / ENC_WARPS);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (reset_counter) {  // the pinned-arena pipeline zeroes all of its counters once, up front, to keep copy engines off this stream
        cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned), stream);
        if (e != cudaSuccess) return e;
    }
    k<<<grid, ENC_WARPS * 32, smem, stream>>>(b, counter);
    return cudaGetLastError();
}

// effort: 0..2 = hash table of 2^10 .. 2^12 slots per warp (2 .. 8 KiB of shared memory); a smaller table leaves room for more
// resident warps and is faster, a larger one finds more matches.  3 = the HC-class search: 2^12 buckets of two positions each
// (a hash chain of depth two, 16 KiB), both candidates verified, and a lazy choice between neighbouring positions.  Measured on the bench corpus (Snappy, 16 384 x 64 KiB):
// 2^10 92 GB/s ratio 1.64 | 2^11 85 GB/s 1.79 | 2^12 58 GB/s 1.92 (default).  lz4 `acceleration` (src/lz4.rs:113-131) lowers
// the effort, HC-class levels (`compression=Some(n)`, lz4 frame level >= 3; src/lz4.rs:17,42-59) raise it.
template <int CODEC>
static cudaError_t launch_enc_effort(int effort, const Batch& b, unsigned* counter, int sm_count, cudaStream_t stream, bool reset_counter) {
    switch (effort) {
    case 0: return launch_enc<CODEC, 10>(b, counter, sm_count, stream, reset_counter);
    case 1: return launch_enc<CODEC, 11>(b, counter, sm_count, stream, reset_counter);
    case 3: return launch_enc<CODEC, 12, 2>(b, counter, sm_count, stream, reset_counter);   // HC class: two candidates per position + lazy choice
    default: return launch_enc<CODEC, 12>(b, counter, sm_count, stream, reset_counter);
    }
}

cudaError_t launch_lz_encode(int codec, const Batch& b, unsigned* counter, int sm_count, int effort, cudaStream_t stream, bool reset_counter) {
    return codec == CJ_LZ4_BLOCK ? launch_enc_effort<CJ_LZ4_BLOCK>(effort, b, counter, sm_count, stream, reset_counter)
                                 : launch_enc_effort<CJ_SNAPPY_RAW>(effort, b, counter, sm_count, stream, reset_counter);
}

}  // namespace cj
