// lz_decode.cu — LZ4 block and Snappy raw block decode kernels (sm_100a), generation 1.
//
// Replaces, for a whole batch of independent blocks per launch:
//   LZ4   : lz4::block::decompress_into -> LZ4_decompress_safe      (reference src/lz4.rs:78-95,140-173)
//   Snappy: snap::raw::Decoder::decompress                          (reference src/snappy.rs:55-60,103-108)
//
// Execution model: persistent grid, one warp per block, blocks handed out by a grid-wide atomic
// work queue (per-block work varies ~5x with the compression ratio).  Each warp owns a
// shared-memory output ring (RING bytes): every produced byte is written to the ring first, so
// back-references up to RING-CHUNK bytes away are served from shared memory, and the ring is
// drained to HBM with 16-byte vector stores (512 B per warp instruction, fully coalesced).  Rare
// far back-references re-read drained output through L2 (ld.global.cg).  Token / tag parsing is
// warp-uniform; literal and match bytes are moved one byte per lane, 32 bytes per instruction;
// overlapping matches use src = base + (i mod offset), which only touches already written bytes
// and therefore needs no intra-copy ordering.
//
// Acceptance rules and status codes are identical to oracle/lz4.c and oracle/snappy.c.
#include "common.cuh"

namespace cj {

template <int RING>
struct OutRing {
    static constexpr uint32_t MASK = RING - 1;
    static constexpr uint32_t CHUNK = 1024;    // largest span moved between room checks
    static constexpr uint32_t FLUSH_T = 2048;  // drain when this many bytes are pending
    static_assert(RING >= 2 * CHUNK + FLUSH_T + 16, "ring too small for the far-match invariant");

    uint8_t* ring;
    uint8_t* dst;
    uint32_t a;        // dst misalignment (dst & 15); ring index = (pos + a) & MASK
    uint32_t op;       // bytes produced so far
    uint32_t flushed;  // bytes already stored to global
    int lane;

    __device__ __forceinline__ void init(uint8_t* ring_, uint8_t* dst_, int lane_) {
        ring = ring_;
        dst = dst_;
        a = (uint32_t)((uintptr_t)dst_ & 15u);
        op = 0;
        flushed = 0;
        lane = lane_;
    }
    __device__ __forceinline__ uint32_t ridx(uint32_t p) const { return (p + a) & MASK; }

    // Stores [flushed, upto) to global; the vector body needs (upto + a) % 16 == 0 unless final.
    __device__ __forceinline__ void flush_to(uint32_t upto, bool final) {
        __syncwarp();
        uint32_t q = flushed;
        if (q == 0 && a != 0) {
            uint32_t head = min(16u - a, upto);
            if ((uint32_t)lane < head) dst[lane] = ring[ridx(lane)];
            q = head;
        }
        uint32_t vend = q + ((upto - q) & ~15u);
        for (uint32_t p = q + lane * 16; p < vend; p += 512) {
            uint4 v = *reinterpret_cast<const uint4*>(ring + ridx(p));
            *reinterpret_cast<uint4*>(dst + p) = v;
        }
        q = vend;
        if (final) {
            if (q + lane < upto) dst[q + lane] = ring[ridx(q + lane)];
            q = upto;
        }
        flushed = q;
        __syncwarp();
    }
    __device__ __forceinline__ void make_room() {
        if (op - flushed >= FLUSH_T) flush_to(((op + a) & ~15u) - a, false);
    }
    __device__ __forceinline__ void finish() { flush_to(op, true); }

    // len literal bytes from global src.
    __device__ __forceinline__ void put_literals(const uint8_t* __restrict__ s, uint32_t len) {
        if (len <= 32 && op - flushed < FLUSH_T) {
            if ((uint32_t)lane < len) ring[ridx(op + lane)] = ldg_u8(s + lane);
            op += len;
        } else {
            while (len) {
                uint32_t c = min(len, CHUNK);
                make_room();
                for (uint32_t i = lane; i < c; i += 32) ring[ridx(op + i)] = ldg_u8(s + i);
                op += c;
                s += c;
                len -= c;
            }
        }
        __syncwarp();
    }

    // Back-reference copy; caller guarantees 1 <= off <= op.
    __device__ __forceinline__ void put_match(uint32_t off, uint32_t len) {
        if (len <= 32 && off + 32 <= (uint32_t)RING && op - flushed < FLUSH_T) {
            uint32_t s = op - off;
            uint32_t i = lane;
            if (off < len) i = i % off;
            if ((uint32_t)lane < len) ring[ridx(op + lane)] = ring[ridx(s + i)];
            op += len;
            __syncwarp();
            return;
        }
        while (len) {
            uint32_t c = min(len, CHUNK);
            make_room();
            uint32_t s = op - off;
            if (off + c <= (uint32_t)RING) {
                if (off >= c) {
                    for (uint32_t i = lane; i < c; i += 32) ring[ridx(op + i)] = ring[ridx(s + i)];
                } else {
                    for (uint32_t i = lane; i < c; i += 32) ring[ridx(op + i)] = ring[ridx(s + i % off)];
                }
            } else {  // far: the source was drained to global long ago (off > RING - CHUNK >= c)
                for (uint32_t i = lane; i < c; i += 32) ring[ridx(op + i)] = __ldcg(dst + s + i);
            }
            op += c;
            len -= c;
            __syncwarp();
        }
    }
};

// ------------------------------------------------------------------------------------------------
// LZ4 block.  Same end-of-block rules as LZ4_decompress_safe (MFLIMIT 12, LASTLITERALS 5).
// ------------------------------------------------------------------------------------------------
template <int RING>
__device__ int32_t lz4_decode_block(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap, uint8_t* ring,
                                    int lane, uint32_t* produced) {
    *produced = 0;
    if (n == 0) return CJ_ST_EMPTY;
    if (cap == 0) return (n == 1 && ldg_u8(src) == 0) ? CJ_OK : CJ_ST_DST_SMALL;
    OutRing<RING> out;
    out.init(ring, dst, lane);
    uint32_t ip = 0;
    int32_t st = CJ_OK;
    for (;;) {
        if (ip >= n) { st = CJ_ST_TRUNCATED; break; }
        uint32_t token = ldg_u8(src + ip++);
        uint64_t len = token >> 4;
        if (len == 15) {
            if (n < 15 || ip >= n - 15) { st = CJ_ST_TRUNCATED; break; }
            uint32_t b;
            do {
                b = ldg_u8(src + ip++);
                len += b;
                if (ip > n - 15) { st = CJ_ST_TRUNCATED; break; }
            } while (b == 255);
            if (st) break;
        }
        if ((uint64_t)out.op + len + 12 > cap || (uint64_t)ip + len + 8 > n) {
            // the tail zone of input or output: this must be the final, literal-only sequence
            if ((uint64_t)ip + len != n) { st = ((uint64_t)ip + len > n) ? CJ_ST_TRUNCATED : CJ_ST_CORRUPT; break; }
            if ((uint64_t)out.op + len > cap) { st = CJ_ST_DST_SMALL; break; }
            out.put_literals(src + ip, (uint32_t)len);
            break;
        }
        out.put_literals(src + ip, (uint32_t)len);
        ip += (uint32_t)len;
        uint32_t off = ldg_u8(src + ip) | (ldg_u8(src + ip + 1) << 8);
        ip += 2;
        len = token & 15;
        if (len == 15) {
            uint32_t b;
            do {
                b = ldg_u8(src + ip++);
                len += b;
                if (ip > n - 4) { st = CJ_ST_TRUNCATED; break; }
            } while (b == 255);
            if (st) break;
        }
        len += 4;
        if (off == 0 || off > out.op) { st = CJ_ST_OFFSET; break; }
        if ((uint64_t)out.op + len + 5 > cap) { st = CJ_ST_DST_SMALL; break; }
        out.put_match(off, (uint32_t)len);
    }
    out.finish();
    *produced = out.op;
    return st;
}

// ------------------------------------------------------------------------------------------------
// Snappy raw block.
// ------------------------------------------------------------------------------------------------
template <int RING>
__device__ int32_t snappy_decode_block(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap, uint8_t* ring,
                                       int lane, uint32_t* produced) {
    *produced = 0;
    if (n == 0) return CJ_ST_EMPTY;
    // uvarint32 preamble
    uint64_t ulen = 0;
    uint32_t ip = 0;
    {
        bool done = false;
        for (int i = 0; i < 5 && ip < n; i++) {
            uint32_t b = ldg_u8(src + ip++);
            ulen |= (uint64_t)(b & 0x7f) << (7 * i);
            if (!(b & 0x80)) { done = true; break; }
        }
        if (!done) return CJ_ST_HEADER;
    }
    if (ulen > 0xFFFFFFFFull) return CJ_ST_TOO_BIG;
    if (ulen > cap) return CJ_ST_DST_SMALL;
    const uint32_t dn = (uint32_t)ulen;
    OutRing<RING> out;
    out.init(ring, dst, lane);
    int32_t st = CJ_OK;
    while (ip < n) {
        uint32_t tag = ldg_u8(src + ip++);
        uint32_t type = tag & 3;
        if (type == 0) {
            uint64_t len = (tag >> 2) + 1;
            if (len > 60) {
                uint32_t nb = (uint32_t)len - 60;
                if (nb > n - ip) { st = CJ_ST_TRUNCATED; break; }
                uint32_t v = 0;
                for (uint32_t i = 0; i < nb; i++) v |= ldg_u8(src + ip + i) << (8 * i);
                ip += nb;
                len = (uint64_t)v + 1;
            }
            if (len > n - ip) { st = CJ_ST_TRUNCATED; break; }
            if (len > dn - out.op) { st = CJ_ST_LEN_MISMATCH; break; }
            out.put_literals(src + ip, (uint32_t)len);
            ip += (uint32_t)len;
            continue;
        }
        uint32_t len, off;
        if (type == 1) {
            if (n - ip < 1) { st = CJ_ST_TRUNCATED; break; }
            len = 4 + ((tag >> 2) & 7);
            off = ((tag >> 5) << 8) | ldg_u8(src + ip);
            ip += 1;
        } else if (type == 2) {
            if (n - ip < 2) { st = CJ_ST_TRUNCATED; break; }
            len = 1 + (tag >> 2);
            off = ldg_u8(src + ip) | (ldg_u8(src + ip + 1) << 8);
            ip += 2;
        } else {
            if (n - ip < 4) { st = CJ_ST_TRUNCATED; break; }
            len = 1 + (tag >> 2);
            off = ldg_u8(src + ip) | (ldg_u8(src + ip + 1) << 8) | (ldg_u8(src + ip + 2) << 16) | (ldg_u8(src + ip + 3) << 24);
            ip += 4;
        }
        if (off == 0 || off > out.op) { st = CJ_ST_OFFSET; break; }
        if (len > dn - out.op) { st = CJ_ST_LEN_MISMATCH; break; }
        out.put_match(off, len);
    }
    if (st == CJ_OK && out.op != dn) st = CJ_ST_LEN_MISMATCH;
    out.finish();
    *produced = out.op;
    return st;
}

template <int CODEC, int RING, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) lz_decode_kernel(Batch b, unsigned* __restrict__ counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* ring = smem + (size_t)warp * RING;
    for (;;) {
        const uint32_t u = next_unit(counter, lane);
        if (u >= b.n) break;
        const uint64_t slen = b.src_len[u], dcap = b.dst_cap[u];
        const uint8_t* src = b.src_base + b.src_off[u];
        uint8_t* dst = b.dst_base + b.dst_off[u];
        uint32_t produced = 0;
        int32_t st;
        if (slen > MAX_UNIT) {
            st = CJ_ST_TOO_BIG;
        } else {
            const uint32_t cap = dcap > MAX_UNIT ? MAX_UNIT : (uint32_t)dcap;
            if (CODEC == CJ_LZ4_BLOCK) st = lz4_decode_block<RING>(src, (uint32_t)slen, dst, cap, ring, lane, &produced);
            else st = snappy_decode_block<RING>(src, (uint32_t)slen, dst, cap, ring, lane, &produced);
        }
        if (lane == 0) {
            b.dst_len[u] = st == CJ_OK ? produced : 0;
            b.status[u] = st;
        }
        __syncwarp();
    }
}

constexpr int DEC_RING = 16384;
constexpr int DEC_WARPS = 4;

cudaError_t launch_lz_decode(int codec, const Batch& b, unsigned* counter, int sm_count, cudaStream_t stream) {
    const size_t smem = (size_t)DEC_RING * DEC_WARPS;
    auto k = codec == CJ_LZ4_BLOCK ? lz_decode_kernel<CJ_LZ4_BLOCK, DEC_RING, DEC_WARPS> : lz_decode_kernel<CJ_SNAPPY_RAW, DEC_RING, DEC_WARPS>;
    static bool attr_done[2] = {false, false};
    const int which = codec == CJ_LZ4_BLOCK ? 0 : 1;
    if (!attr_done[which]) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done[which] = true;
    }
    const int ctas_per_sm = 3;
    int grid = sm_count * ctas_per_sm;
    const int need = (int)((b.n + DEC_WARPS - 1) / DEC_WARPS);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    k<<<grid, DEC_WARPS * 32, smem, stream>>>(b, counter);
    return cudaGetLastError();
}

}  // namespace cj
