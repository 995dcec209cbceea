// frames.cu — frame containers around the block codecs (sm_100a + host plumbing).
//
//   snappy framing format  : cramjam.snappy.compress/decompress(_into)  reference src/snappy.rs:22-42,81-90
//                            (snap::read::FrameEncoder / FrameDecoder)
//   LZ4 frame (LZ4F)       : cramjam.lz4.compress/decompress(_into)     reference src/lz4.rs:27-65 (lz4::Encoder/Decoder)
//
// Snappy framed streams are cut into their (independent, <= 64 KiB) chunks on the host — a walk
// over 4-byte chunk headers, no payload byte is interpreted on the CPU — and every chunk becomes
// one unit of the batched raw-block kernels; masked CRC-32C of each chunk is computed on the device
// and checked / emitted.  LZ4 frames are decoded by one warp per frame walking the blocks on the
// device (linked blocks depend on the previous 64 KiB of output, so blocks of one frame run in
// order); XXH32 header / block / content checksums are verified in the same kernel.  The encoders
// emit independent 64 KiB blocks (legal LZ4F), one unit of the block encoder each.
#include <chrono>
#include <functional>
#include <future>

#include "internal.h"
#include "lz_decode.cuh"

namespace cj {

// ================================================================================================
// Checksums on the device
// ================================================================================================
#define XP1 2654435761u
#define XP2 2246822519u
#define XP3 3266489917u
#define XP4 668265263u
#define XP5 374761393u

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

template <class LoadByte>
__host__ __device__ __forceinline__ uint32_t xxh32_generic(LoadByte ld, uint64_t n, uint32_t seed) {
    uint64_t p = 0;
    uint32_t h;
    auto rd = [&](uint64_t q) { return (uint32_t)ld(q) | ((uint32_t)ld(q + 1) << 8) | ((uint32_t)ld(q + 2) << 16) | ((uint32_t)ld(q + 3) << 24); };
    if (n >= 16) {
        uint32_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
        do {
            v1 = rotl32(v1 + rd(p) * XP2, 13) * XP1;
            v2 = rotl32(v2 + rd(p + 4) * XP2, 13) * XP1;
            v3 = rotl32(v3 + rd(p + 8) * XP2, 13) * XP1;
            v4 = rotl32(v4 + rd(p + 12) * XP2, 13) * XP1;
            p += 16;
        } while (p + 16 <= n);
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = seed + XP5;
    }
    h += (uint32_t)n;
    while (p + 4 <= n) { h = rotl32(h + rd(p) * XP3, 17) * XP4; p += 4; }
    while (p < n) { h = rotl32(h + (uint32_t)ld(p) * XP5, 11) * XP1; p++; }
    h ^= h >> 15; h *= XP2; h ^= h >> 13; h *= XP3; h ^= h >> 16;
    return h;
}

// XXH32 of a global-memory range; word loads when the range is 4-byte aligned.
__device__ uint32_t xxh32_global(const uint8_t* p, uint64_t n) {
    if (((uintptr_t)p & 3) == 0) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
        uint64_t i = 0;
        uint32_t h;
        if (n >= 16) {
            uint32_t v1 = XP1 + XP2, v2 = XP2, v3 = 0, v4 = 0u - XP1;
            do {
                const uint4 q = __ldcg(reinterpret_cast<const uint4*>(w + i));  // 16 B aligned when p is; else falls to scalar path below
                v1 = rotl32(v1 + q.x * XP2, 13) * XP1;
                v2 = rotl32(v2 + q.y * XP2, 13) * XP1;
                v3 = rotl32(v3 + q.z * XP2, 13) * XP1;
                v4 = rotl32(v4 + q.w * XP2, 13) * XP1;
                i += 4;
            } while (i * 4 + 16 <= n);
            h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
        } else {
            h = XP5;
        }
        h += (uint32_t)n;
        uint64_t b = i * 4;
        while (b + 4 <= n) { h = rotl32(h + __ldcg(w + b / 4) * XP3, 17) * XP4; b += 4; }
        while (b < n) { h = rotl32(h + (uint32_t)__ldcg(p + b) * XP5, 11) * XP1; b++; }
        h ^= h >> 15; h *= XP2; h ^= h >> 13; h *= XP3; h ^= h >> 16;
        return h;
    }
    return xxh32_generic([&](uint64_t q) { return __ldcg(p + q); }, n, 0);
}

// XXH32 of every unit, one warp per unit.  The four accumulators of XXH32 are four independent serial chains
// (acc = rotl(acc + w * P2, 13) * P1 over every fourth word), so one unit cannot go faster than one chain step per
// 16 bytes; what the warp adds is the memory side: all 32 lanes stream the unit in 512-byte rows (one coalesced
// 16-byte load per lane, the next row in flight while the current one is hashed), rows are parked in shared memory
// and lanes 0..3 run one accumulator each.  (The one-thread-per-unit loop this replaces exposed a full DRAM round
// trip per 16 bytes: 5 s for a 256 MiB frame; this kernel is chain-bound at ~2 GB/s per unit.)
// Warp-cooperative XXH32 of p[0, L) (p 16-byte aligned, L2-coherent loads); `rows` = 2 x 128 words of this warp's shared memory.
// Every lane returns the hash.  Rows are stored transposed (the 32 words of accumulator i contiguous) so that lane i pulls its
// words with eight 16-byte loads and runs its chain in registers.  The chain itself is folded to two dependent operations per
// step: with s = acc + w*P2, rotl(s,13)*P1 = s*(P1<<13) + (s>>19)*P1 (the two halves of the rotation occupy disjoint bits),
// so the next s is (s>>19)*P1 + (s*(P1<<13) + w'*P2) — a shift and a multiply-add side by side, then one multiply-add.
__device__ __forceinline__ uint32_t xxh32_warp(const uint8_t* p, uint64_t L, uint32_t* rows, int lane) {
    const uint64_t stripes = L / 16;           // 16-byte stripes the main loop consumes
    const uint64_t nrows = (stripes + 31) / 32;
    constexpr uint32_t K13 = XP1 << 13;
    uint32_t acc = lane == 0 ? XP1 + XP2 : (lane == 1 ? XP2 : (lane == 2 ? 0u : 0u - XP1));
    const uint4* g = reinterpret_cast<const uint4*>(p);
    // Four rows (2 KiB) in flight per warp — a single warp has to cover the memory latency by itself — held in four
    // registers that are never copied into each other (a rotation would wait for the newest load every row).
    uint4 nq[4];
#pragma unroll
    for (int j = 0; j < 4; j++) nq[j] = (uint64_t)lane + 32 * j < stripes ? __ldcg(g + lane + 32 * j) : make_uint4(0, 0, 0, 0);
    __syncwarp();
    auto one_row = [&](uint64_t r, uint4& q) {
        uint32_t* row = rows + (r & 1) * 128;
        // every lane pre-multiplies its four words by P2 (off the hashing lanes' critical path)
        row[lane] = q.x * XP2; row[32 + lane] = q.y * XP2; row[64 + lane] = q.z * XP2; row[96 + lane] = q.w * XP2;
        const uint64_t s1 = (r + 4) * 32 + lane;
        if (s1 < stripes) q = __ldcg(g + s1);
        __syncwarp();
        const uint32_t cnt = (uint32_t)min((uint64_t)32, stripes - r * 32);
        if (lane < 4) {
            uint32_t w[32];
            const uint4* mine = reinterpret_cast<const uint4*>(row + lane * 32);
#pragma unroll
            for (int k = 0; k < 8; k++) { const uint4 t = mine[k]; w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w; }
            if (cnt == 32) {
                uint32_t sv = acc + w[0];
#pragma unroll
                for (int k = 1; k < 32; k++) sv = (sv >> 19) * XP1 + (sv * K13 + w[k]);
                acc = rotl32(sv, 13) * XP1;
            } else {
#pragma unroll
                for (int k = 0; k < 32; k++)
                    if ((uint32_t)k < cnt) acc = rotl32(acc + w[k], 13) * XP1;
            }
        }
        // rows alternate between two buffers; a buffer is rewritten only after the __syncwarp of the row in between
    };
    uint64_t r = 0;
    for (; r + 4 <= nrows; r += 4) {
        one_row(r, nq[0]); one_row(r + 1, nq[1]); one_row(r + 2, nq[2]); one_row(r + 3, nq[3]);
    }
    if (r < nrows) one_row(r, nq[0]);
    if (r + 1 < nrows) one_row(r + 1, nq[1]);
    if (r + 2 < nrows) one_row(r + 2, nq[2]);
    const uint32_t v1 = __shfl_sync(FULL, acc, 0), v2 = __shfl_sync(FULL, acc, 1), v3 = __shfl_sync(FULL, acc, 2), v4 = __shfl_sync(FULL, acc, 3);
    uint32_t h = L >= 16 ? rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18) : XP5;
    h += (uint32_t)L;
    if (lane == 0) {
        uint64_t b = stripes * 16;
        while (b + 4 <= L) { h = rotl32(h + __ldcg(reinterpret_cast<const uint32_t*>(p + b)) * XP3, 17) * XP4; b += 4; }
        while (b < L) { h = rotl32(h + (uint32_t)__ldcg(p + b) * XP5, 11) * XP1; b++; }
        h ^= h >> 15; h *= XP2; h ^= h >> 13; h *= XP3; h ^= h >> 16;
    }
    __syncwarp();
    return __shfl_sync(FULL, h, 0);
}

constexpr int XXH_WARPS = 4;
__global__ void __launch_bounds__(XXH_WARPS * 32) xxh32_units_kernel(uint32_t n, const uint8_t* __restrict__ base, const uint64_t* __restrict__ off,
                                                                      const uint64_t* __restrict__ len, uint32_t* __restrict__ out) {
    __shared__ __align__(16) uint32_t rows[XXH_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t u = blockIdx.x * XXH_WARPS + warp;
    if (u >= n) return;
    const uint8_t* p = base + off[u];
    const uint64_t L = len[u];
    uint32_t h;
    if (((uintptr_t)p & 15) != 0) {  // unaligned unit: byte-wise, one lane (engine-made arenas are aligned)
        h = lane == 0 ? xxh32_generic([&](uint64_t q) { return __ldcg(p + q); }, L, 0) : 0u;
    } else {
        h = xxh32_warp(p, L, rows[warp], lane);
    }
    if (lane == 0) out[u] = h;
}

// CRC-32C (Castagnoli), slicing-by-8 tables built in shared memory by each CTA; one thread per unit.
__global__ void __launch_bounds__(128) crc32c_units_kernel(uint32_t n, const uint8_t* __restrict__ base, const uint64_t* __restrict__ off,
                                                           const uint64_t* __restrict__ len, uint32_t* __restrict__ out_masked) {
    __shared__ uint32_t T[8][256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
        T[0][i] = c;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = T[0][i];
        for (int t = 1; t < 8; t++) {
            c = (c >> 8) ^ T[0][c & 0xff];
            T[t][i] = c;
        }
    }
    __syncthreads();
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const uint8_t* p = base + off[u];
    uint64_t L = len[u];
    uint32_t c = 0xFFFFFFFFu;
    while (L && ((uintptr_t)p & 7)) { c = (c >> 8) ^ T[0][(c ^ __ldcg(p)) & 0xff]; p++; L--; }
    while (L >= 8) {
        const uint2 w = __ldcg(reinterpret_cast<const uint2*>(p));
        const uint32_t lo = w.x ^ c, hi = w.y;
        c = T[7][lo & 0xff] ^ T[6][(lo >> 8) & 0xff] ^ T[5][(lo >> 16) & 0xff] ^ T[4][lo >> 24] ^ T[3][hi & 0xff] ^ T[2][(hi >> 8) & 0xff] ^
            T[1][(hi >> 16) & 0xff] ^ T[0][hi >> 24];
        p += 8;
        L -= 8;
    }
    while (L) { c = (c >> 8) ^ T[0][(c ^ __ldcg(p)) & 0xff]; p++; L--; }
    c ^= 0xFFFFFFFFu;
    out_masked[u] = ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

// ================================================================================================
// LZ4 frame decode: one warp per frame stream (concatenated + skippable frames included)
// ================================================================================================
__device__ int32_t lz4f_decode_stream(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap, uint8_t* smem_warp, int lane,
                                      uint32_t* produced) {
    OutRing out;
    out.init(smem_warp, dst, lane);
    uint32_t ip = 0;
    int32_t st = CJ_OK;
    while (ip < n && st == CJ_OK) {
        if (n - ip < 4) { st = CJ_ST_TRUNCATED; break; }
        const uint32_t magic = rd32g(src + ip);
        if (magic >= 0x184D2A50u && magic <= 0x184D2A5Fu) {  // skippable frame
            if (n - ip < 8) { st = CJ_ST_TRUNCATED; break; }
            const uint32_t sz = rd32g(src + ip + 4);
            if (sz > n - ip - 8) { st = CJ_ST_TRUNCATED; break; }
            ip += 8 + sz;
            continue;
        }
        if (magic != 0x184D2204u) { st = CJ_ST_HEADER; break; }
        if (n - ip < 7) { st = CJ_ST_TRUNCATED; break; }
        const uint32_t flg = __ldg(src + ip + 4), bd = __ldg(src + ip + 5);
        if ((flg >> 6) != 1 || (flg & 0x02) || (bd & 0x8F)) { st = CJ_ST_HEADER; break; }
        const bool indep = (flg >> 5) & 1, bsum = (flg >> 4) & 1, csize = (flg >> 3) & 1, csum = (flg >> 2) & 1, dict = flg & 1;
        const uint32_t bid = (bd >> 4) & 7;
        if (bid < 4) { st = CJ_ST_HEADER; break; }
        const uint32_t bmax = 1u << (8 + 2 * bid);
        const uint32_t dlen = 2 + (csize ? 8 : 0) + (dict ? 4 : 0);
        if (n - ip < 4 + dlen + 1) { st = CJ_ST_TRUNCATED; break; }
        const uint8_t* desc = src + ip + 4;
        if (((xxh32_generic([&](uint64_t q) { return __ldg(desc + q); }, dlen, 0) >> 8) & 0xff) != __ldg(desc + dlen)) { st = CJ_ST_CHECKSUM; break; }
        uint64_t content_size = 0;
        if (csize) content_size = (uint64_t)rd32g(desc + 2) | ((uint64_t)rd32g(desc + 6) << 32);
        if (dict) { st = CJ_ST_UNSUPPORTED; break; }
        ip += 4 + dlen + 1;
        const uint32_t frame_start = out.op;
        for (;;) {
            if (n - ip < 4) { st = CJ_ST_TRUNCATED; break; }
            const uint32_t bs = rd32g(src + ip);
            ip += 4;
            if (bs == 0) break;
            const bool stored = bs >> 31;
            const uint32_t blen = bs & 0x7FFFFFFFu;
            if (blen > bmax) { st = CJ_ST_CORRUPT; break; }
            if (blen > n - ip) { st = CJ_ST_TRUNCATED; break; }
            if (bsum) {
                if (n - ip - blen < 4) { st = CJ_ST_TRUNCATED; break; }
                uint32_t h = 0;
                const uint8_t* bp = src + ip;
                if (lane == 0) h = xxh32_generic([&](uint64_t q) { return __ldg(bp + q); }, blen, 0);
                h = __shfl_sync(FULL, h, 0);
                if (h != rd32g(src + ip + blen)) { st = CJ_ST_CHECKSUM; break; }
            }
            if (stored) {
                if (blen > cap - out.op) { st = CJ_ST_DST_SMALL; break; }
                out.put_literals(src + ip, blen);
            } else {
                if (blen == 0) { st = CJ_ST_EMPTY; break; }
                out.base = indep ? out.op : frame_start;
                const uint32_t room = min(cap - out.op, bmax);
                st = decode_stream<CJ_LZ4_BLOCK, true>(src + ip, blen, 0, out.op + room, out, smem_warp, lane);
                if (st != CJ_OK) {
                    if (st == CJ_ST_DST_SMALL && room == bmax) st = CJ_ST_CORRUPT;  // block larger than the frame's block size
                    break;
                }
            }
            ip += blen + (bsum ? 4 : 0);
        }
        if (st != CJ_OK) break;
        if (csum) {
            if (n - ip < 4) { st = CJ_ST_TRUNCATED; break; }
            out.flush_to(out.op, true);
            uint32_t h = 0;
            const uint8_t* fp = dst + frame_start;
            if (((uintptr_t)fp & 15) == 0) {   // the ring is idle between frames: its first KiB parks the rows being hashed
                __threadfence_block();
                h = xxh32_warp(fp, out.op - frame_start, reinterpret_cast<uint32_t*>(smem_warp), lane);
            } else {
                if (lane == 0) h = xxh32_global(fp, out.op - frame_start);
                h = __shfl_sync(FULL, h, 0);
            }
            if (h != rd32g(src + ip)) { st = CJ_ST_CHECKSUM; break; }
            ip += 4;
        }
        if (csize && content_size != (uint64_t)(out.op - frame_start)) { st = CJ_ST_LEN_MISMATCH; break; }
    }
    out.finish();
    *produced = out.op;
    return st;
}

__global__ void __launch_bounds__(DEC_WARPS * 32, CJ_DEC_CTAS) lz4f_decode_kernel(Batch b, unsigned* __restrict__ counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* smem_warp = smem + (size_t)warp * DEC_SMEM_WARP;
    ring_barrier_init(smem_warp, lane);
    for (;;) {
        const uint32_t u = next_unit(counter, lane);
        if (u >= b.n) break;
        const uint64_t slen = b.src_len[u], dcap = b.dst_cap[u];
        uint32_t produced = 0;
        int32_t st;
        if (slen > MAX_UNIT) st = CJ_ST_TOO_BIG;
        else st = lz4f_decode_stream(b.src_base + b.src_off[u], (uint32_t)slen, b.dst_base + b.dst_off[u], dcap > MAX_UNIT ? MAX_UNIT : (uint32_t)dcap,
                                     smem_warp, lane, &produced);
        if (lane == 0) {
            b.dst_len[u] = st == CJ_OK ? produced : 0;
            b.status[u] = st;
        }
        __syncwarp();
    }
}

cudaError_t launch_lz4f_decode(const Batch& b, unsigned* counter, int sm_count, cudaStream_t stream) {
    const size_t smem = (size_t)DEC_SMEM_WARP * DEC_WARPS;
    static cj_per_device_flag attr_flag;
    int& attr_done = attr_flag.here();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(lz4f_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = 1;
    }
    int grid = sm_count * CJ_DEC_CTAS;
    const int need = (int)((b.n + DEC_WARPS - 1) / DEC_WARPS);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    lz4f_decode_kernel<<<grid, DEC_WARPS * 32, smem, stream>>>(b, counter);
    return cudaGetLastError();
}

// ================================================================================================
// Host plumbing
// ================================================================================================
namespace {

// CJ_TRACE=1: wall-clock checkpoints of the frame paths on stderr (each one drains the stream first).
struct PhaseTrace {
    cj_ctx* c;
    const char* what;
    bool on;
    std::chrono::steady_clock::time_point t0, last;
    PhaseTrace(cj_ctx* c_, const char* w) : c(c_), what(w), on(getenv("CJ_TRACE") != nullptr) { t0 = last = std::chrono::steady_clock::now(); }
    void mark(const char* label) {
        if (!on) return;
        cudaStreamSynchronize(c->stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[cj] %s: %-28s +%8.2f ms (total %8.2f)\n", what, label, std::chrono::duration<double, std::milli>(now - last).count(),
                std::chrono::duration<double, std::milli>(now - t0).count());
        last = now;
    }
};

inline uint32_t h_rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline void h_wr32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
inline uint32_t h_xxh32(const uint8_t* p, size_t n) { return xxh32_generic([&](uint64_t q) { return p[q]; }, n, 0); }

// XXH32's four accumulators are serial chains (acc = rotl(acc + w * P2, 13) * P1): one large frame cannot be hashed faster
// than one chain step per 16 bytes, ~1.5 GB/s at GPU clocks (xxh32_warp below) against ~6 GB/s on one host core.  A frame
// of at least HOST_HASH_MIN bytes whose bytes are on the host anyway (CJ_HOST / CJ_PINNED input of the encoder, output of
// the decoder on its way home) is therefore hashed there, on a thread of its own next to the GPU work or the copy home —
// a checksum of host-resident bytes, no codec step.  Device-resident frames (CJ_DEVICE) and small ones stay on the kernel.
constexpr size_t HOST_HASH_MIN = (size_t)1 << 20;
inline uint32_t h_xxh32_words(const uint8_t* p, size_t n) {
    const uint8_t* const end = p + n;
    uint32_t h;
    if (n >= 16) {
        uint32_t v1 = XP1 + XP2, v2 = XP2, v3 = 0, v4 = 0u - XP1;
        const uint8_t* const lim = end - 16;
        do {
            uint32_t w[4];
            memcpy(w, p, 16);
            v1 = rotl32(v1 + w[0] * XP2, 13) * XP1;
            v2 = rotl32(v2 + w[1] * XP2, 13) * XP1;
            v3 = rotl32(v3 + w[2] * XP2, 13) * XP1;
            v4 = rotl32(v4 + w[3] * XP2, 13) * XP1;
            p += 16;
        } while (p <= lim);
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = XP5;
    }
    h += (uint32_t)n;
    while (p + 4 <= end) { uint32_t w; memcpy(&w, p, 4); h = rotl32(h + w * XP3, 17) * XP4; p += 4; }
    while (p < end) { h = rotl32(h + (uint32_t)*p * XP5, 11) * XP1; p++; }
    h ^= h >> 15; h *= XP2; h ^= h >> 13; h *= XP3; h ^= h >> 16;
    return h;
}

// Chunk / block size of the snappy-framed and LZ4-frame encoders.  Both formats allow chunks smaller than 64 KiB; one warp
// needs ~4 ms to compress (1 ms to decompress) 64 KiB, so a small input is cut finer to put more warps on it:
// 64 KiB from 4 MiB up, else n / 64 rounded up to 4 KiB, at least 16 KiB (1 MiB -> 64 chunks of 16 KiB).
inline uint64_t frame_piece(uint64_t n) {
    if (n >= ((uint64_t)4 << 20)) return 65536;
    const uint64_t p = (n / 64 + 4095) & ~(uint64_t)4095;
    return std::min<uint64_t>(65536, std::max<uint64_t>(16384, p));
}

// A flat list of device work items referring to one device source arena and one device destination arena.
struct Items {
    std::vector<uint64_t> so, sl, dof, dc;
    size_t size() const { return so.size(); }
    void add(uint64_t a, uint64_t b, uint64_t c, uint64_t d) { so.push_back(a); sl.push_back(b); dof.push_back(c); dc.push_back(d); }
};

// Uploads item descriptors; returns device pointers inside ctx->f_ddesc (layout: so | sl | dof | dc | dl | st(i32) | aux(u32)).
struct DevItems {
    uint64_t *so, *sl, *dof, *dc, *dl;
    int32_t* st;
    uint32_t* aux;
    uint64_t* h;  // pinned host mirror, same layout
    size_t n;
};

int upload_items(cj_ctx* c, const Items& it, Scratch& dsc, Scratch& hsc, DevItems* out) {
    const size_t n = it.size();
    const size_t bytes = n * (5 * 8 + 4 + 4) + 64;
    int rc;
    if ((rc = dsc.ensure(bytes))) return rc;
    if ((rc = hsc.ensure(bytes))) return rc;
    uint64_t* h = (uint64_t*)hsc.p;
    if (n) {
        memcpy(h, it.so.data(), n * 8);
        memcpy(h + n, it.sl.data(), n * 8);
        memcpy(h + 2 * n, it.dof.data(), n * 8);
        memcpy(h + 3 * n, it.dc.data(), n * 8);
        CUDA_TRY(cudaMemcpyAsync(dsc.p, h, n * 32, cudaMemcpyHostToDevice, c->stream));
    }
    uint64_t* d = (uint64_t*)dsc.p;
    out->so = d; out->sl = d + n; out->dof = d + 2 * n; out->dc = d + 3 * n; out->dl = d + 4 * n;
    out->st = (int32_t*)(d + 5 * n);
    out->aux = (uint32_t*)(out->st + n);
    out->h = h;
    out->n = n;
    return CJ_OK;
}

int fetch_results(cj_ctx* c, const DevItems& d) {  // dl | st | aux back to the pinned mirror
    if (d.n) CUDA_TRY(cudaMemcpyAsync(d.h + 4 * d.n, d.dl, d.n * 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CJ_OK;
}

// Copies every unit's source bytes into a device arena (16-byte aligned starts); base[i] = arena offset of unit i.
int upload_units(cj_ctx* c, const cj_batch* bt, int where, std::vector<uint64_t>& base, size_t extra_tail = 64) {
    const size_t n = bt->n;
    base.resize(n);
    size_t acc = 0;
    for (size_t i = 0; i < n; i++) { base[i] = acc; acc += cj_align16((size_t)bt->src_len[i]); }
    int rc;
    if ((rc = c->f_dsrc.ensure(acc + extra_tail))) return rc;
    const uint8_t* hs = (const uint8_t*)bt->src_base;
    if (where == CJ_PINNED) {
        for (size_t i = 0; i < n; i++)
            if (bt->src_len[i]) CUDA_TRY(cudaMemcpyAsync((uint8_t*)c->f_dsrc.p + base[i], hs + bt->src_off[i], (size_t)bt->src_len[i], cudaMemcpyHostToDevice, c->stream));
    } else {
        if ((rc = c->f_hsrc.ensure(acc + extra_tail))) return rc;
        uint8_t* stage = (uint8_t*)c->f_hsrc.p;
        if (n <= 4) { for (size_t i = 0; i < n; i++) cj_parallel_copy(stage + base[i], hs + bt->src_off[i], (size_t)bt->src_len[i]); }
        else cj_parallel_units(n, acc, [&](size_t i) { memcpy(stage + base[i], hs + bt->src_off[i], (size_t)bt->src_len[i]); });
        if (acc) CUDA_TRY(cudaMemcpyAsync(c->f_dsrc.p, stage, acc, cudaMemcpyHostToDevice, c->stream));
    }
    return CJ_OK;
}

// Copies produced bytes of every OK unit from the device destination arena back to the caller.
// `landed(unit_bytes)`, if given, runs on a thread of its own once the units' bytes are in host memory — next to the copy
// from the pinned staging arena into the caller's pages (CJ_HOST) — with unit_bytes(i) = where unit i's bytes can be read.
int download_units(cj_ctx* c, const cj_batch* bt, int where, const std::vector<uint64_t>& dbase, size_t arena_bytes,
                   const std::function<void(const std::function<const uint8_t*(size_t)>&)>& landed = nullptr) {
    const size_t n = bt->n;
    uint8_t* hd = (uint8_t*)bt->dst_base;
    if (where == CJ_PINNED) {
        for (size_t i = 0; i < n; i++)
            if (bt->status[i] == CJ_OK && bt->dst_len[i])
                CUDA_TRY(cudaMemcpyAsync(hd + bt->dst_off[i], (uint8_t*)c->f_ddst.p + dbase[i], (size_t)bt->dst_len[i], cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (landed) landed([&](size_t i) { return (const uint8_t*)hd + bt->dst_off[i]; });
    } else {
        int rc;
        if ((rc = c->f_hdst.ensure(arena_bytes + 64))) return rc;
        if (arena_bytes) CUDA_TRY(cudaMemcpyAsync(c->f_hdst.p, c->f_ddst.p, arena_bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const uint8_t* stage = (const uint8_t*)c->f_hdst.p;
        std::thread side;
        if (landed) side = std::thread([&]() { landed([&](size_t i) { return stage + dbase[i]; }); });
        struct Join { std::thread& t; ~Join() { if (t.joinable()) t.join(); } } join{side};
        if (n <= 4) {
            for (size_t i = 0; i < n; i++)
                if (bt->status[i] == CJ_OK && bt->dst_len[i]) cj_parallel_copy(hd + bt->dst_off[i], stage + dbase[i], (size_t)bt->dst_len[i]);
        } else {
            cj_parallel_units(n, arena_bytes, [&](size_t i) {
                if (bt->status[i] == CJ_OK && bt->dst_len[i]) memcpy(hd + bt->dst_off[i], stage + dbase[i], (size_t)bt->dst_len[i]);
            });
        }
    }
    return CJ_OK;
}

int launch_crc(cj_ctx* c, uint32_t n, const uint8_t* base, const uint64_t* off, const uint64_t* len, uint32_t* out) {
    if (!n) return CJ_OK;
    crc32c_units_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(n, base, off, len, out);
    c->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return CJ_OK;
}

int launch_xxh32(cj_ctx* c, uint32_t n, const uint8_t* base, const uint64_t* off, const uint64_t* len, uint32_t* out) {
    if (!n) return CJ_OK;
    xxh32_units_kernel<<<(n + XXH_WARPS - 1) / XXH_WARPS, XXH_WARPS * 32, 0, c->stream>>>(n, base, off, len, out);
    c->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return CJ_OK;
}

int copy_units(cj_ctx* c, uint32_t n, const uint8_t* sb, const uint64_t* so, const uint64_t* len, uint8_t* db, const uint64_t* dof) {
    if (!n) return CJ_OK;
    cudaError_t e = launch_copy_units(n, sb, so, len, db, dof, c->sm_count, c->stream);
    c->launches += 1;
    if (e != cudaSuccess) { cj_set_error("copy_units launch failed: %s", cudaGetErrorString(e)); return CJ_E_CUDA; }
    return CJ_OK;
}

// ---- snappy framing: host walk over chunk headers ---------------------------------------------
struct SnChunk { uint64_t body_off; uint32_t body_len; uint32_t ulen; uint32_t crc; bool compressed; };

int32_t snappy_uvarint(const uint8_t* p, size_t n, uint64_t* v) {  // returns header bytes or 0
    uint64_t r = 0;
    for (int i = 0; i < 5 && (size_t)i < n; i++) {
        r |= (uint64_t)(p[i] & 0x7f) << (7 * i);
        if (!(p[i] & 0x80)) { *v = r; return i + 1; }
    }
    return 0;
}

int32_t snappy_frame_walk(const uint8_t* s, size_t n, std::vector<SnChunk>* chunks, uint64_t* total) {
    size_t p = 0;
    bool seen = false;
    uint64_t tot = 0;
    while (p < n) {
        if (n - p < 4) return CJ_ST_TRUNCATED;
        const uint8_t type = s[p];
        const size_t len = (size_t)s[p + 1] | ((size_t)s[p + 2] << 8) | ((size_t)s[p + 3] << 16);
        p += 4;
        if (len > n - p) return CJ_ST_TRUNCATED;
        if (!seen && type != 0xff) return CJ_ST_HEADER;
        if (type == 0xff) {
            if (len != 6 || memcmp(s + p, "sNaPpY", 6) != 0) return CJ_ST_HEADER;
            seen = true;
        } else if (type == 0x00 || type == 0x01) {
            if (len < 4) return CJ_ST_CORRUPT;
            SnChunk c;
            c.crc = h_rd32(s + p);
            c.body_off = p + 4;
            c.body_len = (uint32_t)(len - 4);
            c.compressed = type == 0x00;
            if (c.compressed) {
                uint64_t u;
                if (c.body_len == 0) return CJ_ST_EMPTY;
                if (!snappy_uvarint(s + p + 4, c.body_len, &u)) return CJ_ST_HEADER;
                if (u > 0xFFFFFFFFull) return CJ_ST_TOO_BIG;
                if (u > 65536) return CJ_ST_CORRUPT;
                c.ulen = (uint32_t)u;
            } else {
                if (c.body_len > 65536) return CJ_ST_CORRUPT;
                c.ulen = c.body_len;
            }
            tot += c.ulen;
            if (chunks) chunks->push_back(c);
        } else if (type < 0x80) {
            return CJ_ST_CORRUPT;  // 0x02-0x7f reserved unskippable
        }
        p += len;
    }
    *total = tot;
    return CJ_OK;
}

int snappy_framed_decompress(cj_ctx* c, int where, const cj_batch* bt) {
    PhaseTrace tr(c, "snappy_framed_decompress");
    const size_t n = bt->n;
    const uint8_t* hs = (const uint8_t*)bt->src_base;
    std::vector<uint64_t> sbase, dbase(n);
    std::vector<std::vector<SnChunk>> chunks(n);
    std::vector<uint64_t> total(n, 0);
    size_t dacc = 0;
    for (size_t i = 0; i < n; i++) {
        bt->status[i] = snappy_frame_walk(hs + bt->src_off[i], (size_t)bt->src_len[i], &chunks[i], &total[i]);
        bt->dst_len[i] = 0;
        if (bt->status[i] == CJ_OK && total[i] > bt->dst_cap[i]) bt->status[i] = CJ_ST_DST_SMALL;
        dbase[i] = dacc;
        if (bt->status[i] == CJ_OK) dacc += cj_align16((size_t)total[i]);
    }
    int rc;
    tr.mark("host walk");
    if ((rc = upload_units(c, bt, where, sbase))) return rc;
    tr.mark("upload");
    if ((rc = c->f_ddst.ensure(dacc + 64))) return rc;
    Items comp, stored, all;
    std::vector<uint32_t> want_crc, owner;      // per entry of `all`
    std::vector<int64_t> comp_idx;              // per entry of `all`: index into `comp`, or -1 for a stored chunk
    for (size_t i = 0; i < n; i++) {
        if (bt->status[i] != CJ_OK) continue;
        uint64_t d = dbase[i];
        for (const SnChunk& ch : chunks[i]) {
            if (ch.compressed) { comp_idx.push_back((int64_t)comp.size()); comp.add(sbase[i] + ch.body_off, ch.body_len, d, ch.ulen); }
            else { comp_idx.push_back(-1); stored.add(sbase[i] + ch.body_off, ch.body_len, d, ch.ulen); }
            all.add(d, ch.ulen, 0, 0);
            want_crc.push_back(ch.crc);
            owner.push_back((uint32_t)i);
            d += ch.ulen;
        }
    }
    // compressed chunks -> raw block decoder; stored chunks -> device copy; then CRC of every chunk
    Scratch &dd = c->f_ddesc, &hd = c->f_hdesc;
    // three descriptor groups share the scratch: carve them out of one allocation
    const size_t per = 48 + 8;
    const size_t need = (comp.size() + stored.size() + all.size()) * per + 3 * 64;
    if ((rc = dd.ensure(need))) return rc;
    if ((rc = hd.ensure(need))) return rc;
    auto carve = [&](const Items& it, size_t byte_off, DevItems* out) -> int {
        Scratch ds, hs2;
        ds.p = (uint8_t*)dd.p + byte_off; ds.cap = dd.cap - byte_off;
        hs2.p = (uint8_t*)hd.p + byte_off; hs2.cap = hd.cap - byte_off; hs2.pinned = true;
        int r = upload_items(c, it, ds, hs2, out);
        ds.p = nullptr; hs2.p = nullptr;  // not owned
        return r;
    };
    DevItems dcomp, dstored, dall;
    size_t o0 = 0, o1 = cj_align16(comp.size() * per + 16), o2 = o1 + cj_align16(stored.size() * per + 16);
    if ((rc = carve(comp, o0, &dcomp))) return rc;
    if ((rc = carve(stored, o1, &dstored))) return rc;
    if ((rc = carve(all, o2, &dall))) return rc;
    if (comp.size()) {
        Batch b;
        b.n = (uint32_t)comp.size();
        b.src_base = (const uint8_t*)c->f_dsrc.p; b.src_off = dcomp.so; b.src_len = dcomp.sl;
        b.dst_base = (uint8_t*)c->f_ddst.p; b.dst_off = dcomp.dof; b.dst_cap = dcomp.dc; b.dst_len = dcomp.dl; b.status = dcomp.st;
        if ((rc = cj_run_device_batch(c, CJ_SNAPPY_RAW, false, b, nullptr))) return rc;
    }
    if ((rc = copy_units(c, (uint32_t)stored.size(), (const uint8_t*)c->f_dsrc.p, dstored.so, dstored.sl, (uint8_t*)c->f_ddst.p, dstored.dof))) return rc;
    if ((rc = launch_crc(c, (uint32_t)all.size(), (const uint8_t*)c->f_ddst.p, dall.so, dall.sl, dall.aux))) return rc;
    if ((rc = fetch_results(c, dcomp))) return rc;
    if ((rc = fetch_results(c, dall))) return rc;
    tr.mark("decode + crc");
    // fold chunk results into unit results (first failure in stream order wins)
    {
        const uint64_t* cdl = dcomp.h + 4 * dcomp.n;
        const int32_t* cst = (const int32_t*)(dcomp.h + 5 * dcomp.n);
        const uint32_t* got = (const uint32_t*)((const int32_t*)(dall.h + 5 * dall.n) + dall.n);
        std::vector<int32_t> ust(n, CJ_OK);
        for (size_t k = 0; k < all.size(); k++) {
            const uint32_t u = owner[k];
            int32_t st = CJ_OK;
            if (comp_idx[k] >= 0) {
                const size_t ci = (size_t)comp_idx[k];
                st = cst[ci];
                if (st == CJ_OK && cdl[ci] != all.sl[k]) st = CJ_ST_LEN_MISMATCH;
            }
            if (st == CJ_OK && got[k] != want_crc[k]) st = CJ_ST_CHECKSUM;
            if (ust[u] == CJ_OK && st != CJ_OK) ust[u] = st;
        }
        for (size_t i = 0; i < n; i++)
            if (bt->status[i] == CJ_OK) {
                bt->status[i] = ust[i];
                bt->dst_len[i] = ust[i] == CJ_OK ? total[i] : 0;
            }
    }
    tr.mark("fold");
    rc = download_units(c, bt, where, dbase, dacc);
    tr.mark("download");
    return rc;
}

int snappy_framed_compress(cj_ctx* c, int where, const cj_batch* bt) {
    PhaseTrace tr(c, "snappy_framed_compress");
    const size_t n = bt->n;
    static const uint8_t STREAM_ID[10] = {0xff, 0x06, 0x00, 0x00, 's', 'N', 'a', 'P', 'p', 'Y'};
    std::vector<uint64_t> sbase;
    int rc;
    if ((rc = upload_units(c, bt, where, sbase))) return rc;
    const size_t slot = cj_align16(32 + 65536 + 65536 / 6);
    Items ch;  // one item per 64 KiB chunk: raw chunk -> slot
    std::vector<uint32_t> owner;
    for (size_t i = 0; i < n; i++) {
        bt->status[i] = CJ_OK;
        bt->dst_len[i] = 0;
        const uint64_t piece = frame_piece(bt->src_len[i]);
        for (uint64_t p = 0; p < bt->src_len[i]; p += piece) {
            const uint64_t L = std::min<uint64_t>(piece, bt->src_len[i] - p);
            ch.add(sbase[i] + p, L, ch.size() * slot, slot);
            owner.push_back((uint32_t)i);
        }
    }
    const size_t nc = ch.size();
    if ((rc = c->f_dtmp.ensure(nc * slot + 64))) return rc;
    DevItems dch;
    if ((rc = upload_items(c, ch, c->f_ddesc, c->f_hdesc, &dch))) return rc;
    if (nc) {
        Batch b;
        b.n = (uint32_t)nc;
        b.src_base = (const uint8_t*)c->f_dsrc.p; b.src_off = dch.so; b.src_len = dch.sl;
        b.dst_base = (uint8_t*)c->f_dtmp.p; b.dst_off = dch.dof; b.dst_cap = dch.dc; b.dst_len = dch.dl; b.status = dch.st;
        if ((rc = cj_run_device_batch(c, CJ_SNAPPY_RAW, true, b, nullptr))) return rc;
        if ((rc = launch_crc(c, (uint32_t)nc, (const uint8_t*)c->f_dsrc.p, dch.so, dch.sl, dch.aux))) return rc;
    }
    if ((rc = fetch_results(c, dch))) return rc;
    const uint64_t* clen = dch.h + 4 * nc;
    const int32_t* cst = (const int32_t*)(dch.h + 5 * nc);
    const uint32_t* crc = (const uint32_t*)(cst + nc);
    // layout of every output stream; header bytes are assembled on the host (framing only), bodies are spliced on the device
    std::vector<uint64_t> dbase(n), total(n, 0);
    size_t dacc = 0;
    Items body_comp, body_raw, hdr;
    std::vector<uint8_t> hdr_bytes;
    hdr_bytes.reserve(10 * n + 8 * nc + 64);
    size_t k = 0;
    for (size_t i = 0; i < n; i++) {
        dbase[i] = dacc;
        if (bt->status[i] != CJ_OK) continue;
        uint64_t pos = 10;
        hdr.add(hdr_bytes.size(), 10, dacc, 0);
        hdr_bytes.insert(hdr_bytes.end(), STREAM_ID, STREAM_ID + 10);
        const uint64_t piece = frame_piece(bt->src_len[i]);
        for (uint64_t p = 0; p < bt->src_len[i]; p += piece, k++) {
            const uint64_t L = ch.sl[k];
            if (cst[k] != CJ_OK) { bt->status[i] = cst[k]; }
            const bool use_comp = clen[k] < L - L / 8;  // snap: keep the compressed form only if it saves >= 12.5 %
            const uint64_t body = use_comp ? clen[k] : L;
            uint8_t h8[8];
            h8[0] = use_comp ? 0x00 : 0x01;
            const uint64_t cl = body + 4;
            h8[1] = (uint8_t)cl; h8[2] = (uint8_t)(cl >> 8); h8[3] = (uint8_t)(cl >> 16);
            h_wr32(h8 + 4, crc[k]);
            hdr.add(hdr_bytes.size(), 8, dacc + pos, 0);
            hdr_bytes.insert(hdr_bytes.end(), h8, h8 + 8);
            if (use_comp) body_comp.add(ch.dof[k], body, dacc + pos + 8, 0);
            else body_raw.add(ch.so[k], body, dacc + pos + 8, 0);
            pos += 8 + body;
        }
        total[i] = pos;
        if (bt->status[i] == CJ_OK && pos > bt->dst_cap[i]) bt->status[i] = CJ_ST_DST_SMALL;
        dacc += cj_align16((size_t)pos);
    }
    if ((rc = c->f_ddst.ensure(dacc + hdr_bytes.size() + 128))) return rc;
    // header blob rides at the end of the destination arena
    const size_t blob_off = dacc;
    if (!hdr_bytes.empty()) {
        if ((rc = c->f_hdst.ensure(hdr_bytes.size() + 64))) return rc;
        memcpy(c->f_hdst.p, hdr_bytes.data(), hdr_bytes.size());
        CUDA_TRY(cudaMemcpyAsync((uint8_t*)c->f_ddst.p + blob_off, c->f_hdst.p, hdr_bytes.size(), cudaMemcpyHostToDevice, c->stream));
    }
    auto splice = [&](const Items& it, const uint8_t* sb, uint64_t sadd) -> int {
        if (!it.size()) return CJ_OK;
        Items t = it;
        for (auto& v : t.so) v += sadd;
        DevItems d;
        int r = upload_items(c, t, c->f_ddesc, c->f_hdesc, &d);
        if (r) return r;
        r = copy_units(c, (uint32_t)t.size(), sb, d.so, d.sl, (uint8_t*)c->f_ddst.p, d.dof);
        if (r) return r;
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // descriptor scratch is reused by the next splice
        return CJ_OK;
    };
    tr.mark("host layout");
    if ((rc = splice(hdr, (const uint8_t*)c->f_ddst.p, blob_off))) return rc;
    if ((rc = splice(body_comp, (const uint8_t*)c->f_dtmp.p, 0))) return rc;
    if ((rc = splice(body_raw, (const uint8_t*)c->f_dsrc.p, 0))) return rc;
    tr.mark("splice");
    for (size_t i = 0; i < n; i++) bt->dst_len[i] = bt->status[i] == CJ_OK ? total[i] : 0;
    rc = download_units(c, bt, where, dbase, dacc);
    tr.mark("download");
    return rc;
}

// ---- LZ4 frame compress: independent 64 KiB blocks + content checksum --------------------------------
int lz4f_compress(cj_ctx* c, int where, const cj_batch* bt, const cj_params* params) {
    // level >= 3 is LZ4HC in the reference (src/lz4.rs:17 default 4): the block encoder runs with its largest match table
    cj_params blk{params ? params->level : 4, 1, 0};
    const size_t n = bt->n;
    std::vector<uint64_t> sbase;
    int rc;
    PhaseTrace tr(c, "lz4f_compress");
    // content checksums of large inputs: on host threads, next to the upload and the block encoder (see HOST_HASH_MIN)
    std::vector<std::future<uint32_t>> host_hash(n);
    {
        const uint8_t* hs = (const uint8_t*)bt->src_base;
        for (size_t i = 0; i < n; i++)
            if (bt->src_len[i] >= HOST_HASH_MIN) {
                const uint8_t* p = hs + bt->src_off[i];
                const size_t len = (size_t)bt->src_len[i];
                host_hash[i] = std::async(std::launch::async, [p, len]() { return h_xxh32_words(p, len); });
            }
    }
    if ((rc = upload_units(c, bt, where, sbase))) return rc;
    tr.mark("upload");
    const size_t slot = cj_align16(65536 + 65536 / 255 + 16);
    Items ch, whole;
    for (size_t i = 0; i < n; i++) {
        bt->status[i] = CJ_OK;
        bt->dst_len[i] = 0;
        whole.add(sbase[i], host_hash[i].valid() ? 0 : bt->src_len[i], 0, 0);
        const uint64_t piece = frame_piece(bt->src_len[i]);
        for (uint64_t p = 0; p < bt->src_len[i]; p += piece) ch.add(sbase[i] + p, std::min<uint64_t>(piece, bt->src_len[i] - p), ch.size() * slot, slot);
    }
    const size_t nc = ch.size();
    if ((rc = c->f_dtmp.ensure(nc * slot + 64))) return rc;
    // content checksums (one thread per frame) while the block encoder runs
    DevItems dwhole;
    Scratch& dd = c->f_ddesc; Scratch& hd = c->f_hdesc;
    const size_t per = 48 + 8;
    const size_t need = (nc + n) * per + 2 * 64;
    if ((rc = dd.ensure(need))) return rc;
    if ((rc = hd.ensure(need))) return rc;
    auto carve = [&](const Items& it, size_t byte_off, DevItems* out) -> int {
        Scratch ds, hs2;
        ds.p = (uint8_t*)dd.p + byte_off; ds.cap = dd.cap - byte_off;
        hs2.p = (uint8_t*)hd.p + byte_off; hs2.cap = hd.cap - byte_off; hs2.pinned = true;
        int r = upload_items(c, it, ds, hs2, out);
        ds.p = nullptr; hs2.p = nullptr;
        return r;
    };
    DevItems dch;
    if ((rc = carve(ch, 0, &dch))) return rc;
    if ((rc = carve(whole, cj_align16(nc * per + 16), &dwhole))) return rc;
    if (nc) {
        Batch b;
        b.n = (uint32_t)nc;
        b.src_base = (const uint8_t*)c->f_dsrc.p; b.src_off = dch.so; b.src_len = dch.sl;
        b.dst_base = (uint8_t*)c->f_dtmp.p; b.dst_off = dch.dof; b.dst_cap = dch.dc; b.dst_len = dch.dl; b.status = dch.st;
        if ((rc = cj_run_device_batch(c, CJ_LZ4_BLOCK, true, b, &blk))) return rc;
    }
    tr.mark("items + block encode");
    if ((rc = launch_xxh32(c, (uint32_t)n, (const uint8_t*)c->f_dsrc.p, dwhole.so, dwhole.sl, dwhole.aux))) return rc;
    tr.mark("content xxh32");
    if ((rc = fetch_results(c, dch))) return rc;
    if ((rc = fetch_results(c, dwhole))) return rc;
    const uint64_t* clen = dch.h + 4 * nc;
    const int32_t* cst = (const int32_t*)(dch.h + 5 * nc);
    const uint32_t* content_xxh = (const uint32_t*)((const int32_t*)(dwhole.h + 5 * n) + n);
    std::vector<uint64_t> dbase(n), total(n, 0);
    Items body_comp, body_raw, hdr;
    std::vector<uint8_t> hb;
    size_t dacc = 0, k = 0;
    for (size_t i = 0; i < n; i++) {
        dbase[i] = dacc;
        uint8_t fh[15];
        h_wr32(fh, 0x184D2204u);
        fh[4] = 0x40 | 0x20 | 0x08 | 0x04;  // version 01, independent blocks, content size, content checksum
        fh[5] = 0x40;                       // 64 KiB blocks
        const uint64_t csize = bt->src_len[i];
        memcpy(fh + 6, &csize, 8);
        fh[14] = (uint8_t)(h_xxh32(fh + 4, 10) >> 8);
        hdr.add(hb.size(), 15, dacc, 0);
        hb.insert(hb.end(), fh, fh + 15);
        uint64_t pos = 15;
        const uint64_t piece = frame_piece(bt->src_len[i]);
        for (uint64_t p = 0; p < bt->src_len[i]; p += piece, k++) {
            const uint64_t L = ch.sl[k];
            if (cst[k] != CJ_OK) bt->status[i] = cst[k];
            const bool use_comp = clen[k] < L;
            const uint64_t body = use_comp ? clen[k] : L;
            uint8_t b4[4];
            h_wr32(b4, (uint32_t)body | (use_comp ? 0u : 0x80000000u));
            hdr.add(hb.size(), 4, dacc + pos, 0);
            hb.insert(hb.end(), b4, b4 + 4);
            if (use_comp) body_comp.add(ch.dof[k], body, dacc + pos + 4, 0);
            else body_raw.add(ch.so[k], body, dacc + pos + 4, 0);
            pos += 4 + body;
        }
        uint8_t tail[8];
        h_wr32(tail, 0);
        h_wr32(tail + 4, host_hash[i].valid() ? host_hash[i].get() : content_xxh[i]);
        hdr.add(hb.size(), 8, dacc + pos, 0);
        hb.insert(hb.end(), tail, tail + 8);
        pos += 8;
        total[i] = pos;
        if (bt->status[i] == CJ_OK && pos > bt->dst_cap[i]) bt->status[i] = CJ_ST_DST_SMALL;
        dacc += cj_align16((size_t)pos);
    }
    if ((rc = c->f_ddst.ensure(dacc + hb.size() + 128))) return rc;
    const size_t blob_off = dacc;
    if ((rc = c->f_hdst.ensure(hb.size() + 64))) return rc;
    memcpy(c->f_hdst.p, hb.data(), hb.size());
    CUDA_TRY(cudaMemcpyAsync((uint8_t*)c->f_ddst.p + blob_off, c->f_hdst.p, hb.size(), cudaMemcpyHostToDevice, c->stream));
    auto splice = [&](const Items& it, const uint8_t* sb, uint64_t sadd) -> int {
        if (!it.size()) return CJ_OK;
        Items t = it;
        for (auto& v : t.so) v += sadd;
        DevItems d;
        int r = upload_items(c, t, c->f_ddesc, c->f_hdesc, &d);
        if (r) return r;
        r = copy_units(c, (uint32_t)t.size(), sb, d.so, d.sl, (uint8_t*)c->f_ddst.p, d.dof);
        if (r) return r;
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return CJ_OK;
    };
    tr.mark("host layout");
    if ((rc = splice(hdr, (const uint8_t*)c->f_ddst.p, blob_off))) return rc;
    if ((rc = splice(body_comp, (const uint8_t*)c->f_dtmp.p, 0))) return rc;
    if ((rc = splice(body_raw, (const uint8_t*)c->f_dsrc.p, 0))) return rc;
    tr.mark("splice");
    for (size_t i = 0; i < n; i++) bt->dst_len[i] = bt->status[i] == CJ_OK ? total[i] : 0;
    rc = download_units(c, bt, where, dbase, dacc);
    tr.mark("download");
    return rc;
}

// ---- LZ4 frame decompress: block-parallel for independent-block frames --------------------------
// The warp-per-frame kernel (lz4f_decode_kernel) decodes one frame's blocks one after the other — right for batches
// of many frames and the only option for linked blocks, but a single large frame would run on one warp (62 MB/s).
// Frames whose descriptor says "independent blocks" (what this engine's own encoder and the lz4 CLI write) are cut
// into block units on the host (headers only), decoded by the batched block kernels into fixed slots, spliced into
// place with a device copy, and their block / content checksums computed by xxh32_units_kernel.  A unit with anything
// unusual in it (a failed block, a checksum or size mismatch, a malformed or linked frame) is handed to the
// warp-per-frame kernel, which owns the exact status codes.
struct Lz4fBlock { uint64_t payload_off; uint32_t len; bool stored; bool has_sum; uint32_t want_sum; };
struct Lz4fFrame { uint32_t bmax; bool has_csum, has_csize; uint32_t want_csum; uint64_t csize; size_t first_block, n_blocks; };

// Returns true if every frame of the stream is a well-formed independent-block frame (skippable frames are stepped over).
bool lz4f_plan_host(const uint8_t* s, size_t n, std::vector<Lz4fFrame>& frames, std::vector<Lz4fBlock>& blocks) {
    size_t ip = 0;
    while (ip < n) {
        if (n - ip < 4) return false;
        const uint32_t magic = h_rd32(s + ip);
        if (magic >= 0x184D2A50u && magic <= 0x184D2A5Fu) {
            if (n - ip < 8) return false;
            const uint32_t sz = h_rd32(s + ip + 4);
            if (sz > n - ip - 8) return false;
            ip += 8 + (size_t)sz;
            continue;
        }
        if (magic != 0x184D2204u || n - ip < 7) return false;
        const uint32_t flg = s[ip + 4], bd = s[ip + 5];
        if ((flg >> 6) != 1 || (flg & 0x02) || (bd & 0x8F)) return false;
        const bool indep = (flg >> 5) & 1, bsum = (flg >> 4) & 1, csize = (flg >> 3) & 1, csum = (flg >> 2) & 1, dict = flg & 1;
        const uint32_t bid = (bd >> 4) & 7;
        if (!indep || dict || bid < 4) return false;
        const uint32_t dlen = 2 + (csize ? 8 : 0);
        if (n - ip < 4 + dlen + 1) return false;
        if (((h_xxh32(s + ip + 4, dlen) >> 8) & 0xff) != s[ip + 4 + dlen]) return false;
        Lz4fFrame f{};
        f.bmax = 1u << (8 + 2 * bid);
        f.has_csum = csum; f.has_csize = csize;
        if (csize) memcpy(&f.csize, s + ip + 6, 8);
        f.first_block = blocks.size();
        ip += 4 + dlen + 1;
        for (;;) {
            if (n - ip < 4) return false;
            const uint32_t bs = h_rd32(s + ip);
            ip += 4;
            if (bs == 0) break;
            const uint32_t blen = bs & 0x7FFFFFFFu;
            if (blen > f.bmax || blen > n - ip || blen == 0) return false;
            Lz4fBlock b{ip, blen, (bs >> 31) != 0, bsum, 0};
            if (bsum) {
                if (n - ip - blen < 4) return false;
                b.want_sum = h_rd32(s + ip + blen);
            }
            blocks.push_back(b);
            ip += (size_t)blen + (bsum ? 4 : 0);
        }
        f.n_blocks = blocks.size() - f.first_block;
        if (csum) {
            if (n - ip < 4) return false;
            f.want_csum = h_rd32(s + ip);
            ip += 4;
        }
        frames.push_back(f);
    }
    return true;
}

// Sequentially carves descriptor groups out of ctx->f_ddesc / f_hdesc (sized up front).
struct DescCarver {
    cj_ctx* c;
    size_t off = 0;
    static size_t bytes_for(size_t n) { return cj_align16(n * (48 + 8) + 64); }
    int put(const Items& it, DevItems* out) {
        Scratch ds, hs2;
        ds.p = (uint8_t*)c->f_ddesc.p + off; ds.cap = c->f_ddesc.cap - off;
        hs2.p = (uint8_t*)c->f_hdesc.p + off; hs2.cap = c->f_hdesc.cap - off; hs2.pinned = true;
        const int r = upload_items(c, it, ds, hs2, out);
        ds.p = nullptr; hs2.p = nullptr;  // not owned
        off += bytes_for(it.size());
        return r;
    }
};

int lz4f_decompress(cj_ctx* c, int where, const cj_batch* bt) {
    const size_t n = bt->n;
    const uint8_t* hs = (const uint8_t*)bt->src_base;
    std::vector<uint64_t> sbase, dbase(n);
    int rc;
    PhaseTrace tr(c, "lz4f_decompress");
    if ((rc = upload_units(c, bt, where, sbase))) return rc;
    tr.mark("upload");
    size_t dacc = 0;
    for (size_t i = 0; i < n; i++) { dbase[i] = dacc; dacc += cj_align16((size_t)bt->dst_cap[i]); bt->dst_len[i] = 0; bt->status[i] = CJ_OK; }
    if ((rc = c->f_ddst.ensure(dacc + 64))) return rc;

    // plan: which units go block-parallel
    std::vector<std::vector<Lz4fFrame>> frames(n);
    std::vector<std::vector<Lz4fBlock>> blocks(n);
    std::vector<char> par(n, 0);
    size_t nblk = 0, nfr = 0, slot_bytes = 0;
    for (size_t i = 0; i < n; i++) {
        if (bt->src_len[i] > MAX_UNIT || bt->dst_cap[i] > MAX_UNIT) continue;
        if (bt->src_len[i] < (1u << 17)) continue;   // small streams: one warp is as good
        if (!lz4f_plan_host(hs + bt->src_off[i], (size_t)bt->src_len[i], frames[i], blocks[i]) || blocks[i].size() < 2) continue;
        // Scratch is sized from untrusted block headers: a block can produce at most min(bmax, 255 x its length), and a
        // stream whose slots would far exceed what the caller can take (many tiny blocks under a 4 MiB bmax) is decoded
        // serially like the reference does, instead of failing with CJ_E_NOMEM.
        size_t want = 0;
        for (const Lz4fFrame& f : frames[i])
            for (size_t k = f.first_block; k < f.first_block + f.n_blocks; k++)
                if (!blocks[i][k].stored) want += cj_align16(std::min<size_t>(f.bmax, (size_t)blocks[i][k].len * 255));
        if (want > 16 * (size_t)bt->dst_cap[i] + ((size_t)64 << 20)) continue;
        par[i] = 1;
        nblk += blocks[i].size();
        nfr += frames[i].size();
    }
    Items dec, sums;          // compressed blocks -> temp slots; block checksums over compressed payloads
    std::vector<size_t> dec_of(0), sum_of(0);   // per block (flattened over parallel units): index into dec / sums or ~0
    for (size_t i = 0; i < n; i++) {
        if (!par[i]) continue;
        for (const Lz4fFrame& f : frames[i])
            for (size_t k = f.first_block; k < f.first_block + f.n_blocks; k++) {
                const Lz4fBlock& b = blocks[i][k];
                if (!b.stored) {
                    const size_t bslot = std::min<size_t>(f.bmax, (size_t)b.len * 255);
                    dec_of.push_back(dec.size());
                    dec.add(sbase[i] + b.payload_off, b.len, slot_bytes, bslot);
                    slot_bytes += cj_align16(bslot);
                }
                else dec_of.push_back(~(size_t)0);
                if (b.has_sum) { sum_of.push_back(sums.size()); sums.add(sbase[i] + b.payload_off, b.len, 0, 0); }
                else sum_of.push_back(~(size_t)0);
            }
    }
    struct LateSum { size_t unit, off, len; uint32_t want; };
    std::vector<LateSum> late;   // content checksums verified on the host while the output is copied home (HOST_HASH_MIN)
    std::vector<char> serial(n, 0);
    for (size_t i = 0; i < n; i++) serial[i] = !par[i];
    if (nblk) {
        if ((rc = c->f_dtmp.ensure(slot_bytes + 64))) return rc;
        const size_t need = DescCarver::bytes_for(dec.size()) + DescCarver::bytes_for(sums.size()) + 2 * DescCarver::bytes_for(nblk) + DescCarver::bytes_for(nfr) + 256;
        if ((rc = c->f_ddesc.ensure(need))) return rc;
        if ((rc = c->f_hdesc.ensure(need))) return rc;
        DescCarver carve{c};
        DevItems ddec, dsums;
        if ((rc = carve.put(dec, &ddec))) return rc;
        if ((rc = carve.put(sums, &dsums))) return rc;
        if (dec.size()) {
            Batch b;
            b.n = (uint32_t)dec.size();
            b.src_base = (const uint8_t*)c->f_dsrc.p; b.src_off = ddec.so; b.src_len = ddec.sl;
            b.dst_base = (uint8_t*)c->f_dtmp.p; b.dst_off = ddec.dof; b.dst_cap = ddec.dc; b.dst_len = ddec.dl; b.status = ddec.st;
            if ((rc = cj_run_device_batch(c, CJ_LZ4_BLOCK, false, b, nullptr))) return rc;
        }
        if ((rc = launch_xxh32(c, (uint32_t)sums.size(), (const uint8_t*)c->f_dsrc.p, dsums.so, dsums.sl, dsums.aux))) return rc;
        if ((rc = fetch_results(c, ddec))) return rc;
        if ((rc = fetch_results(c, dsums))) return rc;
        tr.mark("plan + block decode");
        const uint64_t* bdl = ddec.h + 4 * ddec.n;
        const int32_t* bst = (const int32_t*)(ddec.h + 5 * ddec.n);
        const uint32_t* bsum = (const uint32_t*)((const int32_t*)(dsums.h + 5 * dsums.n) + dsums.n);
        // place every block; anything unusual sends the whole unit to the warp-per-frame kernel
        Items mv_dec, mv_raw, whole;      // temp slot -> final, stored payload -> final, frames to checksum
        std::vector<uint32_t> whole_want, whole_owner;
        late.clear();
        size_t flat = 0;
        for (size_t i = 0; i < n; i++) {
            if (!par[i]) continue;
            uint64_t pos = 0;
            bool ok = true;
            const size_t m0 = mv_dec.size(), r0 = mv_raw.size(), w0 = whole.size(), l0 = late.size();
            for (const Lz4fFrame& f : frames[i]) {
                const uint64_t fstart = pos;
                for (size_t k = f.first_block; k < f.first_block + f.n_blocks; k++, flat++) {
                    const Lz4fBlock& b = blocks[i][k];
                    if (!ok) continue;
                    if (b.has_sum && bsum[sum_of[flat]] != b.want_sum) { ok = false; continue; }
                    uint64_t produced;
                    if (b.stored) {
                        produced = b.len;
                        if (pos + produced > bt->dst_cap[i]) { ok = false; continue; }
                        mv_raw.add(sbase[i] + b.payload_off, produced, dbase[i] + pos, 0);
                    } else {
                        const size_t d = dec_of[flat];
                        if (bst[d] != CJ_OK) { ok = false; continue; }
                        produced = bdl[d];
                        if (pos + produced > bt->dst_cap[i]) { ok = false; continue; }
                        mv_dec.add(dec.dof[d], produced, dbase[i] + pos, 0);
                    }
                    pos += produced;
                }
                if (ok && f.has_csize && f.csize != pos - fstart) ok = false;
                if (ok && f.has_csum) {
                    if (where != CJ_DEVICE && pos - fstart >= HOST_HASH_MIN) late.push_back({i, (size_t)fstart, (size_t)(pos - fstart), f.want_csum});
                    else { whole.add(dbase[i] + fstart, pos - fstart, 0, 0); whole_want.push_back(f.want_csum); whole_owner.push_back((uint32_t)i); }
                }
            }
            if (!ok) {
                serial[i] = 1;
                mv_dec.so.resize(m0); mv_dec.sl.resize(m0); mv_dec.dof.resize(m0); mv_dec.dc.resize(m0);
                mv_raw.so.resize(r0); mv_raw.sl.resize(r0); mv_raw.dof.resize(r0); mv_raw.dc.resize(r0);
                whole.so.resize(w0); whole.sl.resize(w0); whole.dof.resize(w0); whole.dc.resize(w0);
                whole_want.resize(w0); whole_owner.resize(w0);
                late.resize(l0);
            } else {
                bt->dst_len[i] = pos;
            }
        }
        DevItems dmv, draw, dwhole;
        if ((rc = carve.put(mv_dec, &dmv))) return rc;
        if ((rc = carve.put(mv_raw, &draw))) return rc;
        if ((rc = carve.put(whole, &dwhole))) return rc;
        if ((rc = copy_units(c, (uint32_t)mv_dec.size(), (const uint8_t*)c->f_dtmp.p, dmv.so, dmv.sl, (uint8_t*)c->f_ddst.p, dmv.dof))) return rc;
        if ((rc = copy_units(c, (uint32_t)mv_raw.size(), (const uint8_t*)c->f_dsrc.p, draw.so, draw.sl, (uint8_t*)c->f_ddst.p, draw.dof))) return rc;
        tr.mark("place blocks");
        if ((rc = launch_xxh32(c, (uint32_t)whole.size(), (const uint8_t*)c->f_ddst.p, dwhole.so, dwhole.sl, dwhole.aux))) return rc;
        if ((rc = fetch_results(c, dwhole))) return rc;
        tr.mark("content xxh32");
        const uint32_t* got = (const uint32_t*)((const int32_t*)(dwhole.h + 5 * dwhole.n) + dwhole.n);
        for (size_t k = 0; k < whole.size(); k++)
            if (got[k] != whole_want[k]) { serial[whole_owner[k]] = 1; bt->dst_len[whole_owner[k]] = 0; }
    }
    // everything else (and every unit that tripped a check above): one warp per frame stream, exact status codes
    Items ser;
    std::vector<size_t> ser_unit;
    for (size_t i = 0; i < n; i++)
        if (serial[i]) { ser.add(sbase[i], bt->src_len[i], dbase[i], bt->dst_cap[i]); ser_unit.push_back(i); }
    if (ser.size()) {
        const size_t need = DescCarver::bytes_for(ser.size()) + 64;
        // results of the parallel part are already on the host: the descriptor scratch can be reused
        if ((rc = c->f_ddesc.ensure(need))) return rc;
        if ((rc = c->f_hdesc.ensure(need))) return rc;
        DescCarver carve{c};
        DevItems dser;
        if ((rc = carve.put(ser, &dser))) return rc;
        Batch b;
        b.n = (uint32_t)ser.size();
        b.src_base = (const uint8_t*)c->f_dsrc.p; b.src_off = dser.so; b.src_len = dser.sl;
        b.dst_base = (uint8_t*)c->f_ddst.p; b.dst_off = dser.dof; b.dst_cap = dser.dc; b.dst_len = dser.dl; b.status = dser.st;
        if ((rc = cj_run_device_batch(c, CJ_LZ4_FRAME, false, b, nullptr))) return rc;
        if ((rc = fetch_results(c, dser))) return rc;
        const uint64_t* dl = dser.h + 4 * dser.n;
        const int32_t* st = (const int32_t*)(dser.h + 5 * dser.n);
        for (size_t k = 0; k < ser.size(); k++) { bt->dst_len[ser_unit[k]] = st[k] == CJ_OK ? dl[k] : 0; bt->status[ser_unit[k]] = st[k]; }
    }
    tr.mark("serial units");
    std::vector<char> bad(n, 0);
    std::function<void(const std::function<const uint8_t*(size_t)>&)> verify;
    if (!late.empty())
        verify = [&](const std::function<const uint8_t*(size_t)>& unit_bytes) {
            for (const LateSum& L : late)
                if (!serial[L.unit] && bt->status[L.unit] == CJ_OK && h_xxh32_words(unit_bytes(L.unit) + L.off, L.len) != L.want) bad[L.unit] = 1;
        };
    rc = download_units(c, bt, where, dbase, dacc, verify);
    for (size_t i = 0; i < n; i++)
        if (bad[i]) { bt->status[i] = CJ_ST_CHECKSUM; bt->dst_len[i] = 0; }
    tr.mark("download");
    return rc;
}

// ---- Zstandard through the single-buffer API ---------------------------------------------------
// The zstd kernels run one warp per frame, which is the right unit for batches of frames but leaves a single large
// buffer on one warp.  Compress: inputs larger than ZS_PIECE are written as a sequence of independent frames of
// ZS_PIECE bytes each (concatenated frames are one valid Zstandard stream — zstd::stream::read::Decoder, which the
// reference's decompress uses (src/zstd.rs:23-28), and ZSTD_decompress read them back to back; the encoder's
// match window is 64 KiB, so the split costs almost no ratio), compressed by as many warps.  Decompress: the host
// walks the frame headers; a stream of two or more frames that all declare their content size is decoded one frame
// per warp straight into its final position.  Anything else, and any stream whose frames do not all come back
// clean and exactly sized, goes through the whole-stream path, which owns the exact status codes.
constexpr size_t ZS_PIECE = 512 * 1024;      // piece size for large inputs
constexpr size_t ZS_PIECE_MIN = 64 * 1024;   // small inputs are cut finer (latency is one warp's time per piece): len / 64, clamped

int zstd_compress_split(cj_ctx* c, int where, const cj_batch* bt, const cj_params* params) {
    const size_t n = bt->n;
    PhaseTrace tr(c, "zstd_compress");
    std::vector<uint64_t> sbase, dbase(n);
    int rc;
    if ((rc = upload_units(c, bt, where, sbase))) return rc;
    tr.mark("upload");
    Items enc;
    std::vector<size_t> first(n + 1, 0);
    size_t slot_acc = 0;
    for (size_t i = 0; i < n; i++) {
        first[i] = enc.size();
        const uint64_t L = bt->src_len[i];
        const uint64_t piece = std::min<uint64_t>(ZS_PIECE, std::max<uint64_t>(ZS_PIECE_MIN, (L / 64 + 65535) & ~(uint64_t)65535));
        uint64_t p = 0;
        do {
            const uint64_t len = std::min<uint64_t>(piece, L - p);
            const size_t cap = cj_align16(cj_compress_bound(CJ_ZSTD, (size_t)len));
            enc.add(sbase[i] + p, len, slot_acc, cap);
            slot_acc += cap;
            p += len;
        } while (p < L);
    }
    first[n] = enc.size();
    if ((rc = c->f_dtmp.ensure(slot_acc + 64))) return rc;
    const size_t need = 2 * DescCarver::bytes_for(enc.size()) + 128;
    if ((rc = c->f_ddesc.ensure(need))) return rc;
    if ((rc = c->f_hdesc.ensure(need))) return rc;
    DescCarver carve{c};
    DevItems denc;
    if ((rc = carve.put(enc, &denc))) return rc;
    Batch b;
    b.n = (uint32_t)enc.size();
    b.src_base = (const uint8_t*)c->f_dsrc.p; b.src_off = denc.so; b.src_len = denc.sl;
    b.dst_base = (uint8_t*)c->f_dtmp.p; b.dst_off = denc.dof; b.dst_cap = denc.dc; b.dst_len = denc.dl; b.status = denc.st;
    if ((rc = cj_run_device_batch(c, CJ_ZSTD, true, b, params))) return rc;
    if ((rc = fetch_results(c, denc))) return rc;
    tr.mark("frames encoded");
    const uint64_t* dl = denc.h + 4 * denc.n;
    const int32_t* st = (const int32_t*)(denc.h + 5 * denc.n);
    Items mv;
    size_t dacc = 0;
    for (size_t i = 0; i < n; i++) {
        dbase[i] = dacc;
        bt->status[i] = CJ_OK;
        uint64_t pos = 0;
        for (size_t k = first[i]; k < first[i + 1]; k++) {
            if (st[k] != CJ_OK && bt->status[i] == CJ_OK) bt->status[i] = st[k];
            pos += dl[k];
        }
        if (bt->status[i] == CJ_OK && pos > bt->dst_cap[i]) bt->status[i] = CJ_ST_DST_SMALL;
        bt->dst_len[i] = bt->status[i] == CJ_OK ? pos : 0;
        if (bt->status[i] != CJ_OK) continue;
        pos = 0;
        for (size_t k = first[i]; k < first[i + 1]; k++) { mv.add(enc.dof[k], dl[k], dacc + pos, 0); pos += dl[k]; }
        dacc += cj_align16((size_t)pos);
    }
    if ((rc = c->f_ddst.ensure(dacc + 64))) return rc;
    DevItems dmv;
    if ((rc = carve.put(mv, &dmv))) return rc;
    if ((rc = copy_units(c, (uint32_t)mv.size(), (const uint8_t*)c->f_dtmp.p, dmv.so, dmv.sl, (uint8_t*)c->f_ddst.p, dmv.dof))) return rc;
    tr.mark("splice");
    rc = download_units(c, bt, where, dbase, dacc);
    tr.mark("download");
    return rc;
}

int zstd_decompress_frames(cj_ctx* c, int where, const cj_batch* bt) {
    const size_t n = bt->n;
    const uint8_t* hs = (const uint8_t*)bt->src_base;
    PhaseTrace tr(c, "zstd_decompress");
    std::vector<uint64_t> sbase, dbase(n);
    int rc;
    if ((rc = upload_units(c, bt, where, sbase))) return rc;
    tr.mark("upload");
    size_t dacc = 0;
    for (size_t i = 0; i < n; i++) { dbase[i] = dacc; dacc += cj_align16((size_t)bt->dst_cap[i]); bt->dst_len[i] = 0; bt->status[i] = CJ_OK; }
    if ((rc = c->f_ddst.ensure(dacc + 64))) return rc;
    // pass 0: frame-parallel where the headers allow it, whole stream otherwise; pass 1: whole stream for every unit that tripped
    std::vector<char> par(n, 0), redo(n, 0);
    std::vector<std::vector<cj_frame_info>> frames(n);
    for (size_t i = 0; i < n; i++) {
        if (bt->src_len[i] > MAX_UNIT || bt->dst_cap[i] > MAX_UNIT) continue;
        size_t tot = 0;
        bool exact = false;
        if (cj_zstd_walk_host(hs + bt->src_off[i], (size_t)bt->src_len[i], &tot, &exact, &frames[i]) != CJ_OK || !exact || tot > bt->dst_cap[i]) continue;
        size_t real = 0;   // frames that are not skippable (exact == true: every one of them declares its content size)
        uint64_t pos = 0;
        bool fits = true;  // never hand a frame more room than the unit has left behind the frames before it
        for (const cj_frame_info& f : frames[i]) {
            uint32_t magic; memcpy(&magic, hs + bt->src_off[i] + f.offset, 4);
            if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) continue;
            real++;
            if (f.content > MAX_UNIT || f.content > bt->dst_cap[i] - pos) { fits = false; break; }
            pos += f.content;
        }
        if (real >= 2 && fits) par[i] = 1;
    }
    for (int pass = 0; pass < 2; pass++) {
        Items it;
        std::vector<uint32_t> owner;
        for (size_t i = 0; i < n; i++) {
            if (pass == 0 && par[i]) {
                uint64_t pos = 0;
                for (const cj_frame_info& f : frames[i]) {
                    uint32_t magic; memcpy(&magic, hs + bt->src_off[i] + f.offset, 4);
                    if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) continue;   // skippable frame: nothing to decode
                    it.add(sbase[i] + f.offset, f.size, dbase[i] + pos, f.content);
                    owner.push_back((uint32_t)i);
                    pos += f.content;
                }
                bt->dst_len[i] = pos;
            } else if ((pass == 0 && !par[i]) || (pass == 1 && redo[i])) {
                it.add(sbase[i], bt->src_len[i], dbase[i], bt->dst_cap[i]);
                owner.push_back((uint32_t)i);
            }
        }
        if (!it.size()) continue;
        const size_t need = DescCarver::bytes_for(it.size()) + 64;
        if ((rc = c->f_ddesc.ensure(need))) return rc;
        if ((rc = c->f_hdesc.ensure(need))) return rc;
        DescCarver carve{c};
        DevItems d;
        if ((rc = carve.put(it, &d))) return rc;
        Batch b;
        b.n = (uint32_t)it.size();
        b.src_base = (const uint8_t*)c->f_dsrc.p; b.src_off = d.so; b.src_len = d.sl;
        b.dst_base = (uint8_t*)c->f_ddst.p; b.dst_off = d.dof; b.dst_cap = d.dc; b.dst_len = d.dl; b.status = d.st;
        if ((rc = cj_run_device_batch(c, CJ_ZSTD, false, b, nullptr))) return rc;
        if ((rc = fetch_results(c, d))) return rc;
        tr.mark(pass == 0 ? "frames decoded" : "whole-stream redo");
        const uint64_t* dl = d.h + 4 * d.n;
        const int32_t* st = (const int32_t*)(d.h + 5 * d.n);
        for (size_t k = 0; k < it.size(); k++) {
            const uint32_t u = owner[k];
            if (pass == 0 && par[u]) {
                if (st[k] != CJ_OK || dl[k] != it.dc[k]) redo[u] = 1;
            } else {
                bt->status[u] = st[k];
                bt->dst_len[u] = st[k] == CJ_OK ? dl[k] : 0;
            }
        }
    }
    rc = download_units(c, bt, where, dbase, dacc);
    tr.mark("download");
    return rc;
}

// ---- Snappy raw compress of one large buffer ----------------------------------------------------
// A raw Snappy block is one element stream under one length preamble, and the block encoder is one warp per block.
// snap itself matches inside 64 KiB sub-blocks with a fresh table each, so a large input is cut the same way here:
// every 64 KiB piece is compressed as its own block by its own warp, the pieces' element streams (their own
// preambles dropped) are spliced behind one preamble for the whole length.  Copies never cross a piece boundary,
// so every offset stays valid.  (decompress_raw of such a block is still one warp: the element chain is serial.)
int snappy_raw_compress_split(cj_ctx* c, int where, const cj_batch* bt) {
    const size_t n = bt->n;
    PhaseTrace tr(c, "snappy_raw_compress");
    std::vector<uint64_t> sbase, dbase(n);
    int rc;
    if ((rc = upload_units(c, bt, where, sbase))) return rc;
    auto varint = [](uint64_t v, uint8_t* out) { int k = 0; do { uint8_t b = v & 0x7f; v >>= 7; if (v) b |= 0x80; out[k++] = b; } while (v); return k; };
    Items enc;
    std::vector<size_t> first(n + 1, 0);
    size_t slot_acc = 0;
    for (size_t i = 0; i < n; i++) {
        first[i] = enc.size();
        const uint64_t L = bt->src_len[i];
        // only units above 128 KiB are cut (the rule run_batch routes by): a smaller unit that shares a batch with a large
        // one is compressed as one block, byte for byte what it is when it comes alone
        const uint64_t piece = L > (128u << 10) ? 65536 : std::max<uint64_t>(L, 1);
        uint64_t p = 0;
        do {
            const uint64_t len = std::min<uint64_t>(piece, L - p);
            const size_t cap = cj_align16(32 + (size_t)len + (size_t)len / 6);
            enc.add(sbase[i] + p, len, slot_acc, cap);
            slot_acc += cap;
            p += len;
        } while (p < L);
    }
    first[n] = enc.size();
    if ((rc = c->f_dtmp.ensure(slot_acc + 64))) return rc;
    const size_t need = 2 * DescCarver::bytes_for(enc.size()) + DescCarver::bytes_for(n) + 192;
    if ((rc = c->f_ddesc.ensure(need))) return rc;
    if ((rc = c->f_hdesc.ensure(need))) return rc;
    DescCarver carve{c};
    DevItems denc;
    if ((rc = carve.put(enc, &denc))) return rc;
    Batch b;
    b.n = (uint32_t)enc.size();
    b.src_base = (const uint8_t*)c->f_dsrc.p; b.src_off = denc.so; b.src_len = denc.sl;
    b.dst_base = (uint8_t*)c->f_dtmp.p; b.dst_off = denc.dof; b.dst_cap = denc.dc; b.dst_len = denc.dl; b.status = denc.st;
    if ((rc = cj_run_device_batch(c, CJ_SNAPPY_RAW, true, b, nullptr))) return rc;
    if ((rc = fetch_results(c, denc))) return rc;
    tr.mark("pieces encoded");
    const uint64_t* dl = denc.h + 4 * denc.n;
    const int32_t* st = (const int32_t*)(denc.h + 5 * denc.n);
    Items mv, hdr;
    std::vector<uint8_t> hb;
    size_t dacc = 0;
    for (size_t i = 0; i < n; i++) {
        dbase[i] = dacc;
        bt->status[i] = CJ_OK;
        uint8_t pre[10];
        const int pk = varint(bt->src_len[i], pre);
        uint64_t pos = pk;
        for (size_t k = first[i]; k < first[i + 1]; k++) {
            if (st[k] != CJ_OK && bt->status[i] == CJ_OK) bt->status[i] = st[k];
            uint8_t tmp[10];
            pos += dl[k] - varint(enc.sl[k], tmp);
        }
        // same rule as the one-warp encoder and snap::raw::Encoder: the output must hold max_compress_len(n)
        if (bt->status[i] == CJ_OK && bt->dst_cap[i] < cj_compress_bound(CJ_SNAPPY_RAW, (size_t)bt->src_len[i])) bt->status[i] = CJ_ST_DST_SMALL;
        bt->dst_len[i] = bt->status[i] == CJ_OK ? pos : 0;
        if (bt->status[i] != CJ_OK) continue;
        hdr.add(hb.size(), pk, dacc, 0);
        hb.insert(hb.end(), pre, pre + pk);
        pos = pk;
        for (size_t k = first[i]; k < first[i + 1]; k++) {
            uint8_t tmp[10];
            const int sk = varint(enc.sl[k], tmp);
            mv.add(enc.dof[k] + sk, dl[k] - sk, dacc + pos, 0);
            pos += dl[k] - sk;
        }
        dacc += cj_align16((size_t)pos);
    }
    if ((rc = c->f_ddst.ensure(dacc + hb.size() + 128))) return rc;
    if ((rc = c->f_hdst.ensure(hb.size() + 64))) return rc;
    memcpy(c->f_hdst.p, hb.data(), hb.size());
    if (hb.size()) CUDA_TRY(cudaMemcpyAsync((uint8_t*)c->f_ddst.p + dacc, c->f_hdst.p, hb.size(), cudaMemcpyHostToDevice, c->stream));
    for (auto& v : hdr.so) v += dacc;
    DevItems dmv, dhdr;
    if ((rc = carve.put(mv, &dmv))) return rc;
    if ((rc = carve.put(hdr, &dhdr))) return rc;
    if ((rc = copy_units(c, (uint32_t)mv.size(), (const uint8_t*)c->f_dtmp.p, dmv.so, dmv.sl, (uint8_t*)c->f_ddst.p, dmv.dof))) return rc;
    if ((rc = copy_units(c, (uint32_t)hdr.size(), (const uint8_t*)c->f_ddst.p, dhdr.so, dhdr.sl, (uint8_t*)c->f_ddst.p, dhdr.dof))) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // f_hdst is reused by download_units
    tr.mark("splice");
    rc = download_units(c, bt, where, dbase, dacc);
    tr.mark("download");
    return rc;
}

}  // namespace

int frames_decompress(cj_ctx* c, int codec, int where, const cj_batch* bt) {
    if (where == CJ_DEVICE) {
        cj_set_error("frame containers (codec %d) take host-visible input: pass CJ_HOST or CJ_PINNED", codec);
        return CJ_E_INVALID_ARG;
    }
    if (bt->n == 0) return CJ_OK;
    if (codec == CJ_SNAPPY_FRAMED) return snappy_framed_decompress(c, where, bt);
    if (codec == CJ_LZ4_FRAME) return lz4f_decompress(c, where, bt);
    if (codec == CJ_ZSTD) return zstd_decompress_frames(c, where, bt);
    cj_set_error("unknown frame codec %d", codec);
    return CJ_E_INVALID_ARG;
}

int frames_compress(cj_ctx* c, int codec, int where, const cj_batch* bt, const cj_params* params) {
    if (where == CJ_DEVICE) {
        cj_set_error("frame containers (codec %d) produce host-visible output: pass CJ_HOST or CJ_PINNED", codec);
        return CJ_E_INVALID_ARG;
    }
    if (bt->n == 0) return CJ_OK;
    if (codec == CJ_SNAPPY_FRAMED) return snappy_framed_compress(c, where, bt);
    if (codec == CJ_LZ4_FRAME) return lz4f_compress(c, where, bt, params);
    if (codec == CJ_ZSTD) return zstd_compress_split(c, where, bt, params);
    if (codec == CJ_SNAPPY_RAW) return snappy_raw_compress_split(c, where, bt);
    cj_set_error("unknown frame codec %d", codec);
    return CJ_E_INVALID_ARG;
}

}  // namespace cj

// ================================================================================================
// Size helpers on HOST memory (header parsing only)
// ================================================================================================
int cj_lz4f_walk_host(const uint8_t* s, size_t n, size_t* out, bool* exact, std::vector<cj_frame_info>* frames) {
    size_t p = 0, tot = 0;
    *exact = true;
    while (p < n) {
        if (n - p < 4) return CJ_ST_TRUNCATED;
        uint32_t magic; memcpy(&magic, s + p, 4);
        const size_t frame_at = p;
        if (magic >= 0x184D2A50u && magic <= 0x184D2A5Fu) {
            if (n - p < 8) return CJ_ST_TRUNCATED;
            uint32_t sz; memcpy(&sz, s + p + 4, 4);
            if (sz > n - p - 8) return CJ_ST_TRUNCATED;
            p += 8 + sz;
            if (frames) frames->push_back({frame_at, p - frame_at, 0, true});
            continue;
        }
        if (magic != 0x184D2204u) return CJ_ST_HEADER;
        if (n - p < 7) return CJ_ST_TRUNCATED;
        const uint8_t flg = s[p + 4], bd = s[p + 5];
        const int bid = (bd >> 4) & 7;
        if ((flg >> 6) != 1 || bid < 4) return CJ_ST_HEADER;
        const size_t bmax = (size_t)1 << (8 + 2 * bid);
        const bool bsum = (flg >> 4) & 1, csize = (flg >> 3) & 1, csum = (flg >> 2) & 1, dict = flg & 1;
        const size_t dlen = 2 + (csize ? 8 : 0) + (dict ? 4 : 0);
        if (n - p < 4 + dlen + 1) return CJ_ST_TRUNCATED;
        uint64_t content = 0;
        if (csize) memcpy(&content, s + p + 6, 8);
        p += 4 + dlen + 1;
        size_t frame_tot = 0;
        for (;;) {
            if (n - p < 4) return CJ_ST_TRUNCATED;
            uint32_t bs; memcpy(&bs, s + p, 4);
            p += 4;
            if (bs == 0) break;
            const size_t blen = bs & 0x7FFFFFFFu;
            if (blen > n - p) return CJ_ST_TRUNCATED;
            const size_t badd = (bs >> 31) ? blen : std::min(bmax, blen * 255);
            if (badd > SIZE_MAX / 2 - frame_tot) return CJ_ST_TOO_BIG;
            frame_tot += badd;
            if (!(bs >> 31) && !csize) *exact = false;
            p += blen + (bsum ? 4 : 0);
            if (p > n) return CJ_ST_TRUNCATED;
        }
        if (csum) { if (n - p < 4) return CJ_ST_TRUNCATED; p += 4; }
        const uint64_t add = csize ? content : (uint64_t)frame_tot;
        if (add > (uint64_t)SIZE_MAX - tot) return CJ_ST_TOO_BIG;   // a wrapped sum must not come out as a small bound
        tot += (size_t)add;
        if (frames) frames->push_back({frame_at, p - frame_at, (size_t)add, (bool)csize});
    }
    *out = tot;
    return CJ_OK;
}


static int decompressed_bound(cj_codec codec, const void* src, size_t n, size_t* out, bool want_exact) {
    if (!out || (!src && n)) return CJ_E_INVALID_ARG;
    const uint8_t* s = (const uint8_t*)src;
    int32_t st = CJ_OK;
    bool exact = true;
    switch (codec) {
    case CJ_SNAPPY_RAW: {
        if (n == 0) { *out = 0; return CJ_OK; }  // snap::raw::decompress_len(b"") == Ok(0)
        uint64_t v;
        if (!cj::snappy_uvarint(s, n, &v)) st = CJ_ST_HEADER;
        else if (v > 0xFFFFFFFFull) st = CJ_ST_TOO_BIG;
        else *out = (size_t)v;
        break;
    }
    case CJ_SNAPPY_FRAMED: {
        uint64_t tot = 0;
        st = cj::snappy_frame_walk(s, n, nullptr, &tot);
        *out = (size_t)tot;
        break;
    }
    case CJ_LZ4_BLOCK:
        *out = n * 255;  // LZ4's maximum expansion; callers normally know the size (prefix / output_len)
        exact = false;
        break;
    case CJ_LZ4_FRAME: st = cj_lz4f_walk_host(s, n, out, &exact, nullptr); break;
    case CJ_ZSTD: st = cj_zstd_walk_host(s, n, out, &exact, nullptr); break;
    default: return CJ_E_INVALID_ARG;
    }
    if (st != CJ_OK) {
        cj_set_error("%s", cj_status_string(st));
        return CJ_E_UNIT_FAILED;
    }
    if (want_exact && !exact) {
        cj_set_error("the stream does not record its decompressed size");
        return CJ_E_UNIT_FAILED;
    }
    return CJ_OK;
}

extern "C" int cj_decompressed_len(cj_codec codec, const void* src, size_t n, size_t* out) { return decompressed_bound(codec, src, n, out, true); }
extern "C" int cj_decompress_bound(cj_codec codec, const void* src, size_t n, size_t* out) { return decompressed_bound(codec, src, n, out, false); }
