// lz_decode3.cu — generation-3 batch decode of Snappy raw blocks / LZ4 blocks (sm_100a): two kernels.
//
// Same reference entry points as lz_decode.cuh (snap::raw::Decoder::decompress, src/snappy.rs:55-60,103-108;
// LZ4_decompress_safe behind lz4::block::decompress_into, src/lz4.rs:78-95,140-173), for batches that are
// large enough to fill the machine with one THREAD per block in the first kernel.
//
//   1. g3_index_kernel  — the only inherently serial part of an LZ77 byte format is finding where
//      each element starts (an element's size is known only after its tag is read).  One thread walks
//      one block's tags (32 blocks per warp instruction instead of the 32-candidate pointer-doubling
//      parse of generation 2, ~9 warp instructions per element) and writes a 4-byte descriptor per
//      element: kind, length, and 12-bit deltas of (input position, output position) against an
//      8-byte base kept per row of 32 descriptors.  Long literals / long matches are cut into <= 60 /
//      <= 64 byte descriptors so the second kernel never sees an unusual element.  Compressed input
//      reaches the threads through per-lane shared-memory rings filled by cooperative, coalesced
//      cp.async (16 B per lane, consumed >= 8 iterations after issue, so no DRAM latency is exposed);
//      descriptors leave through a transposed shared-memory staging row, 128 B per store.
//   2. g3_exec_kernel   — one warp per block.  Every lane is a small state machine: it claims the next
//      descriptor in stream order, then moves up to CH bytes of its element per iteration, and claims
//      again as soon as it is done (no batch barrier: a lane never waits for the longest element of a
//      group).  Output bytes go into a shared-memory ring of 16-bit entries {generation, byte}: the
//      generation tag tells a reader whether the slot already holds the position it wants, which is
//      the whole dependency protocol for back-references — no scan, no frontier test, no rounds.
//      Sources already drained to HBM (ring slot overwritten) are re-read from global memory.
//   Anything this path does not take (unaligned input, a 4-byte-offset Snappy copy, a block the index
//   walk finds malformed, a bad offset) is put on a redo list and decoded by the generation-2 kernel,
//   which owns all error reporting: status codes stay exactly the oracle's.
#include "internal.h"
#include "lz_decode.cuh"

namespace cj {

constexpr uint32_t G3_REDO = 0xFFFFFFFFu;
constexpr uint32_t G3_MAX_SRC = 1u << 24;   // larger units go to generation 2
constexpr uint32_t G3_MAX_DST = 1u << 26;

struct G3 {
    uint32_t* desc_off;   // [n+1] first descriptor slot of each block (multiple of 32); desc_off[n] = total
    uint32_t* count;      // [n]   descriptors written for the block, or G3_REDO
    uint32_t* ulen;       // [n]   decompressed length found by the index walk
    uint32_t* redo_list;  // [n]   units for the generation-2 kernel
    unsigned* ctr;        // [0] exec work queue  [1] redo count  [2] redo work queue
    unsigned long long* total;  // descriptor slots wanted by the plan
    uint32_t* desc;       // descriptor arena
    uint2* rowbase;       // (input position, output position) of the first descriptor of each row of 32
    uint64_t arena;       // descriptor slots available in desc[]
};

// descriptor: [31:30] kind  [29:24] f  [23:12] input-position delta  [11:0] output-position delta
constexpr uint32_t K_LIT = 0, K_M16 = 1, K_M1 = 2, K_NULL = 3;
__device__ __forceinline__ uint32_t mk_desc(uint32_t kind, uint32_t f) { return (kind << 30) | (f << 24); }

__device__ __forceinline__ void g3_redo(const G3& g, uint32_t u) {
    const unsigned i = atomicAdd(&g.ctr[1], 1u);
    g.redo_list[i] = u;
    g.count[u] = G3_REDO;
}

// ------------------------------------------------------------------------------------------------
// plan: descriptor capacity per block (one thread per block, coalesced), then an exclusive scan in place (one CTA)
// ------------------------------------------------------------------------------------------------
template <int CODEC>
__global__ void __launch_bounds__(256) g3_caps_kernel(Batch b, G3 g) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const uint64_t sl = b.src_len[i], dc = b.dst_cap[i];
    const bool ok = sl >= 1 && sl <= G3_MAX_SRC && dc <= G3_MAX_DST && (((uintptr_t)(b.src_base + b.src_off[i])) & 15u) == 0;
    uint32_t cap = 0;
    if (ok) {
        uint64_t c = sl / 2 + 2;
        if (CODEC == CJ_LZ4_BLOCK) c += (dc < sl * 255 ? dc : sl * 255) / 64;
        cap = (uint32_t)((c + 31) & ~31ull) + 32;
    }
    g.desc_off[i] = cap;   // turned into an offset by the scan
    g.count[i] = 0;
}

__global__ void __launch_bounds__(1024) g3_scan_kernel(uint32_t n, G3 g) {
    __shared__ unsigned long long part[1024];
    const uint32_t t = threadIdx.x;
    const uint32_t lo = (uint32_t)((uint64_t)n * t / 1024), hi = (uint32_t)((uint64_t)n * (t + 1) / 1024);
    unsigned long long s = 0;
    for (uint32_t i = lo; i < hi; i++) s += g.desc_off[i];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        unsigned long long v = t >= (uint32_t)d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long acc = part[t] - s;
    const unsigned long long tot = part[1023];
    const bool fits = tot < 0xFFFF0000ull;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t c = fits ? g.desc_off[i] : 0;
        g.desc_off[i] = (uint32_t)acc;
        if (c == 0) g3_redo(g, i);   // not eligible (or nothing fits): generation 2
        acc += c;
    }
    if (t == 1023) {
        g.desc_off[n] = fits ? (uint32_t)tot : 0;
        *g.total = fits ? tot : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// kernel 1: index walk, one thread per block
// ------------------------------------------------------------------------------------------------
constexpr int IX_WARPS = 4;
constexpr int IX_RING = 256;     // per-lane input ring (two 128-byte chunks)
constexpr int IX_STRIDE = 33;    // staging row stride in words (bank-conflict free both ways)
constexpr int IX_SMEM_WARP = 32 * IX_RING + 32 * IX_STRIDE * 4;
constexpr int IX_SMEM_CTA = IX_SMEM_WARP * IX_WARPS + 1024;  // + the 256-entry tag table
#ifndef CJ_G3_UNROLL
#define CJ_G3_UNROLL 4
#endif
constexpr int IX_UNROLL = CJ_G3_UNROLL;   // plain elements a lane may index per iteration (amortises the per-iteration plumbing)
constexpr uint32_t IX_DELAY = 4;  // a chunk is read no earlier than this many iterations after its cp.async

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Snappy tag table: [6:0] compressed size of the element, [14:8] bytes it produces, [23:16] descriptor top byte
// (kind << 6 | f), bit 31 = not a plain element (literal with length bytes, 4-byte-offset copy).
__device__ __forceinline__ uint32_t snappy_tag_entry(uint32_t tag) {
    const uint32_t type = tag & 3, L = (tag >> 2) + 1;
    if (type == 0) return L <= 60 ? ((1 + L) | (L << 8) | (((K_LIT << 6) | (L - 1)) << 16)) : 0x80000000u;
    if (type == 1) {
        const uint32_t len = 4 + ((tag >> 2) & 7);
        return 2 | (len << 8) | (((K_M1 << 6) | ((len - 4) << 3) | (tag >> 5)) << 16);
    }
    if (type == 2) return 3 | (L << 8) | (((K_M16 << 6) | (L - 1)) << 16);
    return 0x80000000u;
}

template <int CODEC>
__global__ void __launch_bounds__(IX_WARPS * 32, 4) g3_index_kernel(Batch b, G3 g) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* wsm = smem + (size_t)warp * IX_SMEM_WARP;
    const uint32_t ring0 = smem_addr(wsm);
    const uint32_t ring = ring0 + lane * IX_RING;
    const uint32_t stage0 = smem_addr(wsm + 32 * IX_RING);
    const uint32_t stage = stage0 + lane * IX_STRIDE * 4;
    const uint32_t lut = smem_addr(smem + (size_t)IX_SMEM_WARP * IX_WARPS);
    if (CODEC == CJ_SNAPPY_RAW) {
        for (uint32_t t = threadIdx.x; t < 256; t += blockDim.x) sts32(lut + 4 * t, snappy_tag_entry(t));
        __syncthreads();
    }
    const uint32_t nwarps = gridDim.x * IX_WARPS;

    for (uint32_t first = (blockIdx.x * IX_WARPS + warp) * 32; first < b.n; first += nwarps * 32) {
        // ---- every lane takes one block ----
        const uint32_t cur = first + lane;
        bool active = false;
        const uint8_t* src = nullptr;
        uint32_t n = 0, ip = 0, op = 0, ulen = 0, e = 0, doff = 0;
        if (cur < b.n) {
            doff = g.desc_off[cur];
            if (g.desc_off[cur + 1] != doff) {  // else: the plan already sent it to generation 2
                const uint64_t dcap = b.dst_cap[cur];
                src = b.src_base + b.src_off[cur];
                n = (uint32_t)b.src_len[cur];
                bool ok;
                if (CODEC == CJ_SNAPPY_RAW) {
                    uint64_t v = 0;
                    bool done = false;
                    for (int i = 0; i < 5 && ip < n; i++) {
                        const uint32_t x = ldg_u8(src + ip++);
                        v |= (uint64_t)(x & 0x7f) << (7 * i);
                        if (!(x & 0x80)) { done = true; break; }
                    }
                    ok = done && v <= dcap && v <= MAX_UNIT;
                    ulen = (uint32_t)v;
                } else {
                    ulen = (uint32_t)dcap;  // capacity: LZ4_decompress_safe's rules are applied against it
                    ok = dcap != 0;
                }
                if (ok) active = true;
                else g3_redo(g, cur);
            }
        }
        const uint32_t n16 = n & ~15u;   // the ragged last granule is read straight from global memory
        uint32_t loaded = 0, ready_old = IX_DELAY, ready_new = IX_DELAY, row_ip = 0, row_op = 0;
        uint32_t litrem = 0;   // bytes of a long literal / LZ4 literal run still to be described (input position = ip)
        uint32_t mrem = 0, mip = 0, nextp = 0;  // LZ4: match bytes still to be described, position of its offset, next token
        bool lz_last = false;
        uint32_t iter = 0;

        while (__any_sync(FULL, active)) {
            // ---- input ring: issue the next 128-byte chunk of every lane that has entered its newest one ----
            const bool reading = active && litrem == 0 && mrem == 0 && ip < n;
            const bool want = reading && ip + 128 >= loaded;
            if (want && ip >= loaded) loaded = ip & ~127u;  // first chunk of the block, or a jump over a long literal
            uint32_t wm = __ballot_sync(FULL, want);
            while (wm) {
                const int j = __ffs(wm) - 1;
                wm &= wm - 1;
                const uint32_t go = __shfl_sync(FULL, loaded, j) + 16 * lane, nj = __shfl_sync(FULL, n16, j);
                const uint8_t* sp = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, (unsigned long long)src, j));
                if (lane < 8 && go < nj) cp_async16(ring0 + j * IX_RING + (go & (IX_RING - 1)), sp + go);
            }
            if (want) {
                loaded += 128;
                ready_old = ready_new;
                ready_new = iter + IX_DELAY;
            }
            const uint32_t limit = iter >= ready_new ? loaded : loaded - 128;
            const bool can_read = reading && iter >= ready_old && ip + 8 <= limit;

            bool emitted = false, rowfull = false, fin = false, fail = false;
            // one descriptor into this lane's staging row; a full row ends the lane's work for this iteration
            auto put_desc = [&](uint32_t dtop, uint32_t ipd, uint32_t opd) {
                const uint32_t k = e & 31;
                if (k == 0) {
                    row_ip = ipd; row_op = opd;
                    g.rowbase[(doff + e) >> 5] = make_uint2(ipd, opd);
                }
                const uint32_t dip = ipd - row_ip, dop = opd - row_op;
                if (CODEC != CJ_SNAPPY_RAW && (dip > 4095 || dop > 4095)) { fail = true; return; }  // Snappy: <= 31 x 65 by construction
                sts32(stage + k * 4, dtop | (dip << 12) | dop);
                e++;
                emitted = true;
                rowfull = (e & 31) == 0;
            };
            if (CODEC == CJ_SNAPPY_RAW) {
                // ---- hot path: up to IX_UNROLL plain elements per lane and iteration, table-driven ----
                bool stop = !(reading && iter >= ready_old);
#pragma unroll
                for (int u = 0; u < IX_UNROLL; u++) {
                    if (!stop && !rowfull && ip < n && ip + 8 <= limit) {
                        const uint32_t tag = ip < n16 ? lds8(ring + (ip & (IX_RING - 1))) : ldg_u8(src + ip);
                        const uint32_t ent = lds32(lut + 4 * tag);
                        const uint32_t adv = ent & 0x7F, len = (ent >> 8) & 0x7F;
                        if (ent >> 31) stop = true;   // literal with length bytes / 4-byte-offset copy: the rare path below
                        else if (adv > n - ip || len > ulen - op) { fail = true; stop = true; }
                        else {
                            put_desc((ent << 8) & 0xFF000000u, ip + 1, op);
                            ip += adv; op += len;
                        }
                    }
                }
                if (active && !emitted && !fail) {  // rare states (a lane that made progress above comes back next iteration)
                    if (litrem) {
                        const uint32_t len = min(litrem, 60u);
                        put_desc(mk_desc(K_LIT, len - 1), ip, op);
                        ip += len; op += len; litrem -= len;
                    } else if (ip >= n) {
                        fin = true;
                        fail = op != ulen;
                    } else if (can_read) {
                        const uint32_t tag = ip < n16 ? lds8(ring + (ip & (IX_RING - 1))) : ldg_u8(src + ip);
                        const uint32_t nb = (tag >> 2) + 1 - 60;
                        if ((tag & 3) != 0 || nb > n - ip - 1) fail = true;   // 4-byte-offset copies: generation 2
                        else {
                            uint32_t v = 0;
                            for (uint32_t i = 0; i < nb; i++) {
                                const uint32_t p = ip + 1 + i;
                                v |= (p < n16 ? lds8(ring + (p & (IX_RING - 1))) : ldg_u8(src + p)) << (8 * i);
                            }
                            const uint64_t LL = (uint64_t)v + 1;
                            ip += 1 + nb;
                            if (LL > n - ip || LL > ulen - op) fail = true;
                            else litrem = (uint32_t)LL;
                        }
                    }
                }
            } else {
                // LZ4: a sequence is described as literal chunks (<= 64 B) followed by match chunks (<= 64 B).
                // The checks are lz4_serial_step's (lz_decode.cuh); whatever they would reject goes to generation 2.
                if (active) {
                    if (litrem == 0 && mrem == 0 && !lz_last) {
                        if (ip >= n) fail = true;  // a valid block ends inside a literal-only sequence
                        else if (can_read) {
                            auto rb = [&](uint32_t p) -> uint32_t { return p < n16 ? lds8(ring + (p & (IX_RING - 1))) : ldg_u8(src + p); };
                            const uint32_t token = rb(ip);
                            uint32_t q = ip + 1, ll = token >> 4, ml = token & 15;
                            bool irregular = false;
                            if (ll == 15) {
                                if (n < 15 || q >= n - 15) irregular = true;
                                else {
                                    uint32_t x = 255, cnt = 0;
                                    while (x == 255 && cnt < 4) {   // the ring guarantees ip + 8 bytes; longer runs: generation 2
                                        x = rb(q);
                                        q++; ll += x; cnt++;
                                        if (q > n - 15) { irregular = true; break; }
                                    }
                                    irregular = irregular || x == 255;
                                }
                            }
                            if (irregular) fail = true;
                            else if ((uint64_t)op + ll + 12 > ulen || (uint64_t)q + ll + 8 > n) {
                                // tail zone: must be the final literal-only sequence, ending exactly at n
                                if ((uint64_t)q + ll != n || (uint64_t)op + ll > ulen) fail = true;
                                else { ip = q; lz_last = true; litrem = ll; }
                            } else {
                                // regular sequence (at least 6 bytes follow the offset); the match-length bytes sit behind the
                                // literals, possibly outside the ring window: read from global memory (L1/L2)
                                const uint32_t po = q + ll;
                                uint32_t r = po + 2;
                                if (ml == 15) {
                                    uint32_t x = 255, cnt = 0;
                                    while (x == 255 && cnt < 16) {
                                        x = ldg_u8(src + r);
                                        r++; ml += x; cnt++;
                                        if (r > n - 4) { irregular = true; break; }
                                    }
                                    irregular = irregular || x == 255;
                                }
                                ml += 4;
                                if (irregular || (uint64_t)op + ll + ml + 5 > ulen) fail = true;
                                else { ip = q; litrem = ll; mip = po; mrem = ml; nextp = r; }
                            }
                        }
                    }
                    if (!fail) {  // describe one chunk of the current sequence (also in the iteration that parsed it)
                        if (litrem) {
                            const uint32_t len = min(litrem, 64u);
                            put_desc(mk_desc(K_LIT, len - 1), ip, op);
                            ip += len; op += len; litrem -= len;
                        } else if (mrem) {
                            const uint32_t len = min(mrem, 64u);
                            put_desc(mk_desc(K_M16, len - 1), mip, op);
                            op += len; mrem -= len;
                            if (mrem == 0) ip = nextp;
                        }
                        if (lz_last && litrem == 0) fin = true;
                    }
                }
            }
            if (fail) {
                g3_redo(g, cur);
                active = false;
                fin = false;
            }
            // ---- rows leave the staging area: full rows, and the partial last row of a finished block ----
            const bool flushrow = active && (rowfull || (fin && (e & 31) != 0));
            uint32_t fm = __ballot_sync(FULL, flushrow);
            if (fm) {
                const uint32_t rcount = (e & 31) ? (e & 31) : 32;
                const uint32_t rstart = doff + ((e - 1) & ~31u);
                __syncwarp();
                while (fm) {
                    const int j = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const uint32_t rs = __shfl_sync(FULL, rstart, j), rc = __shfl_sync(FULL, rcount, j);
                    if ((uint32_t)lane < rc) g.desc[rs + lane] = lds32(stage0 + (j * IX_STRIDE + lane) * 4);
                }
            }
            if (fin) {  // fail cleared fin
                g.count[cur] = e;
                g.ulen[cur] = (CODEC == CJ_SNAPPY_RAW) ? ulen : op;
                active = false;
            }
            cp_async_commit();
            cp_async_wait<IX_DELAY - 1>();
            __syncwarp();
            iter++;
        }
        cp_async_wait<0>();
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// kernel 2: execute, one warp per block, dynamic lane state machine
// ------------------------------------------------------------------------------------------------
#ifndef CJ_G3_LOG_OR
#define CJ_G3_LOG_OR 11
#endif
#ifndef CJ_G3_CH
#define CJ_G3_CH 8
#endif
#ifndef CJ_G3_CTAS
#define CJ_G3_CTAS 7
#endif
constexpr int X_LOG_OR = CJ_G3_LOG_OR;
constexpr uint32_t X_OR = 1u << X_LOG_OR;          // output ring entries (16 bit each)
constexpr uint32_t X_OMASK = X_OR - 1;
constexpr uint32_t X_DR = 256;                      // descriptor ring entries
constexpr uint32_t X_FLUSH = X_OR / 4;              // drain when this many complete bytes are pending
constexpr int X_WARPS = 4;
constexpr int X_SMEM_WARP = X_OR * 2 + IRING + X_DR * 4 + (X_DR / 32) * 8;
constexpr int CH = CJ_G3_CH;

__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

struct XOut {
    uint32_t ring;   // shared address of the entry ring
    uint8_t* dst;
    uint32_t a;      // dst & 15; ring entry of output byte p = (p + a) & X_OMASK
    uint32_t flushed;
    int lane;

    __device__ __forceinline__ uint32_t eaddr(uint32_t p) const { return ring + 2 * ((p + a) & X_OMASK); }

    // Stores [flushed, upto) to global memory (all of it complete): ragged head by bytes, 16-byte vector body, tail if final.
    __device__ __forceinline__ void flush_to(uint32_t upto, bool final) {
        uint32_t q = flushed;
        const uint32_t mis = (q + a) & 15u;
        if (mis != 0 && q < upto) {
            const uint32_t head = min(16u - mis, upto - q);
            if ((uint32_t)lane < head) dst[q + lane] = (uint8_t)lds16(eaddr(q + lane));
            q += head;
        }
        const uint32_t vend = q + ((upto - q) & ~15u);
        for (uint32_t p = q + lane * 16; p < vend; p += 512) {
            const uint32_t sa = eaddr(p);  // (p + a) is a multiple of 16: 32 contiguous, aligned bytes of entries
            const uint4 lo = lds128(sa), hi = lds128(sa + 16);
            uint4 v;
            v.x = __byte_perm(lo.x, lo.y, 0x6420);
            v.y = __byte_perm(lo.z, lo.w, 0x6420);
            v.z = __byte_perm(hi.x, hi.y, 0x6420);
            v.w = __byte_perm(hi.z, hi.w, 0x6420);
            *reinterpret_cast<uint4*>(dst + p) = v;
        }
        q = vend;
        if (final) {
            if (q + lane < upto) dst[q + lane] = (uint8_t)lds16(eaddr(q + lane));
            q = upto;
        }
        flushed = q;
        __syncwarp();
    }
};

template <int CODEC>
__global__ void __launch_bounds__(X_WARPS * 32, CJ_G3_CTAS) g3_exec_kernel(Batch b, G3 g) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* wsm = smem + (size_t)warp * X_SMEM_WARP;
    const uint32_t oring = smem_addr(wsm);
    const uint32_t dring = smem_addr(wsm + X_OR * 2 + IRING);
    const uint32_t rbring = dring + X_DR * 4;
    const uint32_t ltmask = (1u << lane) - 1;

    for (;;) {
        const uint32_t u = next_unit(&g.ctr[0], lane);
        if (u >= b.n) break;
        const uint32_t total_e = g.count[u];
        if (total_e == G3_REDO) continue;
        const uint32_t ulen = g.ulen[u];
        const uint8_t* src = b.src_base + b.src_off[u];
        const uint32_t n = (uint32_t)b.src_len[u];
        uint8_t* dst = b.dst_base + b.dst_off[u];
        const uint32_t doff = g.desc_off[u];
        const uint32_t* __restrict__ gdesc = g.desc + doff;
        const uint2* __restrict__ grow = g.rowbase + (doff >> 5);

        XOut out;
        out.ring = oring; out.dst = dst; out.a = (uint32_t)((uintptr_t)dst & 15u); out.flushed = 0; out.lane = lane;
        // every slot starts with generation 0xFF, which no position below 255 * X_OR carries; later laps overwrite all slots
        for (uint32_t i = lane * 16; i < X_OR * 2; i += 512) sts128(oring + i, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
        InRing in;
        in.init(wsm + X_OR * 2, src, n, lane);
        __syncwarp();

        uint32_t next_e = 0, d_loaded = 0, hi_op = 0, hi_ip = 0;
        // per-lane element state
        uint32_t rem = 0, dpos = 0, sp = 0, estart = 0, eidx = 0;
        bool is_lit = false;
        bool bad = false;
        if (total_e) hi_ip = grow[0].x;
        in.refill(hi_ip);

        for (;;) {
            // ---- descriptor ring: 128 more descriptors + their 4 row bases when the claim front gets close ----
            if (next_e + 32 > d_loaded && d_loaded < total_e) {
                const uint32_t lowE = __reduce_min_sync(FULL, rem ? eidx : next_e);
                if (d_loaded + 128 - lowE <= X_DR) {
                    const uint32_t i4 = d_loaded + lane * 4;   // descriptor slots are padded to rows of 32, always readable
                    const uint4 v = *reinterpret_cast<const uint4*>(gdesc + i4);
                    sts128(dring + (i4 & (X_DR - 1)) * 4, v);
                    if (lane < 4) {
                        const uint32_t r = (d_loaded >> 5) + lane;
                        const uint2 rb = grow[r];
                        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(rbring + (r & (X_DR / 32 - 1)) * 8), "r"(rb.x), "r"(rb.y) : "memory");
                    }
                    d_loaded += 128;
                    __syncwarp();
                }
            }
            // ---- claim: idle lanes take the next descriptors in stream order ----
            {
                const bool need = rem == 0;
                const uint32_t m = __ballot_sync(FULL, need);
                const uint32_t idx = next_e + __popc(m & ltmask);
                const uint32_t lim = min(d_loaded, total_e);
                bool ok = need && idx < lim;
                uint32_t kind = K_NULL, len = 0, ipd = 0, op = 0, off = 0;
                if (ok) {
                    const uint32_t d = lds32(dring + (idx & (X_DR - 1)) * 4);
                    uint32_t bx, by;
                    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(bx), "=r"(by) : "r"(rbring + ((idx >> 5) & (X_DR / 32 - 1)) * 8));
                    kind = d >> 30;
                    const uint32_t f = (d >> 24) & 63;
                    ipd = bx + ((d >> 12) & 0xFFF);
                    op = by + (d & 0xFFF);
                    len = kind == K_M1 ? 4 + (f >> 3) : f + 1;
                    if (kind == K_NULL) len = 0;
                    // room in the rings: the slot of p is free once p - X_OR has been drained; input must be resident
                    ok = op + len <= out.flushed + X_OR && (ipd + 66 + in.a <= in.loaded || in.loaded >= ((n + in.a + 15) & ~15u));
                    if (ok && kind != K_LIT && kind != K_NULL) {
                        const uint32_t b0 = in.byte(ipd);
                        if (kind == K_M16) off = b0 | (in.byte(ipd + 1) << 8);
                        else off = ((f & 7) << 8) | b0;
                        if (off == 0 || off > op) bad = true;
                    }
                }
                uint32_t okm = __ballot_sync(FULL, ok);
                const uint32_t failm = m & ~okm;          // needing lanes that could not claim: every later rank must wait too
                if (failm) okm &= (1u << (__ffs(failm) - 1)) - 1;
                if ((okm >> lane) & 1) {
                    rem = len; dpos = op; estart = op; eidx = idx;
                    is_lit = kind == K_LIT;
                    sp = is_lit ? ipd : off;
                }
                if (okm) {
                    const int last = 31 - __clz(okm);
                    next_e += __popc(okm);
                    hi_op = __shfl_sync(FULL, op + len, last);
                    hi_ip = __shfl_sync(FULL, ipd + (kind == K_LIT ? len : 0), last);
                }
            }
            if (__any_sync(FULL, bad)) break;

            // ---- move up to CH bytes of the lane's element (word-wise: bytes travel packed in 32-bit registers) ----
            if (rem) {
                constexpr int NW = CH / 4;
                const uint32_t di = (dpos + out.a) & X_OMASK;
                uint32_t c = min(min((uint32_t)CH, rem), X_OR - di);
                uint32_t w[NW];
                bool ready = true;
                if (is_lit) {
                    const uint32_t si = (sp + in.a) & IMASK;
                    c = min(c, (uint32_t)IRING - si);
                    const uint32_t sa = in.ring + si, al = sa & ~3u, sh = (sa & 3u) * 8;
                    uint32_t x[NW + 1];
#pragma unroll
                    for (int k = 0; k <= NW; k++) x[k] = lds32(al + 4 * k);  // reads past si + c stay inside this warp's shared memory
#pragma unroll
                    for (int k = 0; k < NW; k++) w[k] = __funnelshift_r(x[k], x[k + 1], sh);
                } else {
                    const uint32_t s = dpos - sp;
                    c = min(c, sp);
                    const uint32_t si = (s + out.a) & X_OMASK;
                    c = min(c, X_OR - si);
                    const uint32_t ea = oring + 2 * si, al = ea & ~7u, sh = (ea & 7u) * 8;  // sh in {0,16,32,48}
                    uint32_t q[2 * NW + 2];
#pragma unroll
                    for (int k = 0; k < NW + 1; k++) {
                        uint32_t lo, hi;
                        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(al + 8 * k));
                        q[2 * k] = lo; q[2 * k + 1] = hi;
                    }
                    const bool up = (sh & 32u) != 0;
                    const uint32_t s5 = sh & 31u;
                    const uint32_t E = (((s + out.a) >> X_LOG_OR) & 0xFF) * 0x01000100u;   // expected generation in both tag bytes
                    uint32_t miss = 0;
#pragma unroll
                    for (int k = 0; k < NW; k++) {
                        const uint32_t b0 = up ? q[2 * k + 1] : q[2 * k];
                        const uint32_t b1 = up ? q[2 * k + 2] : q[2 * k + 1];
                        const uint32_t b2 = up ? q[2 * k + 3] : q[2 * k + 2];
                        const uint32_t e01 = __funnelshift_r(b0, b1, s5), e23 = __funnelshift_r(b1, b2, s5);
                        const uint32_t cc = c > (uint32_t)(4 * k) ? c - 4 * k : 0u;   // entries of this word that matter
                        const uint32_t m01 = cc >= 2 ? 0xFF00FF00u : (cc == 1 ? 0x0000FF00u : 0u);
                        const uint32_t m23 = cc >= 4 ? 0xFF00FF00u : (cc == 3 ? 0x0000FF00u : 0u);
                        miss |= ((e01 ^ E) & m01) | ((e23 ^ E) & m23);
                        w[k] = __byte_perm(e01, e23, 0x6420);
                    }
                    if (miss) {
                        if (s < out.flushed) {  // slot overwritten: the bytes were drained, re-read them from global memory
                            c = min(c, out.flushed - s);
                            const uintptr_t p = (uintptr_t)(dst + s);
                            const uint32_t* ap = reinterpret_cast<const uint32_t*>(p & ~(uintptr_t)3);
                            const uint32_t gsh = ((uint32_t)p & 3u) * 8;
                            uint32_t x[NW + 1];
#pragma unroll
                            for (int k = 0; k <= NW; k++) x[k] = ap[k];   // at most 7 bytes past s + c, still far below dpos
#pragma unroll
                            for (int k = 0; k < NW; k++) w[k] = __funnelshift_r(x[k], x[k + 1], gsh);
                        } else {
                            ready = false;  // source not produced yet
                        }
                    }
                }
                if (ready) {
                    const uint32_t gd = ((dpos + out.a) >> X_LOG_OR) & 0xFF;
                    const uint32_t da = oring + 2 * di;
#pragma unroll
                    for (int k = 0; k < CH; k++)
                        if ((uint32_t)k < c) sts16(da + 2 * k, __byte_perm(w[k / 4], gd, 0x0040 | (k & 3)));
                    dpos += c;
                    rem -= c;
                    if (is_lit) sp += c;
                }
            }
            __syncwarp();

            // ---- maintenance: frontier, drain, input refill, termination ----
            uint32_t F = __reduce_min_sync(FULL, rem ? estart : 0xFFFFFFFFu);
            if (F == 0xFFFFFFFFu) {
                F = hi_op;
                if (next_e >= total_e) break;  // everything claimed and finished
            }
            if (F - out.flushed >= X_FLUSH) out.flush_to(((F + out.a) & ~15u) - out.a, false);
            if (hi_ip + 192 + in.a > in.loaded && in.loaded < ((n + in.a + 15) & ~15u)) {
                const uint32_t first = __reduce_min_sync(FULL, (rem && is_lit) ? sp : hi_ip);
                in.refill(first);
            }
        }
        __syncwarp();
        if (__any_sync(FULL, bad) || hi_op != ulen) {
            if (lane == 0) g3_redo(g, u);
        } else {
            out.flush_to(ulen, true);
            if (lane == 0) {
                b.dst_len[u] = ulen;
                b.status[u] = CJ_OK;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// generation-2 kernel over the redo list
// ------------------------------------------------------------------------------------------------
template <int CODEC>
__global__ void __launch_bounds__(DEC_WARPS * 32, CJ_DEC_CTAS) lz_decode_list_kernel(Batch b, G3 g) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* smem_warp = smem + (size_t)warp * DEC_SMEM_WARP;
    const uint32_t count = g.ctr[1];
    for (;;) {
        const uint32_t i = next_unit(&g.ctr[2], lane);
        if (i >= count) break;
        const uint32_t u = g.redo_list[i];
        const uint64_t slen = b.src_len[u], dcap = b.dst_cap[u];
        uint32_t produced = 0;
        int32_t st;
        if (slen > MAX_UNIT) st = CJ_ST_TOO_BIG;
        else st = decode_block<CODEC, true>(b.src_base + b.src_off[u], (uint32_t)slen, b.dst_base + b.dst_off[u], dcap > MAX_UNIT ? MAX_UNIT : (uint32_t)dcap, smem_warp, lane, &produced);
        if (lane == 0) {
            b.dst_len[u] = st == CJ_OK ? produced : 0;
            b.status[u] = st;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <class K>
static cudaError_t set_smem(K k, size_t bytes) {
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int CODEC>
static cudaError_t launch_g3(const Batch& b, G3Scratch& sc, int sm_count, cudaStream_t stream, G3Debug* dbg) {
    static cj_per_device_flag attr_flag;
    int& attr_done = attr_flag.here();
    if (!attr_done) {
        cudaError_t e;
        if ((e = set_smem(g3_index_kernel<CODEC>, (size_t)IX_SMEM_CTA)) != cudaSuccess) return e;
        if ((e = set_smem(g3_exec_kernel<CODEC>, (size_t)X_SMEM_WARP * X_WARPS)) != cudaSuccess) return e;
        if ((e = set_smem(lz_decode_list_kernel<CODEC>, (size_t)DEC_SMEM_WARP * DEC_WARPS)) != cudaSuccess) return e;
        attr_done = 1;
    }
    const size_t n = b.n;
    // fixed part: desc_off[n+1] count[n] ulen[n] redo_list[n] ctr[4] total
    const size_t fixed_words = (n + 1) + 3 * n + 4 + 4;
    if (sc.ensure_fixed(fixed_words * 4 + 64) != 0) return cudaErrorMemoryAllocation;
    G3 g;
    uint32_t* w = (uint32_t*)sc.fixed();
    g.desc_off = w; w += n + 1;
    g.count = w; w += n;
    g.ulen = w; w += n;
    g.redo_list = w; w += n;
    w = (uint32_t*)(((uintptr_t)w + 15) & ~(uintptr_t)15);
    g.ctr = w; w += 4;
    g.total = (unsigned long long*)w;
    g.desc = nullptr; g.rowbase = nullptr; g.arena = 0;
    cudaError_t e = cudaMemsetAsync(g.ctr, 0, 4 * sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    g3_caps_kernel<CODEC><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(b, g);
    g3_scan_kernel<<<1, 1024, 0, stream>>>((uint32_t)n, g);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    unsigned long long total = 0;
    if ((e = cudaMemcpyAsync(&total, g.total, sizeof total, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    const size_t arena_bytes = (((size_t)total * 4 + 127) & ~(size_t)127) + (size_t)(total / 32 + 1) * 8 + 1024;
    if (sc.ensure_arena(arena_bytes) != 0) return cudaErrorMemoryAllocation;
    g.desc = (uint32_t*)sc.arena();
    g.rowbase = (uint2*)((uint8_t*)sc.arena() + (((size_t)total * 4 + 127) & ~(size_t)127));
    g.arena = total;
    if (total) {
        int grid = (int)std::min<size_t>((n + IX_WARPS * 32 - 1) / (IX_WARPS * 32), (size_t)sm_count * 4);
        g3_index_kernel<CODEC><<<grid, IX_WARPS * 32, (size_t)IX_SMEM_CTA, stream>>>(b, g);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (!(dbg && dbg->index_only)) {
            grid = (int)std::min<size_t>((n + X_WARPS - 1) / X_WARPS, (size_t)sm_count * CJ_G3_CTAS);
            g3_exec_kernel<CODEC><<<grid, X_WARPS * 32, (size_t)X_SMEM_WARP * X_WARPS, stream>>>(b, g);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    }
    if (dbg) {
        dbg->desc_off = g.desc_off; dbg->count = g.count; dbg->ulen = g.ulen; dbg->desc = g.desc; dbg->rowbase = (const uint32_t*)g.rowbase;
        dbg->redo_count = g.ctr + 1; dbg->total = total;
        if (dbg->index_only) return cudaSuccess;
    }
    {
        int grid = (int)std::min<size_t>((n + DEC_WARPS - 1) / DEC_WARPS, (size_t)sm_count * CJ_DEC_CTAS);
        lz_decode_list_kernel<CODEC><<<grid, DEC_WARPS * 32, (size_t)DEC_SMEM_WARP * DEC_WARPS, stream>>>(b, g);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// The generation-2 kernel over a redo list filled by another path (lz_decode4.cu): ctr[1] = entries, ctr[2] = work queue.
cudaError_t launch_lz_decode_list(int codec, const Batch& b, uint32_t* redo_list, unsigned* ctr, int sm_count, cudaStream_t stream) {
    static cj_per_device_flag attr_flag;
    int& attr_done = attr_flag.here();
    if (!attr_done) {
        cudaError_t e;
        if ((e = set_smem(lz_decode_list_kernel<CJ_SNAPPY_RAW>, (size_t)DEC_SMEM_WARP * DEC_WARPS)) != cudaSuccess) return e;
        if ((e = set_smem(lz_decode_list_kernel<CJ_LZ4_BLOCK>, (size_t)DEC_SMEM_WARP * DEC_WARPS)) != cudaSuccess) return e;
        attr_done = 1;
    }
    G3 g{};
    g.redo_list = redo_list;
    g.ctr = ctr;
    const int grid = (int)std::min<size_t>(((size_t)b.n + DEC_WARPS - 1) / DEC_WARPS, (size_t)sm_count * CJ_DEC_CTAS);
    if (codec == CJ_LZ4_BLOCK) lz_decode_list_kernel<CJ_LZ4_BLOCK><<<grid, DEC_WARPS * 32, (size_t)DEC_SMEM_WARP * DEC_WARPS, stream>>>(b, g);
    else lz_decode_list_kernel<CJ_SNAPPY_RAW><<<grid, DEC_WARPS * 32, (size_t)DEC_SMEM_WARP * DEC_WARPS, stream>>>(b, g);
    return cudaGetLastError();
}

cudaError_t launch_lz_decode3(int codec, const Batch& b, G3Scratch& sc, int sm_count, cudaStream_t stream, G3Debug* dbg) {
    return codec == CJ_LZ4_BLOCK ? launch_g3<CJ_LZ4_BLOCK>(b, sc, sm_count, stream, dbg) : launch_g3<CJ_SNAPPY_RAW>(b, sc, sm_count, stream, dbg);
}

}  // namespace cj
