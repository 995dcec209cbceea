// lz_decode.cu — LZ4 block and Snappy raw block decode kernels (sm_100a).
//
// Replaces, for a whole batch of independent blocks per launch:
//   LZ4   : lz4::block::decompress_into -> LZ4_decompress_safe      (reference src/lz4.rs:78-95,140-173)
//   Snappy: snap::raw::Decoder::decompress                          (reference src/snappy.rs:55-60,103-108)
//
// Execution model: persistent grid, one warp per block, blocks handed out by a grid-wide atomic
// work queue.  Per warp, in shared memory: an 8 KiB output ring (every produced byte lands there
// first; back-references up to 6 KiB away are served from it; it drains to HBM with 16-byte vector
// stores, 512 B per warp instruction), a 4 KiB input ring (refilled with 16-byte vector loads) and
// a 64-entry element queue.
//
// Generation 2 ("lane-parallel") data path, per warp:
//   1. PARSE   32 candidate positions at once: every lane computes the compressed size of the
//              element that would start at its byte; 4 pointer-doubling steps (2 shuffles each)
//              resolve which of the 32 positions are real element starts; those positions are
//              compacted into the queue.
//   2. EXECUTE 32 elements at once, one lane per element: decode lengths/offset, warp prefix
//              scan gives every element its output position, validity checks by ballot, then
//              literals are copied lane-parallel and back-references in dependency rounds (a copy
//              is ready once its source lies below the frontier of completed output).
//   Anything unusual — long literals, lz4 length extensions > 1 byte, the final lz4 sequence, any
//   element that fails a check — is cut out of the batch and handled by the warp-uniform SERIAL
//   path (generation 1), which also owns all error reporting.  Acceptance rules and status codes
//   are identical to oracle/lz4.c and oracle/snappy.c.
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace cj {

#ifndef CJ_ORING
#define CJ_ORING 4096
#endif
#ifndef CJ_IRING
#define CJ_IRING 2048
#endif
#ifndef CJ_TMAX
#define CJ_TMAX 1024
#endif
#ifndef CJ_DEC_CTAS
#define CJ_DEC_CTAS 8
#endif
constexpr int ORING = CJ_ORING;   // output ring bytes per warp
constexpr int IRING = CJ_IRING;   // input ring bytes per warp
constexpr int QCAP = 64;          // element queue entries per warp
constexpr uint32_t OMASK = ORING - 1, IMASK = IRING - 1;
constexpr uint32_t TMAX = CJ_TMAX;            // max output bytes of one lane-parallel batch
constexpr uint32_t PARSE_SPAN = IRING / 2 - 512;  // max compressed bytes between first queued element and parse cursor
// refill() leaves at least IRING-1022 bytes resident past the first queued element; a parse window
// starting inside the span looks at most 31 + 288 bytes further (largest lane-parallel element).
static_assert(IRING >= PARSE_SPAN + 1022 + 352, "input ring too small");
constexpr uint32_t FAR_T = ORING - TMAX;    // back-reference distance beyond which the source is re-read from global
constexpr int STEP_CONT = -1, STEP_DONE = 0;  // serial step results (else a CJ_ST_* error)

// Explicit shared-memory accessors on 32-bit shared addresses (keeps every ring access an LDS/STS).
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// ---- bulk asynchronous copies (TMA; cp.async.bulk -> SASS UBLKCP) and the mbarrier that tracks the loads ----
// The rings of the warp-per-block kernels are staged with them (north_star: "input/output staged through shared memory via
// TMA bulk copies"): one elected lane hands a 16-byte aligned span to the copy engine, no lane moves a byte of it.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W;\n\t}" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t mbar) {   // global -> shared, completes on mbar
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t smem_src, uint32_t bytes) {   // shared -> global, joins the thread's bulk group
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }   // the group's writes are performed
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }   // my shared-memory writes -> async proxy

struct OutRing {
    static constexpr uint32_t CHUNK = ORING / 8;    // largest span moved between room checks (serial path)
    static constexpr uint32_t FLUSH_T = ORING / 4;  // drain when this many bytes are pending
    static_assert(ORING >= 2 * CHUNK + FLUSH_T + 16, "ring too small for the far-match invariant");
    static_assert(ORING >= TMAX + FLUSH_T + CHUNK + 64, "ring too small for a batch");
    static_assert(ORING >= 2 * TMAX + FLUSH_T + CHUNK + 32, "far sources of a batch must already be drained");

    uint32_t ring;     // shared-space address of this warp's output ring
    const uint8_t* ring_g;  // the same ring as a generic pointer (lane-parallel copies mix ring and global sources)
    uint8_t* dst;
    uint32_t a;        // dst misalignment (dst & 15); ring index = (pos + a) & OMASK
    uint32_t op;       // bytes produced so far
    uint32_t flushed;  // bytes already stored to global
    uint32_t base;     // lowest output position a back-reference may reach (0; or the block start inside an independent-block frame)
    uint32_t settled;  // output below this position is in global memory for certain; [settled, flushed) may still be on its way (bulk store in flight)
    int lane;

    __device__ __forceinline__ void init(uint8_t* ring_, uint8_t* dst_, int lane_) {
        ring = smem_addr(ring_);
        ring_g = ring_;
        dst = dst_;
        a = (uint32_t)((uintptr_t)dst_ & 15u);
        op = 0;
        flushed = 0;
        base = 0;
        settled = 0;
        lane = lane_;
    }
    __device__ __forceinline__ uint32_t ridx(uint32_t p) const { return ring + ((p + a) & OMASK); }  // shared address of output byte p

    // Waits for the bulk store in flight: afterwards everything below `flushed` can be re-read from global memory (far
    // back-references, frame checksums) and its ring bytes may be overwritten.
    __device__ __forceinline__ void settle() {
        if (settled != flushed) {
            if (lane == 0) bulk_wait_all();
            __syncwarp();
            settled = flushed;
        }
    }

    // Stores [flushed, upto) to global: ragged head by bytes, 16-byte vector body, and (final only) the tail.
    __device__ __forceinline__ void flush_to(uint32_t upto, bool final) {
        __syncwarp();
        uint32_t q = flushed;
        const uint32_t mis = (q + a) & 15u;
        if (mis != 0 && q < upto) {
            const uint32_t head = min(16u - mis, upto - q);
            if ((uint32_t)lane < head) dst[q + lane] = (uint8_t)lds8(ridx(q + lane));
            q += head;
        }
        const uint32_t vend = q + ((upto - q) & ~15u);
        if (vend > q) {
            // the 16-byte aligned body leaves through the copy engine: at most two spans (the ring wraps), issued by lane 0 once
            // every lane's ring writes are visible to the async proxy and the previous store has completed
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_wait_all();
                const uint32_t r0 = (q + a) & OMASK, bytes = vend - q;
                const uint32_t first = min(bytes, (uint32_t)ORING - r0);
                bulk_store(dst + q, ring + r0, first);
                if (bytes > first) bulk_store(dst + q + first, ring, bytes - first);
                bulk_commit();
            }
            settled = q;   // lane 0 waited: everything before this store is in global memory
        }
        q = vend;
        if (final) {
            if (q + lane < upto) dst[q + lane] = (uint8_t)lds8(ridx(q + lane));
            q = upto;
        }
        flushed = q;
        if (final) settle();
        __syncwarp();
    }
    __device__ __forceinline__ void make_room() {
        if (op - flushed >= FLUSH_T) flush_to(((op + a) & ~15u) - a, false);
    }
    __device__ __forceinline__ void finish() { flush_to(op, true); }

    // Serial path: len literal bytes from global src.
    __device__ __forceinline__ void put_literals(const uint8_t* __restrict__ s, uint32_t len) {
        if (len <= 32 && op - flushed < FLUSH_T) {
            if ((uint32_t)lane < len) sts8(ridx(op + lane), ldg_u8(s + lane));
            op += len;
        } else {
            while (len) {
                uint32_t c = min(len, CHUNK);
                make_room();
                for (uint32_t i = lane; i < c; i += 32) sts8(ridx(op + i), ldg_u8(s + i));
                op += c;
                s += c;
                len -= c;
            }
        }
        __syncwarp();
    }

    // Same, for a source this kernel itself wrote earlier (zstd literal buffer): L2-coherent loads.
    __device__ __forceinline__ void put_literals_coherent(const uint8_t* s, uint32_t len) {
        while (len) {
            uint32_t c = min(len, CHUNK);
            make_room();
            for (uint32_t i = lane; i < c; i += 32) sts8(ridx(op + i), __ldcg(s + i));
            op += c;
            s += c;
            len -= c;
        }
        __syncwarp();
    }

    // Serial path: back-reference copy; caller guarantees 1 <= off <= op.
    __device__ __forceinline__ void put_match(uint32_t off, uint32_t len) {
        if (len <= 32 && off + 32 <= (uint32_t)ORING && op - flushed < FLUSH_T) {
            uint32_t s = op - off;
            uint32_t i = lane;
            if (off < len) i = i % off;
            if ((uint32_t)lane < len) sts8(ridx(op + lane), lds8(ridx(s + i)));
            op += len;
            __syncwarp();
            return;
        }
        while (len) {
            uint32_t c = min(len, CHUNK);
            make_room();
            uint32_t s = op - off;
            if (off + c <= (uint32_t)ORING) {
                if (off >= c) {
                    for (uint32_t i = lane; i < c; i += 32) sts8(ridx(op + i), lds8(ridx(s + i)));
                } else {
                    for (uint32_t i = lane; i < c; i += 32) sts8(ridx(op + i), lds8(ridx(s + i % off)));
                }
            } else {  // far: the source was drained to global long ago (off > ORING - CHUNK >= c)
                if (s + c > settled) settle();
                for (uint32_t i = lane; i < c; i += 32) sts8(ridx(op + i), __ldcg(dst + s + i));
            }
            op += c;
            len -= c;
            __syncwarp();
        }
    }
};

// ------------------------------------------------------------------------------------------------
// SERIAL path: one element per call, warp-uniform, input read straight from global.
// ------------------------------------------------------------------------------------------------

// One LZ4 sequence at ip.  Same end-of-block rules as LZ4_decompress_safe (MFLIMIT 12, LASTLITERALS 5).
__device__ __forceinline__ int lz4_serial_step(const uint8_t* __restrict__ src, uint32_t n, uint32_t cap, uint32_t& ip_io, OutRing& out) {
    uint32_t ip = ip_io;
    if (ip >= n) return CJ_ST_TRUNCATED;
    uint32_t token = ldg_u8(src + ip++);
    uint64_t len = token >> 4;
    if (len == 15) {
        if (n < 15 || ip >= n - 15) return CJ_ST_TRUNCATED;
        uint32_t b;
        do {
            b = ldg_u8(src + ip++);
            len += b;
            if (ip > n - 15) return CJ_ST_TRUNCATED;
        } while (b == 255);
    }
    if ((uint64_t)out.op + len + 12 > cap || (uint64_t)ip + len + 8 > n) {
        // the tail zone of input or output: this must be the final, literal-only sequence
        if ((uint64_t)ip + len != n) return ((uint64_t)ip + len > n) ? CJ_ST_TRUNCATED : CJ_ST_CORRUPT;
        if ((uint64_t)out.op + len > cap) return CJ_ST_DST_SMALL;
        out.put_literals(src + ip, (uint32_t)len);
        ip_io = n;
        return STEP_DONE;
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
            if (ip > n - 4) return CJ_ST_TRUNCATED;
        } while (b == 255);
    }
    len += 4;
    if (off == 0 || off > out.op - out.base) return CJ_ST_OFFSET;
    if ((uint64_t)out.op + len + 5 > cap) return CJ_ST_DST_SMALL;
    out.put_match(off, (uint32_t)len);
    ip_io = ip;
    return STEP_CONT;
}

// One Snappy element at ip (ip < n).
__device__ __forceinline__ int snappy_serial_step(const uint8_t* __restrict__ src, uint32_t n, uint32_t dn, uint32_t& ip_io, OutRing& out) {
    uint32_t ip = ip_io;
    uint32_t tag = ldg_u8(src + ip++);
    uint32_t type = tag & 3;
    if (type == 0) {
        uint64_t len = (tag >> 2) + 1;
        if (len > 60) {
            uint32_t nb = (uint32_t)len - 60;
            if (nb > n - ip) return CJ_ST_TRUNCATED;
            uint32_t v = 0;
            for (uint32_t i = 0; i < nb; i++) v |= ldg_u8(src + ip + i) << (8 * i);
            ip += nb;
            len = (uint64_t)v + 1;
        }
        if (len > n - ip) return CJ_ST_TRUNCATED;
        if (len > dn - out.op) return CJ_ST_LEN_MISMATCH;
        out.put_literals(src + ip, (uint32_t)len);
        ip_io = ip + (uint32_t)len;
        return STEP_CONT;
    }
    uint32_t len, off;
    if (type == 1) {
        if (n - ip < 1) return CJ_ST_TRUNCATED;
        len = 4 + ((tag >> 2) & 7);
        off = ((tag >> 5) << 8) | ldg_u8(src + ip);
        ip += 1;
    } else if (type == 2) {
        if (n - ip < 2) return CJ_ST_TRUNCATED;
        len = 1 + (tag >> 2);
        off = ldg_u8(src + ip) | (ldg_u8(src + ip + 1) << 8);
        ip += 2;
    } else {
        if (n - ip < 4) return CJ_ST_TRUNCATED;
        len = 1 + (tag >> 2);
        off = ldg_u8(src + ip) | (ldg_u8(src + ip + 1) << 8) | (ldg_u8(src + ip + 2) << 16) | (ldg_u8(src + ip + 3) << 24);
        ip += 4;
    }
    if (off == 0 || off > out.op - out.base) return CJ_ST_OFFSET;
    if (len > dn - out.op) return CJ_ST_LEN_MISMATCH;
    out.put_match(off, len);
    ip_io = ip;
    return STEP_CONT;
}

// ------------------------------------------------------------------------------------------------
// LANE-PARALLEL path
// ------------------------------------------------------------------------------------------------
struct InRing {
    uint32_t ring;  // shared-space address of this warp's input ring
    const uint8_t* ring_g;  // generic pointer to the same ring
    const uint8_t* src;
    uint32_t n;
    uint32_t a;       // src misalignment; "g" coordinate = pos + a, ring index = g & IMASK
    uint32_t loaded;  // g coordinate up to which the ring has been filled (multiple of 512)
    uint32_t mbar;    // shared address of this warp's mbarrier (bulk loads complete on it); its phase parity lives in the word behind it
    uint32_t parity;
    int lane;

    // ring_ = this warp's input ring; the warp's mbarrier sits behind the element queue (ring_barrier_init, once per kernel)
    __device__ __forceinline__ void init(uint8_t* ring_, const uint8_t* src_, uint32_t n_, int lane_) {
        ring = smem_addr(ring_);
        ring_g = ring_;
        src = src_;
        n = n_;
        a = (uint32_t)((uintptr_t)src_ & 15u);
        loaded = 0;
        lane = lane_;
        mbar = ring + IRING + QCAP * 4;
        parity = lds32(mbar + 8);
    }
    __device__ __forceinline__ uint32_t byte(uint32_t pos) const { return lds8(ring + ((pos + a) & IMASK)); }

    // Makes [first, first + IRING - 1024) resident (as far as the input goes).  Everything before
    // `first` is dead.  16-byte vector loads; only the ragged first/last group is moved by bytes,
    // so no byte outside [src, src+n) is ever touched.
    __device__ __forceinline__ void refill(uint32_t first) {
        const uint32_t limit = ((first + a) & ~511u) + IRING;
        const uint32_t end_g = (n + a + 15) & ~15u;
        if (loaded + IRING < limit) loaded = limit - IRING;  // the serial path skipped far ahead
        if (loaded < end_g && loaded + 512 <= limit) {
            __syncwarp();
            const uint8_t* base = src - a;  // 16-byte aligned
            uint32_t tx = 0;                // bytes handed to the copy engine in this call
            do {
                const uint32_t g0 = loaded + lane * 16;
                if (loaded >= a && loaded + 512 <= a + n) {
                    // a 512-byte row that lies wholly inside the input: one bulk copy (global -> ring), no lane touches it
                    if (lane == 0) bulk_load(ring + (loaded & IMASK), base + loaded, 512u, mbar);
                    tx += 512;
                } else if (g0 >= a && g0 + 16 <= a + n) {
                    uint4 v = __ldg(reinterpret_cast<const uint4*>(base + g0));
                    sts128(ring + (g0 & IMASK), v);
                } else if (g0 + 16 > a && g0 < a + n) {
                    for (uint32_t t = 0; t < 16; t++) {
                        uint32_t g = g0 + t;
                        if (g >= a && g < a + n) sts8(ring + (g & IMASK), __ldg(base + g));
                    }
                }
                loaded += 512;
            } while (loaded < end_g && loaded + 512 <= limit);
            if (tx) {   // (warp-uniform) wait for the rows: the parser reads them right away
                if (lane == 0) mbar_arrive_expect_tx(mbar, tx);
                mbar_wait(mbar, parity);
                parity ^= 1u;
                if (lane == 0) sts32(mbar + 8, parity);
            }
            __syncwarp();
        }
    }
};

// Once per kernel and warp: the mbarrier of the warp's input ring (one arrival per refill: lane 0's expect_tx) and its parity word.
__device__ __forceinline__ void ring_barrier_init(uint8_t* smem_warp, int lane);

// Compressed size of the element that starts at p if it can go through the lane-parallel path,
// else 0 (needs the serial path: unusual encoding, near the end of the block, or past it).
template <int CODEC>
__device__ __forceinline__ uint32_t elem_size(const InRing& in, uint32_t p) {
    const uint32_t n = in.n;
    const uint32_t b0 = in.byte(p);  // p >= n reads stale ring bytes; the end-of-input tests below turn that into sz = 0
    uint32_t sz;
    if (CODEC == CJ_SNAPPY_RAW) {
        const uint32_t type = b0 & 3, L = b0 >> 2;
        if (type == 0) sz = L < 60 ? L + 2 : 0;  // literals longer than 60 bytes carry their length in extra bytes: serial path
        else sz = type == 1 ? 2 : (type == 2 ? 3 : 0);
        if (p + sz > n) sz = 0;
    } else {
        uint32_t ll = b0 >> 4;
        const uint32_t ml = b0 & 15;
        uint32_t q = p + 1;
        bool cx = false;
        if (ll == 15) {
            uint32_t e = in.byte(q);
            cx = e == 255;
            ll += e;
            q++;
        }
        q += ll + 2;
        if (ml == 15) {
            cx |= in.byte(q) == 255;
            q++;
        }
        sz = q - p;
        // anything within 12 bytes of the end belongs to the serial path (end-of-block rules)
        if (cx || p + sz + 12 > n) sz = 0;
    }
    return sz;
}

// Resolves the element starts among the 32 bytes at pp; appends them to the queue; advances pp.
// stop = the chain ran into a position that needs the serial path (pp then points at it).
template <int CODEC>
__device__ __forceinline__ void parse_window(const InRing& in, uint32_t& pp, uint32_t q, uint32_t& qn, bool& stop, int lane) {
    const uint32_t p = pp + lane;
    const uint32_t sz = elem_size<CODEC>(in, p);
    uint32_t J = sz ? lane + sz : lane;  // where the chain goes from this lane (>= 32: leaves the window; self: stop)
    uint32_t R = 1u << lane;             // element starts visited from this lane
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const uint32_t Jn = __shfl_sync(FULL, J, J & 31), Rn = __shfl_sync(FULL, R, J & 31);
        if (J < 32) {
            R |= Rn;
            J = Jn;
        }
    }
    const uint32_t Jf = __shfl_sync(FULL, J, 0), Rf = __shfl_sync(FULL, R, 0);
    uint32_t tags = Rf;
    if (Jf < 32) {
        stop = true;
        tags = Rf & ((1u << Jf) - 1);
    }
    if ((tags >> lane) & 1) sts32(q + 4 * (qn + __popc(tags & ((1u << lane) - 1))), p);
    qn += __popc(tags);
    pp += Jf;
}

// Lane-parallel copy of up to 16 bytes per lane (nl = 0 for lanes that sit out); src/dst must not overlap.
// The source is a GENERIC pointer: the input ring or the output ring in shared memory, or drained output
// in global memory (re-read through L1/L2 with plain coherent loads) — one instruction stream serves all.
__device__ __forceinline__ void lane_copy16(const uint8_t* gp, uint32_t dp, uint32_t nl) {
    const uint32_t mx = __reduce_max_sync(FULL, nl);
#pragma unroll
    for (uint32_t b = 0; b < 16; b += 4) {  // fully unrolled: constant offsets, warp-uniform early exit
        if (b >= mx) break;
        uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;
        if (b < nl) v0 = gp[b];
        if (b + 1 < nl) v1 = gp[b + 1];
        if (b + 2 < nl) v2 = gp[b + 2];
        if (b + 3 < nl) v3 = gp[b + 3];
        if (b < nl) sts8(dp + b, v0);
        if (b + 1 < nl) sts8(dp + b + 1, v1);
        if (b + 2 < nl) sts8(dp + b + 2, v2);
        if (b + 3 < nl) sts8(dp + b + 3, v3);
    }
}

// How exec_lanes bounds the output: RULE_LEN  = no element may pass `limit` (Snappy announced length, zstd capacity);
//                                    RULE_LZ4  = LZ4_decompress_safe's dstCapacity rules (MFLIMIT 12 / LASTLITERALS 5).
constexpr int RULE_LEN = 0, RULE_LZ4 = 1;

// Executes up to 32 elements, one lane each.  Lane j (< cnt) holds a literal run of LL bytes whose k-th byte is
// lbase[(lidx + k) & lmask] (input ring: lmask = IMASK; a linear buffer: lmask = ~0u) followed by a back-reference
// of ML bytes at distance off (ML == 0: none).  Returns how many elements were executed (a prefix of the lanes);
// element k (if k < cnt) failed a check or did not fit the batch and must be handled by the caller.
template <int RULE, bool EARLY_COPIES>
__device__ __forceinline__ uint32_t exec_lanes(OutRing& out, uint32_t cnt, uint32_t LL, const uint8_t* lbase, uint32_t lidx, uint32_t lmask,
                                               uint32_t ML, uint32_t off, uint32_t limit, int lane) {
    if ((uint32_t)lane >= cnt) { LL = 0; ML = 0; }
    const uint32_t tot = LL + ML;
    uint32_t incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += t;
    }
    const uint32_t O = out.op;
    const uint32_t o = O + incl - tot;
    bool bad = false;
    if ((uint32_t)lane < cnt) {
        bad = incl > TMAX;
        if (ML) bad |= off == 0 || off > o + LL - out.base;
        if (RULE == RULE_LEN) bad |= incl > limit - O;
        else bad |= (o + LL + 12 > limit) || (o + tot + 5 > limit);
    }
    const uint32_t cm = __ballot_sync(FULL, bad);
    const uint32_t k = cm ? __ffs(cm) - 1 : cnt;
    if (k == 0) return 0;
    if ((uint32_t)lane >= k) { LL = 0; ML = 0; }
    const uint32_t T = __shfl_sync(FULL, incl, k - 1);

    // Per-lane addressing.  m = where the match part lands; se = end of the distinct source bytes it needs.
    const uint32_t m = o + LL;
    const uint32_t se = m - off + min(ML, off);
    const uint32_t ldi = (o + out.a) & OMASK, lsi = lidx & lmask;
    const uint32_t mdi = (m + out.a) & OMASK, msi = (m - off + out.a) & OMASK;
    const bool shortL = LL != 0 && LL <= 16 && ldi + LL <= (uint32_t)ORING && (lmask == 0xFFFFFFFFu || lsi + LL <= lmask + 1);
    const bool far = off > FAR_T;                                                              // source re-read from global (L1/L2)
    const bool smallM = ML != 0 && ML <= 16 && off >= ML && mdi + ML <= (uint32_t)ORING && (far || msi + ML <= (uint32_t)ORING);
    const uint8_t* msrc = far ? out.dst + (m - off) : out.ring_g + msi;                          // lane-parallel match source
    bool pending = ML != 0;
    if (__any_sync(FULL, ML != 0 && far && se > out.settled)) out.settle();   // a far source inside the bulk store still in flight (rare)

    // ---- pass 1: literals (always ready); with EARLY_COPIES also every copy of a literal-free lane whose source precedes the batch ----
    {
        const uint8_t* sp = lbase + lsi;
        uint32_t dp = out.ring + ldi, nl = shortL ? LL : 0u;
        if (EARLY_COPIES) {
            const bool early = pending && LL == 0 && smallM && se <= O;
            if (early) { sp = msrc; dp = out.ring + mdi; nl = ML; pending = false; }
        }
        lane_copy16(sp, dp, nl);
        uint32_t lm = __ballot_sync(FULL, LL != 0 && !shortL);
        while (lm) {
            const int j = __ffs(lm) - 1;
            lm &= lm - 1;
            const uint32_t jo = __shfl_sync(FULL, o, j), jl = __shfl_sync(FULL, LL, j), js = __shfl_sync(FULL, lidx, j), jmask = __shfl_sync(FULL, lmask, j);
            const uint8_t* jb = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, (unsigned long long)lbase, j));
            for (uint32_t i = lane; i < jl; i += 32) sts8(out.ridx(jo + i), jb[(js + i) & jmask]);
        }
    }
    __syncwarp();

    // ---- back-references: dependency rounds ----
    {
        uint32_t pm = __ballot_sync(FULL, pending);
        while (pm) {
            const uint32_t F = __shfl_sync(FULL, m, __ffs(pm) - 1);  // all output below F is complete
            const bool ready = pending && se <= F;
            lane_copy16(msrc, out.ring + mdi, (ready && smallM) ? ML : 0u);
            uint32_t lm = __ballot_sync(FULL, ready && !smallM);
            while (lm) {  // long, self-overlapping or ring-wrapping copies: the whole warp moves one at a time
                const int j = __ffs(lm) - 1;
                lm &= lm - 1;
                const uint32_t jm = __shfl_sync(FULL, m, j), jl = __shfl_sync(FULL, ML, j), jf = __shfl_sync(FULL, off, j);
                const uint32_t s = jm - jf;
                if (jf >= jl) {
                    if (jf > FAR_T) for (uint32_t i = lane; i < jl; i += 32) sts8(out.ridx(jm + i), out.dst[s + i]);
                    else for (uint32_t i = lane; i < jl; i += 32) sts8(out.ridx(jm + i), lds8(out.ridx(s + i)));
                } else {
                    for (uint32_t i = lane; i < jl; i += 32) sts8(out.ridx(jm + i), lds8(out.ridx(s + i % jf)));
                }
            }
            pending = pending && !ready;
            __syncwarp();
            pm = __ballot_sync(FULL, pending);
        }
    }
    out.op = O + T;
    return k;
}

// Executes the first cnt (<= 32) queued LZ4 sequences / Snappy elements, one lane each.  Returns how many were
// executed (a prefix); 0 means the first element must go through the serial path.
template <int CODEC>
__device__ __forceinline__ uint32_t exec_batch(const InRing& in, OutRing& out, uint32_t q, uint32_t cnt, uint32_t limit, int lane) {
    __syncwarp();
    uint32_t LL = 0, lsrc = 0, ML = 0, off = 0;
    if ((uint32_t)lane < cnt) {
        const uint32_t p = lds32(q + 4 * lane);
        const uint32_t b0 = in.byte(p);
        if (CODEC == CJ_SNAPPY_RAW) {
            const uint32_t b1 = in.byte(p + 1), b2 = in.byte(p + 2);
            const uint32_t type = b0 & 3, L = b0 >> 2;
            if (type == 0) {
                LL = L + 1;  // the lane-parallel path only sees one-byte literal tags (L < 60)
                lsrc = p + 1;
            } else if (type == 1) {
                ML = 4 + (L & 7);
                off = ((b0 >> 5) << 8) | b1;
            } else {
                ML = L + 1;
                off = b1 | (b2 << 8);
            }
        } else {
            uint32_t ll = b0 >> 4, ml = b0 & 15, r = p + 1;
            if (ll == 15) { ll += in.byte(r); r++; }
            lsrc = r;
            r += ll;
            off = in.byte(r) | (in.byte(r + 1) << 8);
            r += 2;
            if (ml == 15) ml += in.byte(r);
            LL = ll;
            ML = ml + 4;
        }
    }
    return exec_lanes<CODEC == CJ_SNAPPY_RAW ? RULE_LEN : RULE_LZ4, CODEC == CJ_SNAPPY_RAW>(out, cnt, LL, in.ring_g, lsrc + in.a, IMASK, ML, off, limit, lane);
}

// Decodes one compressed stream (an LZ4 block, or the element stream of a Snappy raw block after its
// preamble) starting at src[ip], appending to `out` (which may already hold history).  `limit` is the
// output position the stream may not pass: the capacity for LZ4 (LZ4_decompress_safe's dstCapacity
// rules apply against it), the announced length for Snappy.
template <int CODEC, bool FAST>
__device__ __forceinline__ int32_t decode_stream(const uint8_t* __restrict__ src, uint32_t n, uint32_t ip, uint32_t limit, OutRing& out,
                                                 uint8_t* smem_warp, int lane) {
    int32_t st = CJ_OK;
    if (!FAST) {
        for (;;) {
            if (CODEC == CJ_SNAPPY_RAW && ip >= n) break;
            int r = CODEC == CJ_SNAPPY_RAW ? snappy_serial_step(src, n, limit, ip, out) : lz4_serial_step(src, n, limit, ip, out);
            if (r == STEP_CONT) continue;
            st = r;  // STEP_DONE == CJ_OK
            break;
        }
        return st;
    }
    InRing in;
    in.init(smem_warp + ORING, src, n, lane);
    const uint32_t q = smem_addr(smem_warp + ORING + IRING);
    uint32_t pp = ip, qn = 0, qfront = ip;
    bool stop = false;
    for (;;) {
        in.refill(qfront);
        while (qn < 32 && !stop && pp - qfront < PARSE_SPAN) parse_window<CODEC>(in, pp, q, qn, stop, lane);
        if (qn) {
            const uint32_t cnt = min(qn, 32u);
            out.make_room();
            const uint32_t k = exec_batch<CODEC>(in, out, q, cnt, limit, lane);
            if (k == cnt) {  // whole batch done; keep the leftover queue entries
                const uint32_t rem = qn - cnt;
                uint32_t v = 0;
                if ((uint32_t)lane < rem) v = lds32(q + 4 * (cnt + lane));
                __syncwarp();
                if ((uint32_t)lane < rem) sts32(q + 4 * lane, v);
                qn = rem;
                qfront = rem ? __shfl_sync(FULL, v, 0) : pp;
            } else {  // cut: element k failed a check -> re-parse from it (k == 0: serial path)
                pp = lds32(q + 4 * k);
                __syncwarp();
                qn = 0;
                qfront = pp;
                stop = k == 0;
            }
            out.make_room();
            if (qn || !stop) continue;
        }
        // stop with an empty queue: pp is at an element the lane-parallel path will not take
        if (pp >= n) {
            if (CODEC == CJ_LZ4_BLOCK) st = CJ_ST_TRUNCATED;  // a valid block ends inside the serial step
            break;
        }
        const int r = CODEC == CJ_SNAPPY_RAW ? snappy_serial_step(src, n, limit, pp, out) : lz4_serial_step(src, n, limit, pp, out);
        if (r != STEP_CONT) { st = r; break; }
        qfront = pp;
        stop = false;
    }
    return st;
}

// One independent unit: an LZ4 block or a Snappy raw block.
template <int CODEC, bool FAST>
__device__ int32_t decode_block(const uint8_t* __restrict__ src, uint32_t n, uint8_t* dst, uint32_t cap, uint8_t* smem_warp, int lane,
                                uint32_t* produced) {
    *produced = 0;
    if (n == 0) return CJ_ST_EMPTY;
    uint32_t ip = 0, limit = cap;
    if (CODEC == CJ_SNAPPY_RAW) {
        uint64_t ulen = 0;
        bool done = false;
        for (int i = 0; i < 5 && ip < n; i++) {
            uint32_t b = ldg_u8(src + ip++);
            ulen |= (uint64_t)(b & 0x7f) << (7 * i);
            if (!(b & 0x80)) { done = true; break; }
        }
        if (!done) return CJ_ST_HEADER;
        if (ulen > 0xFFFFFFFFull) return CJ_ST_TOO_BIG;
        if (ulen > cap) return CJ_ST_DST_SMALL;
        limit = (uint32_t)ulen;
    } else {
        if (cap == 0) return (n == 1 && ldg_u8(src) == 0) ? CJ_OK : CJ_ST_DST_SMALL;
    }
    OutRing out;
    out.init(smem_warp, dst, lane);
    int32_t st = decode_stream<CODEC, FAST>(src, n, ip, limit, out, smem_warp, lane);
    if (CODEC == CJ_SNAPPY_RAW && st == CJ_OK && out.op != limit) st = CJ_ST_LEN_MISMATCH;
    out.finish();
    *produced = out.op;
    return st;
}

constexpr int DEC_WARPS = 4;
constexpr int DEC_SMEM_WARP = ORING + IRING + QCAP * 4 + 16;   // output ring | input ring | element queue | mbarrier + parity word

__device__ __forceinline__ void ring_barrier_init(uint8_t* smem_warp, int lane) {
    const uint32_t mbar = smem_addr(smem_warp + ORING + IRING + QCAP * 4);
    if (lane == 0) {
        mbar_init(mbar, 1);
        sts32(mbar + 8, 0);
    }
    __syncwarp();
}

}  // namespace cj
