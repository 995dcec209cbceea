// lz_decode4.cu — generation-4 batch decode of Snappy raw blocks and LZ4 blocks (sm_100a): one THREAD per block.
//
// Same reference entry points as lz_decode.cuh (snap::raw::Decoder::decompress behind cramjam.snappy.decompress_raw /
// decompress_raw_into, src/snappy.rs:52-60,102-108; LZ4_decompress_safe behind lz4::block::decompress_into,
// src/lz4.rs:78-95,140-173).  Generations 2 and 3 spend 20-30 warp instructions per ~8-byte element because a whole warp
// cooperates on one block; here a lane owns a block and decodes it the way a CPU does, so one warp instruction advances
// 32 blocks.  What makes that workable on a GPU:
//
//   * compressed input reaches a lane through its own 128-byte shared-memory ring, filled 16 bytes per sub-iteration by
//     the lane's own cp.async (no cooperation, no shuffles), 64-80 bytes ahead of the read cursor;
//   * every sub-iteration a lane ISSUES at most one 8-byte CHUNK of its current element and RETIRES the chunk it issued
//     D sub-iterations earlier.  Issue starts the source fetch: literal bytes are in the input ring already, a
//     back-reference further than 64 bytes is fetched from the block's own output in global memory by cp.async into a
//     per-lane staging slot (asynchronous: no register scoreboard is held across the loop back-edge).  Retire reads 16
//     bytes from shared memory, shifts them to the output alignment, appends them to a 16-byte register accumulator and
//     stores a completed word with one aligned 16-byte st.global;
//   * back-references that may reach into bytes not yet stored (offset <= 64) are read at retire time from a 128-byte
//     per-lane mirror of the most recent output in shared memory; offsets below 8 are expanded to a periodic pattern;
//   * rings, mirror and staging slots are interleaved across lanes in 16-byte granules, so lane-private accesses at
//     unrelated positions fall into different banks;
//   * the common path has no branches (selects, predicated PTX); everything unusual takes one slow branch;
//   * one CTA per SM with as many warps as the batch needs and the smallest shared-memory carve-out that holds it: the
//     L1 share of the SM decides the kernel's speed (re-reads of back-reference sectors hit it).
//   Anything unusual (unaligned unit, 4-byte-offset copy, malformed element, bad offset, length mismatch) puts the
//   block on the redo list of the generation-2 kernel, which owns all error reporting: status codes stay the oracle's.
#include "internal.h"
#include "lz_decode.cuh"

namespace cj {

#ifndef CJ_G4_D
#define CJ_G4_D 3
#endif
constexpr int G4_MAX_WARPS = 20;       // warps per CTA (96 registers x 640 threads fill the register file): chosen at launch, see launch_g4
constexpr int G4_DMAX = 4;             // sub-iterations between issue and retire of a chunk: template parameter D <= G4_DMAX
constexpr uint32_t G4_INB = 128;       // input ring bytes per lane (8 granules of 16 bytes)
constexpr uint32_t G4_RECB = 128;      // recent-output mirror bytes per lane
constexpr uint32_t G4_NEAR = 64;       // back-references up to this offset are read from the mirror at retire time
constexpr int g4_smem_warp(int D) { return 32 * (int)(G4_INB + G4_RECB + 16 * D); }   // + one 16-byte staging slot per chunk in flight
constexpr int g4_smem_cta(int D, int warps) { return g4_smem_warp(D) * warps + 1024; }   // + the 256-entry tag table
constexpr uint32_t G4_MAX = 1u << 30;
static_assert(8 * (G4_DMAX - 1) + 15 + 8 <= G4_NEAR + 1, "a far back-reference must lie entirely below the stored frontier");

struct G4 {
    uint32_t* redo_list;   // units for the generation-2 kernel
    unsigned* ctr;         // [1] redo count  [2] redo work queue
};

__device__ __forceinline__ void g4_redo(const G4& g, uint32_t u) {
    const unsigned i = atomicAdd(&g.ctr[1], 1u);
    g.redo_list[i] = u;
}

// predicated forms (no branch around them)
//
// Ordering of the far fetch.  A far back-reference is read with cp.async.ca (LDGSTS, L1-allocating) from the lane's OWN
// output, which the same lane wrote earlier with plain st.global.v4.  What is relied on, and why it holds:
//   (1) the bytes were stored by an instruction that precedes the fetch in program order by at least one whole
//       sub-iteration (static_assert below: a far source lies entirely below the stored frontier), and no other thread
//       ever writes or reads this block's output while the kernel runs;
//   (2) the store and the fetch are issued by one warp through one LSU queue in program order; both are `asm volatile`
//       with a "memory" clobber, so neither the compiler nor ptxas moves one across the other;
//   (3) the L1 is write-through and a store that hits a resident line updates (or evicts) it — the same property that
//       makes an ordinary ld.global after st.global by the same thread return the stored value — and LDGSTS performs the
//       same L1 lookup as LDG; a fetch that misses goes to L2 behind the store on the same path.
// The PTX memory model words the guarantee for ld/st; for the asynchronous copy we rely on (2) and (3) as properties of the
// sm_100 memory pipeline.  Evidence: every decode test compares against the oracle bit for bit (including 131 072 blocks
// per launch, tests/test_gpu_lz_decode4.py::test_multi_round_batch_configs4_size), bench.py verifies 4 GiB per run, and
// compute-sanitizer memcheck / racecheck report nothing (profiles/r01_sanitizers.md).  (cp.async.cg — L2 only, no L1 line
// that could go stale — exists for 16-byte copies only and would give up the L1 hits on the re-reads of a back-reference's
// sector, which are what this kernel's speed hangs on)
__device__ __forceinline__ void g4_cp_async8_if(uint32_t saddr, const void* gptr, uint32_t pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 8;\n\t}" ::"r"(saddr), "l"(gptr), "r"(pred) : "memory");
}
__device__ __forceinline__ void g4_cp_async16_if(uint32_t saddr, const void* gptr, uint32_t pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(saddr), "l"(gptr), "r"(pred) : "memory");
}
__device__ __forceinline__ void g4_sts128(uint32_t saddr, uint32_t x, uint32_t y, uint32_t z, uint32_t w, uint32_t pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n\t}" ::"r"(saddr), "r"(x), "r"(y), "r"(z), "r"(w), "r"(pred) : "memory");
}
__device__ __forceinline__ void g4_stg128_if(void* p, uint32_t x, uint32_t y, uint32_t z, uint32_t w, uint32_t pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.global.v4.u32 [%0], {%1,%2,%3,%4};\n\t}" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w), "r"(pred) : "memory");
}
__device__ __forceinline__ uint2 g4_lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

// 8 bytes starting `shb` (0..7) bytes into the 16-byte window {a0, a1}
__device__ __forceinline__ uint64_t g4_funnel(uint2 a0, uint2 a1, uint32_t shb) {
    uint32_t w0 = a0.x, w1 = a0.y, w2 = a1.x;
    if (shb & 4u) { w0 = w1; w1 = w2; w2 = a1.y; }
    const uint32_t s = (shb & 3u) * 8;
    const uint32_t r0 = __funnelshift_r(w0, w1, s), r1 = __funnelshift_r(w1, w2, s);
    return ((uint64_t)r1 << 32) | r0;
}

// Snappy tag table: [6:0] compressed size of the element, [14:8] bytes it produces, [23:22] kind, [21:16] field,
// bit 31 = not a plain element (literal with length bytes, 4-byte-offset copy).
constexpr uint32_t G4_LIT = 0, G4_M16 = 1, G4_M1 = 2;
constexpr uint32_t G4_CK_LIT = 1, G4_CK_NEAR = 2, G4_CK_FAR = 3;   // where a chunk's source bytes are read from when it retires
__device__ __forceinline__ uint32_t g4_tag_entry(uint32_t tag) {
    const uint32_t type = tag & 3, L = (tag >> 2) + 1;
    if (type == 0) return L <= 60 ? ((1 + L) | (L << 8) | (((G4_LIT << 6) | (L - 1)) << 16)) : 0x80000000u;
    if (type == 1) {
        const uint32_t len = 4 + ((tag >> 2) & 7);
        return 2 | (len << 8) | (((G4_M1 << 6) | (tag >> 5)) << 16);
    }
    if (type == 2) return 3 | (L << 8) | (((G4_M16 << 6) | (L - 1)) << 16);
    return 0x80000000u;
}

template <int CODEC, int G4_D>
__global__ void __launch_bounds__(G4_MAX_WARPS * 32, 1) g4_kernel(Batch b, G4 g) {
    constexpr int G4_SMEM_WARP = g4_smem_warp(G4_D);
    const int G4_WARPS = blockDim.x >> 5;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t wbase = smem_addr(smem + (size_t)warp * G4_SMEM_WARP);
    const uint32_t inl = wbase + lane * 16;                               // granule q of this lane's input ring: inl + q * 512
    const uint32_t recl = wbase + 32 * G4_INB + lane * 16;                // 16-byte slot q of the output mirror: recl + q * 512
    const uint32_t stl = wbase + 32 * (G4_INB + G4_RECB) + lane * 16;     // staging slot u of a far chunk: stl + u * 512
    const uint32_t lut = smem_addr(smem + (size_t)G4_SMEM_WARP * G4_WARPS);
    if (CODEC == CJ_SNAPPY_RAW) {
        for (uint32_t t = threadIdx.x; t < 256; t += blockDim.x) sts32(lut + 4 * t, g4_tag_entry(t));
        __syncthreads();
    }
    auto in_a = [&](uint32_t x) -> uint32_t { return inl + ((x & (G4_INB - 16)) << 5) + (x & 15u); };
    auto rec_s = [&](uint32_t p) -> uint32_t { return recl + ((p & 0x70u) << 5); };             // mirror slot (16 bytes) holding output byte p
    auto rec_a = [&](uint32_t p) -> uint32_t { return recl + ((p & 0x70u) << 5) + (p & 8u); };   // its 8-byte half
    const uint32_t nwarps = gridDim.x * G4_WARPS;

    for (uint32_t first = (blockIdx.x * G4_WARPS + warp) * 32; first < b.n; first += nwarps * 32) {
        // ---- every lane takes one block ----
        const uint32_t cur = first + lane;
        bool active = false;
        const uint8_t* src = nullptr;
        uint8_t* dst = nullptr;
        uint32_t n = 0, ulen = 0, ip = 0;
        if (cur < b.n) {
            const uint64_t sl = b.src_len[cur], dcap = b.dst_cap[cur];
            src = b.src_base + b.src_off[cur];
            dst = b.dst_base + b.dst_off[cur];
            bool ok = sl >= 1 && sl <= G4_MAX && (((uintptr_t)src | (uintptr_t)dst) & 15u) == 0;
            if (ok) {
                n = (uint32_t)sl;
                if constexpr (CODEC == CJ_SNAPPY_RAW) {
                    uint64_t v = 0;
                    bool done = false;
                    for (int i = 0; i < 5 && ip < n; i++) {
                        const uint32_t x = ldg_u8(src + ip++);
                        v |= (uint64_t)(x & 0x7f) << (7 * i);
                        if (!(x & 0x80)) { done = true; break; }
                    }
                    ok = done && v >= 1 && v <= dcap && v <= G4_MAX;
                    ulen = (uint32_t)v;
                } else {
                    ok = dcap >= 1 && dcap <= G4_MAX;
                    ulen = (uint32_t)dcap;   // LZ4: the capacity; LZ4_decompress_safe's rules are applied against it
                }
            }
            if (ok) active = true;
            else g4_redo(g, cur);
        }
        const uint32_t n16 = n & ~15u;   // the ragged last granule is read straight from global memory
        uint32_t loaded = 0, lim = 0;    // input granules requested / bytes known to have arrived in the ring
        uint32_t opi = 0, opr = 0;       // output position of the next chunk to issue / to retire
        uint32_t rem = 0, sp = 0;        // bytes of the current element still to issue; literal: input position, copy: offset
        bool is_lit = false, fin = false;
        bool lz_phase = false, lz_last = false;   // LZ4: the next thing to decode is an offset (not a token); the final sequence has been seen
        uint32_t lz_ml = 0;                       // LZ4: match-length nibble of the token whose literals are being issued
        uint64_t lo = 0, hi = 0;         // output bytes [opr & ~15, opr)
        uint32_t tw0 = 0, tw1 = 0;       // the two ring words around ip, fetched one sub-iteration ahead
        bool tw_ok = false;              // ... and whether they had arrived in the ring when they were fetched
        uint32_t M[G4_D], P[G4_D];       // chunk in flight: M = bytes | kind << 4;  P = input position / offset / byte shift
        uint32_t L[G4_D];                // `loaded` as it was at the end of the slot's previous visit
#pragma unroll
        for (int u = 0; u < G4_D; u++) { M[u] = 0; P[u] = 0; L[u] = 0; }

        // The loop body is written without branches on the common path (selects and predicated PTX), so that the retire
        // chain, the tag-decode chain and the issue of the next chunk interleave inside one basic block: with one thread
        // per block only ~3.5 warps share a scheduler, and instruction-level parallelism has to hide what they cannot.
        while (__any_sync(FULL, active)) {
#pragma unroll
            for (int u = 0; u < G4_D; u++) {
                asm volatile("cp.async.wait_group %0;" ::"n"(G4_D - 1) : "memory");
                lim = L[u];   // what had been requested when this slot was last visited has arrived now
                // Shared-memory accesses are volatile asm and keep their program order, so the loads of the two independent
                // chains (retire, tag decode) are written first and their arithmetic afterwards.
                // ---- [A] retire: load the source window of the chunk issued G4_D sub-iterations ago (in shared memory by now) ----
                const uint32_t m = M[u], rc = m & 15u, rkind = m >> 4, rp_ = P[u];
                const uint32_t rs = rkind == G4_CK_LIT ? rp_ : opr - rp_;
                const uint32_t rs0 = rs & ~7u, rs1 = rs0 + 8;
                const uint32_t rsa = stl + u * 512;
                // input ring and mirror have the same geometry (8 granules of 16 bytes, 512 bytes apart): one address form, two bases
                static_assert(G4_INB == G4_RECB, "retire addresses assume equal ring geometry");
                const uint32_t rbase = rkind == G4_CK_LIT ? inl : recl;
                const uint32_t ra = rkind == G4_CK_FAR ? rsa : rbase + ((rs0 & 0x70u) << 5) + (rs0 & 8u);
                const uint32_t ra2 = rkind == G4_CK_FAR ? rsa + 8 : rbase + ((rs1 & 0x70u) << 5) + (rs1 & 8u);
                // the mirror always holds the completed 16-byte words; a near chunk also needs the bytes still in the accumulator
                g4_sts128(rec_s(opr), (uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32), rkind == G4_CK_NEAR ? 1u : 0u);
                const uint2 a0 = g4_lds64(ra), a1 = g4_lds64(ra2);
                // ---- [B] tag decode: table entry of the tag fetched at the end of the previous sub-iteration ----
                const uint32_t t = __funnelshift_r(tw0, tw1, (ip & 3u) * 8);
                uint32_t ent = 0;
                if constexpr (CODEC == CJ_SNAPPY_RAW) ent = lds32(lut + 4 * (t & 255u));
                // ---- [C] retire: align, append to the accumulator, store ----
                {
                    const uint32_t shb = rkind == G4_CK_FAR ? rp_ : (rs & 7u);
                    uint64_t v = g4_funnel(a0, a1, shb);
                    if (rkind == G4_CK_NEAR && rp_ < 8) {   // overlapping copy: the `rp_` bytes repeat
                        v &= ~0ull >> (64 - 8 * rp_);
                        v |= v << (8 * rp_);
                        if (rp_ < 4) v |= v << (16 * rp_);
                        if (rp_ < 2) v |= v << 32;
                    }
                    {   // keep the chunk's rc bytes: two clamped funnel shifts build the 32-bit halves of the mask (rc = 0 gives 0)
                        const uint32_t mlo = __funnelshift_rc(0xFFFFFFFFu, 0u, 32u - 8u * min(rc, 4u));
                        const uint32_t mhi = __funnelshift_rc(0xFFFFFFFFu, 0u, 64u - 8u * max(rc, 4u));
                        v &= ((uint64_t)mhi << 32) | mlo;
                    }
                    const uint32_t k = opr & 15u, sh = (k & 7u) * 8;
                    const uint64_t vl = v << sh, vh = sh ? v >> (64 - sh) : 0ull;
                    const bool lowhalf = k < 8;
                    lo |= lowhalf ? vl : 0ull;
                    hi |= lowhalf ? vh : vl;
                    const uint32_t cross = k + rc >= 16 ? 1u : 0u;   // the 16-byte word is complete (only from the upper half)
                    g4_sts128(rec_s(opr), (uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32), cross);
                    g4_stg128_if(dst + (opr & ~15u), (uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32), cross);
                    lo = cross ? vh : lo;
                    hi = cross ? 0ull : hi;
                    opr += rc;
                }
                // ---- [D] decode the next element if the current one is fully issued (plain tags that are in the ring) ----
                const bool need = active && rem == 0 && !fin;
                const bool atend = ip >= n;
                bool slow;
                if constexpr (CODEC == CJ_SNAPPY_RAW) {
                    fin = fin || (need && atend);
                    const bool fast = need && !atend && tw_ok;
                    const uint32_t adv = ent & 0x7Fu, len = (ent >> 8) & 0x7Fu, kind = (ent >> 22) & 3u;
                    const uint32_t off = kind == G4_M1 ? (((ent >> 16) & 7u) << 8) | ((t >> 8) & 0xFFu) : (t >> 8) & 0xFFFFu;
                    const bool lit = kind == G4_LIT;
                    const bool bad = (ent >> 31) != 0 || adv > n - ip || len > ulen - opi || (!lit && (off == 0 || off > opi));
                    const bool take = fast && !bad;
                    slow = (fast && bad) || (need && !atend && !fast && ip + 4 > n16);   // rare tag, failed check, or the block's last bytes
                    is_lit = take ? lit : is_lit;
                    sp = take ? (lit ? ip + 1 : off) : sp;
                    rem = take ? len : rem;
                    ip = take ? ip + adv : ip;
                } else {
                    // LZ4: a sequence is decoded in two steps, its token (-> the literal run) and, once the literals are issued, its
                    // offset (-> the match); a token without literals does both at once.  One length-extension byte is taken here,
                    // longer runs and the end-of-block zone (MFLIMIT / LASTLITERALS rules of lz4_serial_step) go to the slow path.
                    fin = fin || (need && atend && lz_last);
                    const bool fast = need && !atend && tw_ok;
                    const uint32_t b0 = t & 255u, b1 = (t >> 8) & 255u;
                    const uint32_t ll = b0 >> 4, llx = ll == 15u ? 1u : 0u;
                    const uint32_t ll_tot = ll + (llx ? b1 : 0u), q = ip + 1 + llx;
                    const bool lit0 = !lz_phase && ll_tot != 0;
                    const uint32_t o8 = lz_phase ? 0u : 8u;                       // the offset sits at ip (after literals) or at ip + 1 (no literals)
                    const uint32_t mln = lz_phase ? lz_ml : (b0 & 15u);
                    const uint32_t offv = (t >> o8) & 0xFFFFu, xb = (t >> (o8 + 16)) & 255u;
                    const uint32_t mlx = mln == 15u ? 1u : 0u;
                    const uint32_t mtot = mln + 4 + (mlx ? xb : 0u), madv = (o8 >> 3) + 2 + mlx;
                    const bool longrun = lit0 ? (llx && b1 == 255u) : (mlx && xb == 255u);
                    const bool tailz = !lz_phase && ((uint64_t)opi + ll_tot + 12 > ulen || (uint64_t)q + ll_tot + 8 > n);
                    const bool mbad = !lit0 && (offv == 0 || offv > opi || (uint64_t)opi + mtot + 5 > ulen);
                    const bool bad = longrun || tailz || mbad;
                    const bool take = fast && !bad;
                    slow = (fast && bad) || (need && !atend && !fast && ip + 4 > n16) || (need && atend && !lz_last);
                    is_lit = take ? lit0 : is_lit;
                    sp = take ? (lit0 ? q : offv) : sp;
                    rem = take ? (lit0 ? ll_tot : mtot) : rem;
                    ip = take ? (lit0 ? q + ll_tot : ip + madv) : ip;
                    lz_ml = (take && lit0) ? (b0 & 15u) : lz_ml;
                    lz_phase = take ? lit0 : lz_phase;
                }
                // ---- issue one chunk of the current element ----
                {
                    uint32_t c = min(rem, 8u);
                    const bool litok = sp + c <= lim;
                    const bool isnear = !is_lit && sp <= G4_NEAR;
                    const bool isfar = !is_lit && !isnear && c != 0;
                    slow = slow || (is_lit && c != 0 && !litok && sp + c > n16);   // literal bytes that never enter the ring
                    c = (is_lit && !litok) ? 0u : c;
                    const uint32_t fs = opi - sp;
                    const uint8_t* gp = dst + (fs & ~7u);
                    g4_cp_async8_if(stl + u * 512, gp, isfar ? 1u : 0u);
                    g4_cp_async8_if(stl + u * 512 + 8, gp + 8, (isfar && (fs & 7u) + c > 8) ? 1u : 0u);   // second half only if the chunk reaches into it
                    const uint32_t kind = is_lit ? G4_CK_LIT : (isnear ? G4_CK_NEAR : G4_CK_FAR);
                    M[u] = c ? (c | (kind << 4)) : 0u;
                    P[u] = (is_lit || isnear) ? sp : (fs & 7u);
                    sp += is_lit ? c : 0u;
                    opi += c;
                    rem -= c;
                }
                // ---- everything unusual, at most a few times per block ----
                if (slow) {
                    bool fail = false;
                    if (rem == 0 && CODEC == CJ_SNAPPY_RAW) {   // tag decode from global memory, with every check
                        uint32_t t = ldg_u8(src + ip);
                        if (ip + 1 < n) t |= ldg_u8(src + ip + 1) << 8;
                        if (ip + 2 < n) t |= ldg_u8(src + ip + 2) << 16;
                        const uint32_t tag = t & 255u;
                        const uint32_t ent = g4_tag_entry(tag);
                        if (ent >> 31) {
                            const uint32_t nb = (tag >> 2) - 59;
                            if ((tag & 3u) != 0 || nb > n - ip - 1) fail = true;   // 4-byte-offset copies: generation 2
                            else {
                                uint32_t v = 0;
                                for (uint32_t i = 0; i < nb; i++) v |= ldg_u8(src + ip + 1 + i) << (8 * i);
                                const uint64_t LL = (uint64_t)v + 1;
                                const uint32_t q = ip + 1 + nb;
                                if (LL > n - q || LL > ulen - opi) fail = true;
                                else { is_lit = true; sp = q; rem = (uint32_t)LL; ip = q + (uint32_t)LL; }
                            }
                        } else {
                            const uint32_t adv = ent & 0x7Fu, len = (ent >> 8) & 0x7Fu, kind = (ent >> 22) & 3u;
                            const uint32_t off = kind == G4_M1 ? (((ent >> 16) & 7u) << 8) | ((t >> 8) & 0xFFu) : (t >> 8) & 0xFFFFu;
                            const bool lit = kind == G4_LIT;
                            if (adv > n - ip || len > ulen - opi || (!lit && (off == 0 || off > opi))) fail = true;
                            else { is_lit = lit; sp = lit ? ip + 1 : off; rem = len; ip += adv; }
                        }
                    } else if (rem == 0) {   // LZ4, step by step as lz4_serial_step (lz_decode.cuh) does it; whatever it rejects is generation 2's
                        if (ip >= n) fail = true;
                        else if (!lz_phase) {
                            const uint32_t token = ldg_u8(src + ip);
                            uint32_t p = ip + 1;
                            uint64_t len = token >> 4;
                            if (len == 15) {
                                if (n < 15 || p >= n - 15) fail = true;
                                else {
                                    uint32_t bb;
                                    do {
                                        bb = ldg_u8(src + p++);
                                        len += bb;
                                        if (p > n - 15) { fail = true; break; }
                                    } while (bb == 255);
                                }
                            }
                            if (!fail) {
                                if ((uint64_t)opi + len + 12 > ulen || (uint64_t)p + len + 8 > n) {   // must be the final, literal-only sequence
                                    if ((uint64_t)p + len != n || (uint64_t)opi + len > ulen) fail = true;
                                    else { is_lit = true; sp = p; rem = (uint32_t)len; ip = n; lz_last = true; }
                                } else {
                                    is_lit = true; sp = p; rem = (uint32_t)len; ip = p + (uint32_t)len;
                                    lz_phase = true; lz_ml = token & 15u;
                                }
                            }
                        } else {
                            const uint32_t off = ldg_u8(src + ip) | (ldg_u8(src + ip + 1) << 8);
                            uint32_t p = ip + 2;
                            uint64_t len = lz_ml;
                            if (len == 15) {
                                uint32_t bb;
                                do {
                                    bb = ldg_u8(src + p++);
                                    len += bb;
                                    if (p > n - 4) { fail = true; break; }
                                } while (bb == 255);
                            }
                            len += 4;
                            if (fail || off == 0 || off > opi || (uint64_t)opi + len + 5 > ulen) fail = true;
                            else { is_lit = false; sp = off; rem = (uint32_t)len; ip = p; lz_phase = false; }
                        }
                    } else {          // a literal chunk at the end of the block: staged by hand, retired like a far chunk
                        const uint32_t c = min(rem, 8u);
                        uint32_t x0 = 0, x1 = 0;
                        for (uint32_t j = 0; j < c; j++) {
                            const uint32_t bb = ldg_u8(src + sp + j);
                            if (j < 4) x0 |= bb << (8 * j);
                            else x1 |= bb << (8 * (j - 4));
                        }
                        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(stl + u * 512), "r"(x0), "r"(x1) : "memory");
                        M[u] = c | (G4_CK_FAR << 4);
                        P[u] = 0;
                        sp += c;
                        opi += c;
                        rem -= c;
                    }
                    if (fail) {
                        g4_redo(g, cur);
                        active = false;
                        rem = 0;
                        fin = true;
                    }
                }
                // ---- fetch the tag words of the next element; their latency overlaps the loop bookkeeping ----
                tw0 = lds32(in_a(ip & ~3u));
                tw1 = lds32(in_a((ip & ~3u) + 4));
                tw_ok = ip + 4 <= lim;
                // ---- input ring: one more granule if it fits ahead of everything still needed ----
                {
                    const uint32_t rp = (rem && is_lit) ? sp : ip;
                    const uint32_t keep = (rp > 16u * G4_D ? rp - 16u * G4_D : 0u) & ~15u;   // chunks in flight read at most this far back
                    const bool go = active && loaded < n16 && loaded <= keep + (G4_INB - 16);
                    g4_cp_async16_if(inl + ((loaded & (G4_INB - 16)) << 5), src + loaded, go ? 1u : 0u);
                    loaded += go ? 16u : 0u;
                    L[u] = loaded;
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            // ---- a lane is done when its input is consumed and every chunk has retired ----
            if (active && fin) {
                bool empty = true;
#pragma unroll
                for (int u = 0; u < G4_D; u++) empty = empty && M[u] == 0;
                if (empty) {
                    if (CODEC == CJ_SNAPPY_RAW && opi != ulen) g4_redo(g, cur);
                    else {
                        const uint32_t k = opr & 15u;
                        uint8_t* tail = dst + (opr & ~15u);
                        for (uint32_t j = 0; j < k; j++) tail[j] = (uint8_t)((j < 8 ? lo >> (8 * j) : hi >> (8 * (j - 8))));
                        b.dst_len[cur] = opi;
                        b.status[cur] = CJ_OK;
                    }
                    active = false;
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
}

cudaError_t launch_lz_decode_list(int codec, const Batch& b, uint32_t* redo_list, unsigned* ctr, int sm_count, cudaStream_t stream);

template <int CODEC, int D>
static cudaError_t launch_g4(const Batch& b, const G4& g, int sm_count, cudaStream_t stream) {
    static cj_per_device_flag attr_flag;
    int& attr_done = attr_flag.here();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(g4_kernel<CODEC, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, g4_smem_cta(D, G4_MAX_WARPS));
        if (e != cudaSuccess) return e;
        attr_done = 1;
    }
    // One CTA per SM, as many warps as the batch needs (a lane per block): the kernel then takes an even share of every
    // SM, and whatever runs beside it (the warp-per-block kernel of the co-scheduled split) finds room on all of them.
    const size_t warps = ((size_t)b.n + 31) / 32;
    static const int force_w = [] { const char* e = getenv("CJ_G4_WARPS"); return e ? atoi(e) : 0; }();       // experiments: warps per CTA
    // batches beyond sm_count x G4_MAX_WARPS x 32 blocks are decoded in several equal rounds by the same CTAs
    const size_t rounds = std::max<size_t>(1, (warps + (size_t)sm_count * G4_MAX_WARPS - 1) / ((size_t)sm_count * G4_MAX_WARPS));
    int w = (int)std::min<size_t>(G4_MAX_WARPS, std::max<size_t>(1, (warps + (size_t)sm_count * rounds - 1) / ((size_t)sm_count * rounds)));
    if (force_w >= 1 && force_w <= G4_MAX_WARPS) w = force_w;
    const int grid = (int)std::min<size_t>((warps + w - 1) / w, (size_t)sm_count * (size_t)std::max(1, G4_MAX_WARPS / w));
    const size_t smem = (size_t)g4_smem_cta(D, w);
    // The L1 share of the 256 KB SM memory decides this kernel's speed (re-reads of back-reference sectors hit it): ask for the
    // smallest shared-memory carve-out that holds the CTAs of one SM (8.8 ms with 60 KB of L1, 14.4 ms with 28 KB).
    static cj_per_device_flag carve_flag;   // warps per CTA the carve-out was last set for on this device
    int& carve_for_w = carve_flag.here();
    if (carve_for_w != w) {
        const int per_sm = (grid + sm_count - 1) / sm_count;
        const int pct = (int)std::min<size_t>(100, ((smem + 1024) * per_sm * 100 + 228 * 1024 - 1) / (228 * 1024));
        cudaError_t e = cudaFuncSetAttribute(g4_kernel<CODEC, D>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        if (e != cudaSuccess) return e;
        carve_for_w = w;
    }
    g4_kernel<CODEC, D><<<grid, w * 32, smem, stream>>>(b, g);
    return cudaGetLastError();
}

cudaError_t launch_lz_decode4(int codec, const Batch& b, LzScratch& sc, int sm_count, cudaStream_t stream) {
    static const int depth = [] { const char* e = getenv("CJ_G4_D"); return e ? atoi(e) : CJ_G4_D; }();
    const size_t n = b.n;
    if (sc.ensure_fixed((n + 8) * 4 + 64) != 0) return cudaErrorMemoryAllocation;
    G4 g;
    g.ctr = (unsigned*)sc.fixed();
    g.redo_list = (uint32_t*)sc.fixed() + 8;
    cudaError_t e = cudaMemsetAsync(g.ctr, 0, 4 * sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    if (codec == CJ_LZ4_BLOCK) e = depth <= 2 ? launch_g4<CJ_LZ4_BLOCK, 2>(b, g, sm_count, stream) : launch_g4<CJ_LZ4_BLOCK, 3>(b, g, sm_count, stream);
    else e = depth <= 2 ? launch_g4<CJ_SNAPPY_RAW, 2>(b, g, sm_count, stream) : (depth == 3 ? launch_g4<CJ_SNAPPY_RAW, 3>(b, g, sm_count, stream) : launch_g4<CJ_SNAPPY_RAW, 4>(b, g, sm_count, stream));
    if (e != cudaSuccess) return e;
    return launch_lz_decode_list(codec, b, g.redo_list, g.ctr, sm_count, stream);
}

}  // namespace cj
