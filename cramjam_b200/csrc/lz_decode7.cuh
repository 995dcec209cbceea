// lz_decode7.cuh — generation-7 batch decode of Snappy raw blocks and LZ4 blocks: the LANE PROGRAM, shared by the device
// kernel (lz_decode7.cu) and the host-side emulation that the CPU tests run (tests/emu/g7_emu.cpp).
//
// Reference entry points: snap::raw::Decoder::decompress behind cramjam.snappy.decompress_raw / decompress_raw_into
// (src/snappy.rs:52-60,102-108) and LZ4_decompress_safe behind lz4::block::decompress_into (src/lz4.rs:78-95,140-173).
//
// One THREAD per block, like generation 4 (lz_decode4.cu), rebuilt around what ncu showed of it and of this kernel's first
// version (profiles/r02b_g7_*, profiles/README.md): with 65 536 lanes at unrelated positions the kernel is bound by the SM's L1 / shared-memory
// data pipe — one wavefront per cycle, and a scattered 4-byte or 16-byte access of a warp costs a wavefront per bank conflict
// or per cache line — long before it is bound by instruction issue.  So every access is made as wide and as conflict-free as
// the format allows:
//
//   * CHUNKS are up to 16 bytes (9 900 instead of 12 600 iterations per 64 KiB block of the bench corpus);
//   * every per-lane structure in shared memory is a ring of 16-byte GRANULES, interleaved across the lanes of the warp
//     (granule q of lane l at q * 512 + l * 16): a 16-byte access of a quarter warp then touches each bank once, whatever
//     granule each lane is at.  The input ring has 8 granules (filled by 16-byte cp.async, two per pass), the ring of recent
//     output 8, and every chunk in flight owns a staging pair;
//   * a chunk's SOURCE is always two granules, read with two ld.shared.v4 into eight registers when the chunk retires: from the
//     input ring (literal bytes), from the output ring (back-references of at most 96 bytes), or from the chunk's staging pair,
//     which one or two 16-byte cp.async filled from the block's own output in global memory one pass earlier (everything
//     further back).  A two-stage select picks the six words the chunk starts in, five funnel shifts align them to the output
//     phase.  (Fetching far sources with ld.global straight into the registers was measured and is slower: the loads of
//     different slots share hardware scoreboards, so waiting for the oldest waits for the newest, profiles/README.md.)
//   * the OUTPUT is assembled in registers: a second two-stage select places the five words at the output position inside a
//     32-byte window whose lower half is the granule being filled; a finished granule waits for its sector partner and the two
//     leave with one 32-byte st.global.v8 (the kernel was bound by the number of L1 <-> L2 transactions, not by their size).  The
//     partial granule is mirrored into the output ring every iteration, which is all a near back-reference needs;
//   * selects on the retire path are integer multiply-adds (pick): the ALU pipe is the busy one, the FMA pipe idles;
//   * a back-reference further than 96 bytes whose source is not yet in global memory waits for it (the chunk shrinks to what
//     is there, or the lane issues nothing for an iteration); short periods double (1, 2, 4, 8, 16 bytes per chunk);
//   * the last input granule is fetched with cp.async's src-size operand (the bytes beyond the block arrive as zeros), so the
//     end of the input needs no separate path and nothing beyond the unit is read.
//
// Everything unusual (literal with length bytes, LZ4 length runs, the LZ4 end-of-block zone) takes one slow branch that
// decodes from global memory with the oracle's rules; whatever fails a check, and every unit that is not 16-byte aligned,
// goes on the redo list of the warp-per-block kernel (generation 2, lz_decode.cuh), which owns all status codes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define G7_HD __host__ __device__ __forceinline__
#else
#define G7_HD inline
#endif

namespace cj {
namespace g7 {

constexpr int CODEC_SNAPPY = 0, CODEC_LZ4 = 2;   // = CJ_SNAPPY_RAW, CJ_LZ4_BLOCK (include/cramjam_cuda.h)
constexpr uint32_t IN_G = 8;             // input ring granules per lane
constexpr uint32_t INB = IN_G * 16;      // input ring bytes per lane
#ifndef CJ_G7_OUT_G
#define CJ_G7_OUT_G 8
#endif
constexpr uint32_t OUT_G = CJ_G7_OUT_G;   // granules of recent output per lane: the one being filled, OUT_G - 2 behind it, one ahead
constexpr uint32_t GROW = 512;           // bytes between consecutive granules of one lane (32 lanes x 16 bytes)
constexpr uint32_t NEAR = (OUT_G - 2) * 16;   // back-references up to this offset are read from the output ring at retire time
constexpr uint32_t MAXU = 1u << 30;
constexpr uint32_t warp_bytes(int D) { return (IN_G + OUT_G + 2u * (uint32_t)D) * GROW; }   // + a staging pair per chunk in flight
constexpr uint32_t cta_bytes(int D, int warps) { return warp_bytes(D) * (uint32_t)warps + 1024u; }   // + the 256-entry tag table

struct u4 {
    uint32_t x, y, z, w;
};

G7_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {   // ((hi:lo) >> (sh & 31)) & 0xffffffff
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
G7_HD uint32_t low_bits(uint32_t nbits) {   // mask of the low min(nbits, 32) bits
#if defined(__CUDA_ARCH__)
    return __funnelshift_lc(0xFFFFFFFFu, 0u, nbits);
#else
    return nbits >= 32u ? 0xFFFFFFFFu : ((1u << nbits) - 1u);
#endif
}
G7_HD uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
// m ? b : a for m in {0, 1}.  On the device two integer multiply-adds instead of a select: the kernel is bound by the ALU pipe
// (selects, logic, shifts, compares: one warp instruction per two cycles), while the FMA pipe that executes IMAD idles.
G7_HD uint32_t pick(uint32_t a, uint32_t b, uint32_t m) {
#if defined(__CUDA_ARCH__)
    uint32_t d, r;
    asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(d) : "r"(a), "r"(b));   // b - a
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(m), "r"(d), "r"(a));
    return r;
#else
    return m ? b : a;
#endif
}
G7_HD int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

// Snappy tag table: [6:0] compressed size of the element, [14:8] bytes it produces, [23:22] kind, [21:16] field,
// bit 31 = not a plain element (literal with length bytes, 4-byte-offset copy).
constexpr uint32_t K_LIT = 0, K_M16 = 1, K_M1 = 2;
G7_HD uint32_t tag_entry(uint32_t tag) {
    const uint32_t type = tag & 3, L = (tag >> 2) + 1;
    if (type == 0) return L <= 60 ? ((1 + L) | (L << 8) | (((K_LIT << 6) | (L - 1)) << 16)) : 0x80000000u;
    if (type == 1) {
        const uint32_t len = 4 + ((tag >> 2) & 7);
        return 2 | (len << 8) | (((K_M1 << 6) | (tag >> 5)) << 16);
    }
    if (type == 2) return 3 | (L << 8) | (((K_M16 << 6) | (L - 1)) << 16);
    return 0x80000000u;
}

// The program of one lane over one block.  `Env` supplies the memory operations (shared-memory accesses by 32-bit address,
// predicated stores, cp.async, the warp vote), the addresses in_l / out_l / st_l of granule 0 of the lane's input ring,
// output ring and staging pairs, and the tag table lut.  has == false: the lane has no block and only keeps step with its warp.
//
// One PASS of the loop is D iterations (slots u = 0..D-1, unrolled); an iteration retires the chunk issued into its slot one
// pass earlier and issues a new one.  Each iteration commits one cp.async group, so wait_group(D-1) at its top guarantees
// everything issued one pass ago.  The input ring is refilled once per pass, two granules at a time.
template <int CODEC, int D, class Env>
G7_HD void decode_block(Env& env, bool has, const uint8_t* src, uint8_t* dst, uint64_t sl, uint64_t dcap) {
    const uint32_t in_l = env.in_l, out_l = env.out_l, st_l = env.st_l, lut = env.lut;
    constexpr uint32_t RUN = 0, DRAIN = 1, IDLE = 2;   // decoding / input consumed, chunks still in flight / nothing to do
    uint32_t st = IDLE;
    uint32_t n = 0, ulen = 0, ip = 0;
    if (has) {
        bool ok = sl >= 1 && sl <= MAXU && (((uintptr_t)src | (uintptr_t)dst) & 15u) == 0;
        if (ok) {
            n = (uint32_t)sl;
            if (CODEC == CODEC_SNAPPY) {
                uint64_t v = 0;
                bool done = false;
                for (int i = 0; i < 5 && ip < n; i++) {
                    const uint32_t x = env.ldg8(src + ip++);
                    v |= (uint64_t)(x & 0x7f) << (7 * i);
                    if (!(x & 0x80)) { done = true; break; }
                }
                ok = done && v >= 1 && v <= dcap && v <= MAXU;
                ulen = (uint32_t)v;
            } else {
                ok = dcap >= 1 && dcap <= MAXU;
                ulen = (uint32_t)dcap;   // LZ4: the capacity; LZ4_decompress_safe's rules are applied against it
            }
        }
        if (ok) st = RUN;
        else env.redo();
    }
    const uint32_t nload = st == RUN ? ((n + 31u) & ~31u) : 0u;   // granule pairs are requested up to here; bytes beyond n arrive as zeros
    uint32_t loaded = 0, lim = 0, lnext = 0;   // input bytes requested / known to have arrived in the ring / requested one pass ago
    uint32_t opi = 0, opr = 0;       // output position of the next chunk to issue / to retire
    uint32_t rem = 0, sp = 0;        // bytes of the current element still to issue; literal: input position, copy: (effective) offset
    bool is_lit = false;
    bool lz_phase = false, lz_last = false;   // LZ4: the next thing to decode is an offset (not a token); the final sequence has been seen
    uint32_t lz_ml = 0;                       // LZ4: match-length nibble of the token whose literals are being issued
    uint32_t tw0 = 0, tw1 = 0;       // the two ring words around ip, fetched one iteration ahead
    bool tw_ok = false;              // ... and whether they had arrived when they were fetched
    u4 acc = {0, 0, 0, 0};           // the output granule being filled: bytes [opr & ~15, opr)
    u4 pend = acc;                   // a finished granule that starts a 32-byte sector waits here for its partner: the two leave with one 32-byte store
    bool have_pend = false;
    const uint32_t ph = (uint32_t)((uintptr_t)dst >> 4) & 1u;   // granule g of the output starts a 32-byte sector of memory iff g + ph is even
    // chunk in flight, per slot: M = bytes, DD = byte offset of its source window in the first of its two source granules,
    // GA / GB = shared-memory addresses of those granules (registers are not the scarce resource here, ALU instructions are)
    uint32_t M[D], DD[D], GA[D], GB[D], IPH[D];
#pragma unroll
    for (int u = 0; u < D; u++) { M[u] = 0; DD[u] = 0; GA[u] = out_l; GB[u] = out_l; IPH[u] = 0; }

    while (env.any(st != IDLE)) {
#pragma unroll
        for (int u = 0; u < D; u++) {
            env.tick();
            env.template wait<D - 1>();
            if (u == 0) lim = lnext;   // the pair requested one pass ago has arrived
            // ---- loads first: the two source granules of the chunk issued one pass ago, and the table entry of the tag fetched one
            //      iteration ago ----
            const uint32_t c = M[u], d = DD[u];
            const u4 sa = env.lds128(GA[u]), sb = env.lds128(GB[u]);
            const uint32_t t = funnel_r(tw0, tw1, ip * 8u);   // the funnel shift takes its amount mod 32
            uint32_t ent = 0;
            if (CODEC == CODEC_SNAPPY) ent = env.lds32(lut + 4u * (t & 255u));
            // ---- retire: pick the six words the source window starts in, align them to the output phase, place them behind the
            //      bytes of the granule being filled ----
            {
                const uint32_t w1 = (d >> 2) & 1u, w2 = (d >> 3) & 1u;
                const uint32_t u0 = pick(sa.x, sa.y, w1), u1 = pick(sa.y, sa.z, w1), u2 = pick(sa.z, sa.w, w1), u3 = pick(sa.w, sb.x, w1),
                               u4_ = pick(sb.x, sb.y, w1), u5 = pick(sb.y, sb.z, w1), u6 = pick(sb.z, sb.w, w1), u7 = sb.w;
                const uint32_t t0 = pick(u0, u2, w2), t1 = pick(u1, u3, w2), t2 = pick(u2, u4_, w2), t3 = pick(u3, u5, w2), t4 = pick(u4_, u6, w2),
                               t5 = pick(u5, u7, w2);
                const uint32_t sh = d * 8u;
                const uint32_t x0 = funnel_r(t0, t1, sh), x1 = funnel_r(t1, t2, sh), x2 = funnel_r(t2, t3, sh), x3 = funnel_r(t3, t4, sh),
                               x4 = funnel_r(t4, t5, sh);
                // x0..x4 hold the chunk from byte (opr & 3) of x0 on; word j of the 32-byte window takes x[j - wo], wo = word of opr
                const uint32_t o1 = (opr >> 2) & 1u, o2 = (opr >> 3) & 1u;
                const uint32_t z0 = x0, z1 = pick(x1, x0, o1), z2 = pick(x2, x1, o1), z3 = pick(x3, x2, o1), z4 = pick(x4, x3, o1), z5 = x4;
                const uint32_t y0 = z0, y1 = z1, y2 = pick(z2, z0, o2), y3 = pick(z3, z1, o2), y4 = pick(z4, z2, o2), y5 = pick(z5, z3, o2), y6 = z4, y7 = z5;
                // bytes below opr come from the accumulator
                const int32_t ob = (int32_t)((opr & 15u) * 8u);
                const uint32_t m0 = low_bits((uint32_t)ob), m1 = low_bits((uint32_t)imax(ob - 32, 0)), m2 = low_bits((uint32_t)imax(ob - 64, 0)),
                               m3 = low_bits((uint32_t)imax(ob - 96, 0));
                u4 lo;
                lo.x = (acc.x & m0) | (y0 & ~m0);
                lo.y = (acc.y & m1) | (y1 & ~m1);
                lo.z = (acc.z & m2) | (y2 & ~m2);
                lo.w = (acc.w & m3) | (y3 & ~m3);
                const bool cross = (opr & 15u) + c >= 16u;   // the granule is complete: store it, go on with the upper half
                const uint32_t ga = out_l + ((opr >> 4) & (OUT_G - 1)) * GROW, gb = out_l + (((opr >> 4) + 1u) & (OUT_G - 1)) * GROW;
                const bool odd = (((opr >> 4) + ph) & 1u) != 0;   // the finished granule is the second half of its sector
                env.stg256_if(dst + (opr & ~15u) - 16, pend, lo, cross && odd && have_pend);
                env.stg128_if(dst + (opr & ~15u), lo, cross && odd && !have_pend);   // no partner: only the first granule of a block that starts mid-sector
                const uint32_t cr = cross ? 1u : 0u;
                pend.x = pick(pend.x, lo.x, cr);
                pend.y = pick(pend.y, lo.y, cr);
                pend.z = pick(pend.z, lo.z, cr);
                pend.w = pick(pend.w, lo.w, cr);
                have_pend = cross ? !odd : have_pend;
                acc.x = pick(lo.x, y4, cross ? 1u : 0u);
                acc.y = pick(lo.y, y5, cross ? 1u : 0u);
                acc.z = pick(lo.z, y6, cross ? 1u : 0u);
                acc.w = pick(lo.w, y7, cross ? 1u : 0u);
                env.sts128(ga, lo);   // the output ring always holds the granule being filled as far as it is known
                env.sts128_if(gb, acc, cross);
                opr += c;
            }
            // ---- decode the next element if the current one is fully issued ----
            const bool need = st == RUN && rem == 0;
            const bool atend = ip >= n;
            bool slow;
            if (CODEC == CODEC_SNAPPY) {
                const bool fast = need && !atend && tw_ok;
                const uint32_t adv = ent & 0x7Fu, len = (ent >> 8) & 0x7Fu, kind = (ent >> 22) & 3u;
                const uint32_t off = kind == K_M1 ? (((ent >> 16) & 7u) << 8) | ((t >> 8) & 0xFFu) : (t >> 8) & 0xFFFFu;
                const bool lit = kind == K_LIT;
                const bool bad = (int32_t)ent < 0 || ip + adv > n || opi + len > ulen || (!lit && off - 1u >= opi);
                const bool take = fast && !bad;
                slow = fast && bad;   // rare tag or failed check
                st = (need && atend) ? DRAIN : st;
                is_lit = take ? lit : is_lit;
                sp = take ? (lit ? ip + 1 : off) : sp;
                rem = take ? len : rem;
                ip = take ? ip + adv : ip;
            } else {
                // LZ4: a sequence is decoded in two steps, its token (-> the literal run) and, once the literals are issued, its
                // offset (-> the match); a token without literals does both at once.  One length-extension byte is taken here,
                // longer runs and the end-of-block zone (MFLIMIT / LASTLITERALS rules of LZ4_decompress_safe) go to the slow path.
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
                // 32-bit sums cannot wrap: positions are at most MAXU = 2^30, the lengths of this path at most 15 + 255 + 4
                const bool tailz = !lz_phase && (opi + ll_tot + 12u > ulen || q + ll_tot + 8u > n);
                const bool mbad = !lit0 && (offv - 1u >= opi || opi + mtot + 5u > ulen || ip + madv > n);
                const bool bad = longrun || tailz || mbad;
                const bool take = fast && !bad;
                slow = (fast && bad) || (need && atend && !lz_last);
                st = (need && atend && lz_last) ? DRAIN : st;
                is_lit = take ? lit0 : is_lit;
                sp = take ? (lit0 ? q : offv) : sp;
                rem = take ? (lit0 ? ll_tot : mtot) : rem;
                ip = take ? (lit0 ? q + ll_tot : ip + madv) : ip;
                lz_ml = (take && lit0) ? (b0 & 15u) : lz_ml;
                lz_phase = take ? lit0 : lz_phase;
            }
            // ---- issue one chunk of the current element: as many of its next 16 bytes as their source allows ----
            {
                const uint32_t c16 = umin(rem, 16u);
                const uint32_t k = opi & 3u;
                const uint32_t fs = opi - sp;             // copy: output position of its source
                const uint32_t F = (opr & ~15u) - (have_pend ? 16u : 0u);   // output below F is in global memory
                const bool isnear = sp <= NEAR;
                IPH[u] = (is_lit && rem != 0) ? sp : ip;
                // the source window starts k bytes before the source (those bytes are replaced by the accumulator's): its position
                // in the input (literal) or in the output (copy), and the offset of its first byte in the first of its two granules
                const uint32_t sk = (is_lit ? sp : fs) - k;
                const uint32_t dd = sk & 15u;
                // bytes the source can supply now: literal bytes that have arrived / a near copy never overlaps its own source (a
                // short period doubles below) / the part of a far source that is in global memory already; and a chunk never
                // reaches beyond its two granules
                const int32_t avail = (int32_t)(is_lit ? lim - sp : (isnear ? sp : F - fs));
                const uint32_t cn = avail <= 0 ? 0u : umin(umin(c16, (uint32_t)avail), 32u - dd - k);
                const bool isfar = !is_lit && !isnear && cn != 0;
                const int32_t g0 = (int32_t)(sk & ~15u);   // far: output position of the first source granule (-16: in front of the block)
                const uint32_t s0 = st_l + 2u * (uint32_t)u * GROW;   // ... and this slot's staging pair
                env.cp16_far_if(s0, dst + (g0 < 0 ? 0 : g0), isfar && g0 >= 0 && dd + k < 16u);   // not if it would only hold the k bytes in front of the source
                env.cp16_far_if(s0 + GROW, dst + (g0 + 16), isfar && dd + k + cn > 16u);   // second granule only if the chunk reaches into it
                const uint32_t rbase = is_lit ? in_l : out_l;
                const uint32_t rmask = is_lit ? (IN_G - 1) : (OUT_G - 1);
                const uint32_t gi = sk >> 4;
                const bool ring = is_lit || isnear;
                const uint32_t a0 = ring ? rbase + (gi & rmask) * GROW : s0, a1 = ring ? rbase + ((gi + 1u) & rmask) * GROW : s0 + GROW;
                GA[u] = a0;
                GB[u] = a1;
                M[u] = cn;
                DD[u] = dd;
                sp = is_lit ? sp + cn : ((cn == sp && sp < 16u) ? sp + sp : sp);
                opi += cn;
                rem -= cn;
            }
            // ---- everything unusual, at most a few times per block ----
            if (slow) {
                bool fail = false;
                if (CODEC == CODEC_SNAPPY) {   // a literal with length bytes, decoded from global memory with every check
                    const uint32_t tag = env.ldg8(src + ip);
                    const uint32_t nb = (tag >> 2) - 59;
                    if ((int32_t)tag_entry(tag) >= 0 || (tag & 3u) != 0 || nb > n - ip - 1) fail = true;   // failed element; 4-byte-offset copies: generation 2
                    else {
                        uint32_t v = 0;
                        for (uint32_t i = 0; i < nb; i++) v |= env.ldg8(src + ip + 1 + i) << (8 * i);
                        const uint64_t LL = (uint64_t)v + 1;
                        const uint32_t q = ip + 1 + nb;
                        if (LL > n - q || LL > ulen - opi) fail = true;
                        else { is_lit = true; sp = q; rem = (uint32_t)LL; ip = q + (uint32_t)LL; }
                    }
                } else {   // LZ4, step by step as LZ4_decompress_safe does it; whatever it rejects is generation 2's
                    if (ip >= n) fail = true;
                    else if (!lz_phase) {
                        const uint32_t token = env.ldg8(src + ip);
                        uint32_t p = ip + 1;
                        uint64_t len = token >> 4;
                        if (len == 15) {
                            if (n < 15 || p >= n - 15) fail = true;
                            else {
                                uint32_t bb;
                                do {
                                    bb = env.ldg8(src + p++);
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
                        if (ip + 2 > n) fail = true;
                        else {
                            const uint32_t off = env.ldg8(src + ip) | (env.ldg8(src + ip + 1) << 8);
                            uint32_t p = ip + 2;
                            uint64_t len = lz_ml;
                            if (len == 15) {
                                uint32_t bb;
                                do {
                                    if (p >= n) { fail = true; break; }
                                    bb = env.ldg8(src + p++);
                                    len += bb;
                                    if (p > n - 4) { fail = true; break; }
                                } while (bb == 255);
                            }
                            len += 4;
                            if (fail || off == 0 || off > opi || (uint64_t)opi + len + 5 > ulen) fail = true;
                            else { is_lit = false; sp = off; rem = (uint32_t)len; ip = p; lz_phase = false; }
                        }
                    }
                }
                if (fail) {
                    env.redo();
                    st = IDLE;   // chunks in flight still retire (inside the capacity); the redo pass rewrites the unit
                    rem = 0;
                }
            }
            // ---- fetch the tag words of the next element; their latency overlaps the loop bookkeeping ----
            {
                const uint32_t w0 = ip >> 2, w1 = w0 + 1u;
                tw0 = env.lds32(in_l + ((w0 >> 2) & (IN_G - 1)) * GROW + (w0 & 3u) * 4u);
                tw1 = env.lds32(in_l + ((w1 >> 2) & (IN_G - 1)) * GROW + (w1 & 3u) * 4u);
                tw_ok = umin(ip + 4u, nload) <= lim;
            }
            // ---- input ring, once per pass: one more pair of granules if it fits ahead of everything still needed ----
            if (u == 0) {
                const uint32_t low = IPH[(u + 1) % D];   // the oldest slot in flight: no chunk in flight reads input below this (less 3 bytes)
                const uint32_t keep = (low < 3u ? 0u : low - 3u) & ~15u;
                const bool go = st == RUN && loaded < nload && loaded + 32u <= keep + INB;
                const uint32_t ra = in_l + ((loaded >> 4) & (IN_G - 1)) * GROW;   // loaded is a multiple of 32: the pair does not wrap
                const uint32_t left = n - loaded;   // >= 1 when go
                const uint32_t z1 = umin(left, 16u), z2 = left > 16u ? umin(left - 16u, 16u) : 0u;   // bytes beyond the block are zero-filled, not read
                const uint8_t* g1 = src + loaded;
                const uint8_t* g2 = g1 + (left > 16u ? 16u : 0u);
                env.cp16_in_if(ra, g1, z1, go);
                env.cp16_in_if(ra + GROW, g2, z2, go);
                loaded += go ? 32u : 0u;
                lnext = loaded;   // this pair joins the group committed below: it has arrived when the next pass begins
            }
            env.commit();
        }
        // ---- a lane is done when its input is consumed and every chunk has retired ----
        if (st == DRAIN) {
            bool empty = true;
#pragma unroll
            for (int u = 0; u < D; u++) empty = empty && M[u] == 0;
            if (empty) {
                if (CODEC == CODEC_SNAPPY && opi != ulen) env.redo();
                else {
                    const uint32_t k = opr & 15u;   // the bytes of the unfinished granule
                    const uint32_t ga = out_l + ((opr >> 4) & (OUT_G - 1)) * GROW;
                    env.stg128_if(dst + (opr & ~15u) - 16, pend, have_pend);   // a finished granule still waiting for its partner
                    for (uint32_t j = 0; j < k; j++) env.stg8(dst + (opr & ~15u) + j, env.lds8(ga + j));
                    env.finish_ok(opi);
                }
                st = IDLE;
            }
        }
    }
    env.template wait<0>();
}

}  // namespace g7
}  // namespace cj
