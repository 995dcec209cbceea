// lz_decode7.cuh — generation-7 batch decode of Snappy raw blocks and LZ4 blocks: the LANE PROGRAM, shared by the device
// kernel (lz_decode7.cu) and the host-side emulation that the CPU tests run (tests/emu/g7_emu.cpp).
//
// Reference entry points: snap::raw::Decoder::decompress behind cramjam.snappy.decompress_raw / decompress_raw_into
// (src/snappy.rs:52-60,102-108) and LZ4_decompress_safe behind lz4::block::decompress_into (src/lz4.rs:78-95,140-173).
//
// One THREAD per block, like generation 4 (lz_decode4.cu), rebuilt around its two measured costs (229 warp
// instructions per <= 8-byte chunk, integer pipe 76 % busy):
//
//   * CHUNKS are up to 16 bytes (9 900 instead of 12 600 sub-iterations per 64 KiB block of the bench corpus);
//   * every per-lane shared-memory structure is LINEAR in the lane's own record, so a chunk's source is one byte address
//     computed when the chunk is issued and its six source words are read at immediate offsets from it:
//       - the 128-byte input ring carries a 32-byte copy of its head behind its end (written by the same cp.async group), so a
//         24-byte read never wraps;
//       - a far back-reference is fetched from the block's own output in global memory by one or two 16-byte cp.async into a
//         48-byte staging slot (16 bytes of slack in front of the data);
//       - the output is assembled in a 64-byte window [P1][P0][A][B]: A is the 16-byte granule being filled, B takes what
//         spills over, P1/P0 are the two granules before A.  When A is complete it is stored with one st.global.v4 and the
//         window moves down by one granule.  Back-references of at most 32 bytes are read straight from the window when the
//         chunk retires (no stall, and short periods double: 1, 2, 4, 8, 16 bytes per chunk);
//   * a back-reference further than 32 bytes whose source is not yet in global memory waits for it (the chunk shrinks to what
//     is there, or the lane issues nothing for an iteration) instead of carrying a third source path through every chunk;
//   * retire does no source-kind dispatch at all: load six words, five funnel shifts, merge the first word with the bytes
//     already in the window, five word stores, the granule hand-over;
//   * the last input granule is fetched with cp.async's src-size operand (the bytes beyond the block arrive as zeros), so the
//     end of the input needs no separate path.
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
constexpr uint32_t INB = 128;            // input ring bytes per lane
constexpr uint32_t IN_SLOT = 16 + INB + 32;   // 16 bytes of slack in front, the copy of the ring's first 32 bytes behind
constexpr uint32_t ST_SLOT = 48;         // staging slot of one far chunk: 16 bytes of slack + two 16-byte granules
constexpr uint32_t ASM_SLOT = 80;        // 16 bytes of slack + [P1][P0][A][B]
constexpr uint32_t ASM_P1 = 16, ASM_P0 = 32, ASM_A = 48, ASM_B = 64;
constexpr uint32_t NEAR = 32;            // back-references up to this offset are read from the assembly window at retire time
constexpr uint32_t MAXU = 1u << 30;
constexpr uint32_t lane_payload(int D) { return IN_SLOT + (uint32_t)D * ST_SLOT + ASM_SLOT; }
// lane records are 16 bytes (mod 128) apart: same-offset 16-byte accesses of a quarter warp fall into different banks
constexpr uint32_t lane_stride(int D) { return lane_payload(D) + (16u + 128u - lane_payload(D) % 128u) % 128u; }
constexpr uint32_t warp_bytes(int D) { return 32u * lane_stride(D); }
constexpr uint32_t cta_bytes(int D, int warps) { return warp_bytes(D) * (uint32_t)warps + 1024u + 64u; }   // + the 256-entry tag table + slack behind the last record

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
G7_HD uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

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
// predicated cp.async / global stores, the warp vote) and the per-lane record addresses in_l / st_l / asm_l and the tag table
// lut.  has == false: the lane has no block and only keeps step with its warp.
//
// One PASS of the loop is D iterations (slots u = 0..D-1, unrolled); an iteration retires the chunk issued into its slot one
// pass earlier and issues a new one.  Each iteration commits one cp.async group, so wait_group(D-1) at its top guarantees
// everything issued one pass ago.  The input ring is refilled once per pass, two granules at a time.
template <int CODEC, int D, class Env>
G7_HD void decode_block(Env& env, bool has, const uint8_t* src, uint8_t* dst, uint64_t sl, uint64_t dcap) {
    const uint32_t in_l = env.in_l, st_l = env.st_l, asm_l = env.asm_l, lut = env.lut;
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
    uint32_t loaded = 0, lim = 0, lnext = 0;   // input bytes requested / known to have arrived / requested one pass ago
    uint32_t opi = 0, opr = 0;       // output position of the next chunk to issue / to retire
    uint32_t rem = 0, sp = 0;        // bytes of the current element still to issue; literal: input position, copy: (effective) offset
    bool is_lit = false;
    bool lz_phase = false, lz_last = false;   // LZ4: the next thing to decode is an offset (not a token); the final sequence has been seen
    uint32_t lz_ml = 0;                       // LZ4: match-length nibble of the token whose literals are being issued
    uint32_t tw0 = 0, tw1 = 0;       // the two ring words around ip, fetched one iteration ahead
    bool tw_ok = false;              // ... and whether they had arrived when they were fetched
    uint32_t M[D], P[D];             // chunk in flight: bytes, shared-memory byte address of its source
    uint32_t IPH[D];                 // lowest input position the decoder needed when the slot was issued (its literal chunk reads no lower)
#pragma unroll
    for (int u = 0; u < D; u++) { M[u] = 0; P[u] = asm_l + ASM_A; IPH[u] = 0; }

    while (env.any(st != IDLE)) {
#pragma unroll
        for (int u = 0; u < D; u++) {
            env.template wait<D - 1>();
            if (u == 0) lim = lnext;   // the pairs requested one pass ago have arrived
            const uint32_t st_u = st_l + (uint32_t)u * ST_SLOT;
            // ---- loads first (shared-memory accesses keep their program order): the six source words of the chunk issued one
            //      pass ago, the window word it starts in, and the table entry of the tag fetched one iteration ago ----
            const uint32_t c = M[u], a = P[u];
            const uint32_t a0 = a & ~3u;
            const uint32_t wa = asm_l + ASM_A + (opr & 12u);
            const uint32_t s0 = env.lds32(a0), s1 = env.lds32(a0 + 4), s2 = env.lds32(a0 + 8), s3 = env.lds32(a0 + 12), s4 = env.lds32(a0 + 16),
                           s5 = env.lds32(a0 + 20);
            const uint32_t w = env.lds32(wa);
            const uint32_t t = funnel_r(tw0, tw1, ip * 8u);   // the funnel shift takes its amount mod 32
            uint32_t ent = 0;
            if (CODEC == CODEC_SNAPPY) ent = env.lds32(lut + 4u * (t & 255u));
            // ---- retire: align to the output phase, merge with the bytes already in the window, store five words ----
            {
                const uint32_t sh = a * 8u;
                uint32_t x0 = funnel_r(s0, s1, sh);
                const uint32_t x1 = funnel_r(s1, s2, sh), x2 = funnel_r(s2, s3, sh), x3 = funnel_r(s3, s4, sh), x4 = funnel_r(s4, s5, sh);
                const uint32_t hm = 0xFFFFFFFFu << ((opr * 8u) & 31u);   // bytes of the first word that are not output yet
                x0 = (x0 & hm) | (w & ~hm);
                env.sts32(wa, x0);
                env.sts32(wa + 4, x1);
                env.sts32(wa + 8, x2);
                env.sts32(wa + 12, x3);
                env.sts32(wa + 16, x4);   // bytes beyond opr + c are not output yet: whatever lands there is overwritten later
                const bool cross = (opr & 15u) + c >= 16u;   // granule A is complete: store it, move the window down
                const u4 vp = env.lds128(asm_l + ASM_P0), va = env.lds128(asm_l + ASM_A), vb = env.lds128(asm_l + ASM_B);
                env.stg128_if(dst + (opr & ~15u), va, cross);
                env.sts128_if(asm_l + ASM_P1, vp, cross);
                env.sts128_if(asm_l + ASM_P0, va, cross);
                env.sts128_if(asm_l + ASM_A, vb, cross);
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
                const bool tailz = !lz_phase && ((uint64_t)opi + ll_tot + 12 > ulen || (uint64_t)q + ll_tot + 8 > n);
                const bool mbad = !lit0 && (offv - 1u >= opi || (uint64_t)opi + mtot + 5 > ulen || ip + madv > n);
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
                const uint32_t F = opr & ~15u;            // output below F is in global memory
                const bool isnear = sp <= NEAR;
                IPH[u] = (is_lit && rem != 0) ? sp : ip;
                // bytes the source can supply now: literal bytes that have arrived / a near copy never overlaps its own source (a
                // short period doubles below) / the part of a far source that is in global memory already
                const int32_t avail = (int32_t)(is_lit ? lim - sp : (isnear ? sp : F - fs));
                const uint32_t cn = avail <= 0 ? 0u : umin(c16, (uint32_t)avail);
                const bool isfar = !is_lit && !isnear && cn != 0;
                const uint8_t* gp = dst + (fs & ~15u);
                env.cp16_far_if(st_u + 16, gp, isfar);
                env.cp16_far_if(st_u + 32, gp + 16, isfar && (fs & 15u) + cn > 16u);   // second granule only if the chunk reaches into it
                const uint32_t base = is_lit ? in_l : (isnear ? asm_l + ASM_A : st_u + 16);
                const uint32_t boff = is_lit ? (sp & (INB - 1)) : (isnear ? (opi & 15u) - sp : (fs & 15u));
                P[u] = base + boff - k;
                M[u] = cn;
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
            tw0 = env.lds32(in_l + (ip & (INB - 4)));
            tw1 = env.lds32(in_l + (ip & (INB - 4)) + 4);
            tw_ok = umin(ip + 4u, nload) <= lim;
            // ---- input ring, once per pass: one more pair of granules if it fits ahead of everything still needed ----
            if (u == 0) {
                const uint32_t low = IPH[(u + 1) % D];   // the oldest slot in flight: no chunk in flight reads input below this (less 3 bytes)
                const uint32_t keep = (low < 3u ? 0u : low - 3u) & ~15u;
                const bool go = st == RUN && loaded < nload && loaded + 32u <= keep + INB;
                const uint32_t rpos = loaded & (INB - 1);
                const uint32_t left = n - loaded;   // >= 1 when go
                const uint32_t z1 = umin(left, 16u), z2 = left > 16u ? umin(left - 16u, 16u) : 0u;   // bytes beyond the block are zero-filled, not read
                const uint8_t* g1 = src + loaded;
                const uint8_t* g2 = g1 + (left > 16u ? 16u : 0u);
                env.cp16_in_if(in_l + rpos, g1, z1, go);
                env.cp16_in_if(in_l + rpos + 16, g2, z2, go);
                env.cp16_in_if(in_l + INB, g1, z1, go && rpos == 0);        // the copy of the ring's head behind its end
                env.cp16_in_if(in_l + INB + 16, g2, z2, go && rpos == 0);
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
                    const uint32_t k = opr & 15u;   // the bytes of the unfinished granule A
                    for (uint32_t j = 0; j < k; j++) env.stg8(dst + (opr & ~15u) + j, env.lds8(asm_l + ASM_A + j));
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
