// lz_match.cuh — warp-parallel greedy hash match finder shared by the LZ4 / Snappy block encoders and
// the zstd encoder.  32 consecutive positions are hashed per step (one per lane), looked up in a
// 4096-entry table of 32-bit positions in shared memory, then inserted — when several lanes share a
// bucket the highest position wins, chosen with __match_any_sync so the result never depends on store
// ordering (the encoders are deterministic).  Every lane verifies its candidate with one 4-byte
// compare; the matches of a step are taken greedily in position order and extended 32 bytes per ballot.
#pragma once
#include "common.cuh"

namespace cj {

#ifndef CJ_ENC_HBITS
#define CJ_ENC_HBITS 12
#endif
constexpr int ENC_HBITS = CJ_ENC_HBITS;
constexpr int ENC_HSIZE = 1 << ENC_HBITS;
constexpr uint32_t ENC_MAXOFF = 65535;
// A table slot keeps the low 16 bits of a position.  The candidate is rebuilt as pos - ((pos - slot) & 0xFFFF):
// the most recent position congruent to the slot, at most 65535 bytes back (the LZ4 / Snappy-2 / window limit
// all three encoders share).  There is no "empty" value: a zeroed table proposes the 64 KiB-aligned base, and
// every candidate is verified by a 4-byte compare, so a stale or never-written slot only costs a failed compare.
// Halving the slot size doubles the warps an SM can hold (8 KiB instead of 16 KiB of shared memory per warp).
using enc_slot_t = uint16_t;
constexpr size_t ENC_TABLE_BYTES = (size_t)ENC_HSIZE * sizeof(enc_slot_t);

__device__ __forceinline__ uint32_t load32u(const uint8_t* p) {  // unaligned little-endian 32-bit load
    const uint32_t a = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p - a);
    const uint32_t lo = __ldg(w);
    if (a == 0) return lo;
    const uint32_t hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, a * 8);
}

template <int HBITS = ENC_HBITS>
__device__ __forceinline__ void match_table_reset(enc_slot_t* table, int lane) {
    for (uint32_t i = lane; i < ((size_t)sizeof(enc_slot_t) << HBITS) / 16; i += 32) reinterpret_cast<uint4*>(table)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
}

// One 32-position window in flight between the two halves of a step (registers only).
struct MatchWindow {
    uint32_t v, v4, v8;        // this lane's position-side words at +0, +4, +8
    uint32_t w0, w1, w2, w3;   // candidate-side words (aligned loads), shifted by `sh` when compared
    uint32_t sh;               // candidate misalignment in bits
    uint32_t cand, d;          // candidate position and distance
    bool probe, deep;          // a candidate exists; all 12 bytes on both sides are inside the input
    // second way of the bucket (WAYS == 2, the HC-class search): the position the bucket held before the most recent one
    uint32_t x0, x1, x2, x3, xsh, cand2, d2;
    bool probe2;
};

// First half of a step: hash the 32 positions of the window at p, look them up, insert them (highest position of a bucket
// wins, whatever the store order), and START the loads of the candidate words.  Nothing here waits for global memory.
// WAYS = 2: a bucket keeps its two most recent positions (a hash chain of depth two); both candidates are verified and the
// longer match wins.  The bucket count is 2^HBITS either way, so the table is WAYS x 2^HBITS slots.
template <int HBITS, int WAYS = 1>
__device__ __forceinline__ void match_probe(const uint8_t* __restrict__ src, uint32_t p, uint32_t start_limit, uint32_t v, uint32_t vn,
                                            enc_slot_t* table, int lane, MatchWindow& w) {
    const uint32_t pos = p + lane;
    const bool valid = pos < start_limit;
    const uint32_t h = (v * 0x9E3779B1u) >> (32 - HBITS);
    uint32_t slot = 0, slot2 = 0;
    if (valid) {
        slot = table[WAYS * h];
        if (WAYS == 2) slot2 = table[2 * h + 1];
    }
    __syncwarp();
    const uint32_t grp = __match_any_sync(FULL, valid ? h : (0x80000000u | lane));
    if (valid && lane == 31 - __clz(grp)) {   // highest position of the bucket wins; with two ways the old head moves down
        table[WAYS * h] = (enc_slot_t)pos;
        if (WAYS == 2) table[2 * h + 1] = (enc_slot_t)slot;
    }
    __syncwarp();
    w.d = (pos - slot) & 0xFFFFu;
    w.cand = pos - w.d;
    w.probe = valid && w.d != 0 && w.d <= pos;
    // position-side words at +4 and +8 (real only while pos + 8 < start_limit)
    const uint32_t v4a = __shfl_sync(FULL, v, (lane + 4) & 31), v4b = __shfl_sync(FULL, vn, (lane + 4) & 31);
    const uint32_t v8a = __shfl_sync(FULL, v, (lane + 8) & 31), v8b = __shfl_sync(FULL, vn, (lane + 8) & 31);
    w.v = v;
    w.v4 = lane + 4 < 32 ? v4a : v4b;
    w.v8 = lane + 8 < 32 ? v8a : v8b;
    w.deep = pos + 8 < start_limit;
    w.w0 = w.w1 = w.w2 = w.w3 = 0;
    w.sh = 0;
    if (w.probe) {
        const uint32_t a = (uint32_t)((uintptr_t)(src + w.cand) & 3u);
        const uint32_t* q = reinterpret_cast<const uint32_t*>(src + w.cand - a);
        w.sh = a * 8;
        w.w0 = __ldg(q);
        w.w1 = __ldg(q + 1);   // holds cand + 3 or lies inside [cand, pos): always in bounds
        if (w.deep) { w.w2 = __ldg(q + 2); w.w3 = __ldg(q + 3); }
    }
    w.probe2 = false;
    if (WAYS == 2) {
        w.d2 = (pos - slot2) & 0xFFFFu;
        w.cand2 = pos - w.d2;
        w.probe2 = valid && w.d2 != 0 && w.d2 <= pos && w.d2 != w.d;
        w.x0 = w.x1 = w.x2 = w.x3 = 0;
        w.xsh = 0;
        if (w.probe2) {
            const uint32_t a = (uint32_t)((uintptr_t)(src + w.cand2) & 3u);
            const uint32_t* q = reinterpret_cast<const uint32_t*>(src + w.cand2 - a);
            w.xsh = a * 8;
            w.x0 = __ldg(q);
            w.x1 = __ldg(q + 1);
            if (w.deep) { w.x2 = __ldg(q + 2); w.x3 = __ldg(q + 3); }
        }
    }
}

// Second half: this lane's match length at its position: 0 none, 4..11 final, 12 = at least 12 (or not extended, see deep).
__device__ __forceinline__ uint32_t match_len12(bool probe, bool deep, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t sh, uint32_t v,
                                                uint32_t v4, uint32_t v8) {
    if (!probe) return 0;
    const uint32_t c0 = __funnelshift_r(a0, a1, sh);
    if (c0 != v) return 0;
    if (!deep) return 4;
    const uint32_t x1 = __funnelshift_r(a1, a2, sh) ^ v4, x2 = __funnelshift_r(a2, a3, sh) ^ v8;
    return x1 ? 4 + ((__ffs(x1) - 1) >> 3) : (x2 ? 8 + ((__ffs(x2) - 1) >> 3) : 12);
}
// With two ways the longer of the two candidates' matches is kept (the nearer one on a tie): cand / d are switched to it.
template <int WAYS>
__device__ __forceinline__ uint32_t match_verify(MatchWindow& w) {
    const uint32_t m1 = match_len12(w.probe, w.deep, w.w0, w.w1, w.w2, w.w3, w.sh, w.v, w.v4, w.v8);
    if (WAYS == 1) return m1;
    const uint32_t m2 = match_len12(w.probe2, w.deep, w.x0, w.x1, w.x2, w.x3, w.xsh, w.v, w.v4, w.v8);
    if (m2 > m1) { w.cand = w.cand2; w.d = w.d2; return m2; }
    return m1;
}

// Scans src[begin, end): positions in [begin, start_limit) may start a match, a match may not pass
// match_limit (start_limit <= match_limit - 3).  Returns the position where the trailing literal run starts.
//
// The scan advances one 32-position window per step and is software-pipelined over two windows: the table lookups and
// candidate loads of window p + 32 (match_probe) are issued BEFORE the matches of window p are verified, chosen and emitted,
// so the L2 / DRAM round trip of the candidate words — what a warp of the unpipelined scan sat waiting for most of its
// time (ncu: 4.9 long-scoreboard stalls per issue, throughput linear in resident warps) — overlaps a whole step of work.
// Every lane that finds a verified candidate has its match up to 12 bytes from registers (the position-side words come
// from the neighbouring lanes by shuffle); 84 % of the bench corpus' matches need nothing more, longer ones are extended by
// the whole warp, 32 bytes per ballot.  A match that reaches into later windows masks the lanes it covers there (they are
// still inserted into the table); a match that covers whole windows makes the scan jump ahead.  The matches of a step are
// chosen greedily in position order (registers only), then handed to the emitter together:
//   em.window(src, p, v, anchor, sel, mlen, off)  all matches of the step at once, lane i of `sel` holding match
//                                              (p + i, mlen, off); v = the lane's 4-byte word (its low byte is
//                                              src[p + lane]); anchor = start of the pending literal run.
//                                              Returns false if it wants the step one match at a time instead:
//   em.serial(anchor, literal_len, offset, match_len)  warp-uniform.
// HBITS = log2 of the table size: the speed / ratio knob of the block encoders (lz4 `acceleration`, HC-class levels).
// WAYS = 2 adds the second candidate per position and a one-step lazy choice (a match is passed over when the next position
// starts a longer one): the HC-class search of the block encoders (lz4 `compression=Some(n)`, lz4 frame level >= 3).
template <class Emitter, int HBITS = ENC_HBITS, int WAYS = 1>
__device__ __forceinline__ uint32_t find_matches(const uint8_t* __restrict__ src, uint32_t begin, uint32_t start_limit, uint32_t match_limit,
                                                 enc_slot_t* table, int lane, Emitter& em) {
    uint32_t anchor = begin;
    uint32_t p = begin;
    if (p >= start_limit) return anchor;
    auto words = [&](uint32_t at) { return at + lane < start_limit ? load32u(src + at + lane) : 0u; };
    uint32_t v1 = words(p + 32);
    MatchWindow cur, nxt;
    match_probe<HBITS, WAYS>(src, p, start_limit, words(p), v1, table, lane, cur);
    for (;;) {
        // ---- first half of the NEXT window: its loads fly while this window is finished ----
        const uint32_t p1 = p + 32;
        const bool more = p1 < start_limit;
        uint32_t v2 = 0;
        if (more) {
            v2 = words(p1 + 32);
            match_probe<HBITS, WAYS>(src, p1, start_limit, v1, v2, table, lane, nxt);
        }
        // ---- second half of THIS window ----
        uint32_t mlen = match_verify<WAYS>(cur);
        const bool deep = cur.deep;
        uint32_t mm = __ballot_sync(FULL, mlen != 0);
        if (anchor > p) mm &= anchor - p >= 32 ? 0u : ~((1u << (anchor - p)) - 1);  // lanes covered by the previous match
        if (mm) {
            // greedy choice in position order; only matches that may be longer than 12 touch memory
            const uint32_t open_ended = __ballot_sync(FULL, mlen == 12 || (mlen != 0 && !deep));
            uint32_t sel = 0, last_end = 0;
            while (mm) {
                const int i = __ffs(mm) - 1;
                const uint32_t mpos = p + i;
                uint32_t len = __shfl_sync(FULL, mlen, i);
                if (WAYS == 2 && i < 31 && ((mm >> (i + 1)) & 1)) {   // lazy: the next position starts a longer match -> one more literal
                    const uint32_t len1 = __shfl_sync(FULL, mlen, i + 1);
                    if (len < 12 && len1 > len) { mm &= mm - 1; continue; }
                }
                if ((open_ended >> i) & 1) {  // extend the match, 32 bytes per ballot
                    const uint32_t c = __shfl_sync(FULL, cur.cand, i);
                    const uint32_t maxlen = match_limit - mpos;
                    for (;;) {
                        const uint32_t k = len + lane;
                        const bool eq = k < maxlen && __ldg(src + c + k) == __ldg(src + mpos + k);
                        const uint32_t ne = __ballot_sync(FULL, !eq);
                        if (ne) { len += __ffs(ne) - 1; break; }
                        len += 32;
                    }
                    if (lane == i) mlen = len;
                }
                sel |= 1u << i;
                last_end = mpos + len;
                const uint32_t covered = last_end - p;  // lanes below this were swallowed by the match
                mm = covered >= 32 ? 0u : mm & ~((1u << covered) - 1);
            }
            if (!em.window(src, p, cur.v, anchor, sel, mlen, cur.d)) {
                uint32_t s = sel, a = anchor;
                while (s) {
                    const int i = __ffs(s) - 1;
                    s &= s - 1;
                    const uint32_t len = __shfl_sync(FULL, mlen, i), off = __shfl_sync(FULL, cur.d, i);
                    em.serial(a, p + i - a, off, len);
                    a = p + i + len;
                }
            }
            anchor = last_end;
        }
        if (!more) break;
        if (anchor >= p1 + 32) {
            // the last match covers the next window entirely (long runs): jump to the window that holds its end; the
            // pipeline restarts there (the windows in between are not inserted into the table)
            p = p1 + ((anchor - p1) & ~31u);
            if (p >= start_limit) break;
            v1 = words(p + 32);
            match_probe<HBITS, WAYS>(src, p, start_limit, words(p), v1, table, lane, cur);
            continue;
        }
        cur = nxt;
        v1 = v2;
        p = p1;
    }
    return anchor;
}

}  // namespace cj
