// lz_match.cuh — warp-parallel greedy hash match finder shared by the LZ4 / Snappy block encoders and
// the zstd encoder.  32 consecutive positions are hashed per step (one per lane), looked up in a
// 4096-entry table of 32-bit positions in shared memory, then inserted — when several lanes share a
// bucket the highest position wins, chosen with __match_any_sync so the result never depends on store
// ordering (the encoders are deterministic).  Every lane verifies its candidate with one 4-byte
// compare; the matches of a step are taken greedily in position order and extended 32 bytes per ballot.
#pragma once
#include "common.cuh"

namespace cj {

constexpr int ENC_HBITS = 12;
constexpr int ENC_HSIZE = 1 << ENC_HBITS;
constexpr uint32_t ENC_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t ENC_MAXOFF = 65535;

__device__ __forceinline__ uint32_t load32u(const uint8_t* p) {  // unaligned little-endian 32-bit load
    const uint32_t a = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p - a);
    const uint32_t lo = __ldg(w);
    if (a == 0) return lo;
    const uint32_t hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, a * 8);
}

__device__ __forceinline__ void match_table_reset(uint32_t* table, int lane) {
    for (uint32_t i = lane; i < ENC_HSIZE / 4; i += 32) reinterpret_cast<uint4*>(table)[i] = make_uint4(ENC_EMPTY, ENC_EMPTY, ENC_EMPTY, ENC_EMPTY);
    __syncwarp();
}

// Scans src[begin, end): positions in [begin, start_limit) may start a match, a match may not pass
// match_limit.  For every match taken, emit(anchor, literal_len, offset, match_len) is called warp-uniformly.
// Returns the position where the trailing literal run starts.
template <class Emit>
__device__ __forceinline__ uint32_t find_matches(const uint8_t* __restrict__ src, uint32_t begin, uint32_t start_limit, uint32_t match_limit,
                                                 uint32_t* table, int lane, Emit&& emit) {
    uint32_t anchor = begin;
    uint32_t p = begin;
    while (p < start_limit) {
        const uint32_t pos = p + lane;
        const bool valid = pos < start_limit;
        uint32_t v = 0, h = 0, cand = ENC_EMPTY;
        if (valid) {
            v = load32u(src + pos);
            h = (v * 0x9E3779B1u) >> (32 - ENC_HBITS);
            cand = table[h];
        }
        __syncwarp();
        const uint32_t grp = __match_any_sync(FULL, valid ? h : (0x80000000u | lane));
        if (valid && lane == 31 - __clz(grp)) table[h] = pos;  // highest position of the bucket wins
        __syncwarp();
        const bool ok = valid && cand != ENC_EMPTY && pos - cand <= ENC_MAXOFF && load32u(src + cand) == v;
        uint32_t mm = __ballot_sync(FULL, ok);
        if (anchor > p) mm &= anchor - p >= 32 ? 0u : ~((1u << (anchor - p)) - 1);  // lanes covered by the previous match
        while (mm) {
            const int i = __ffs(mm) - 1;
            const uint32_t mpos = p + i;
            const uint32_t c = __shfl_sync(FULL, cand, i);
            uint32_t len = 4;  // extend the match, 32 bytes per ballot
            const uint32_t maxlen = match_limit - mpos;
            for (;;) {
                const uint32_t k = len + lane;
                const bool eq = k < maxlen && __ldg(src + c + k) == __ldg(src + mpos + k);
                const uint32_t ne = __ballot_sync(FULL, !eq);
                if (ne) { len += __ffs(ne) - 1; break; }
                len += 32;
            }
            emit(anchor, mpos - anchor, mpos - c, len);
            anchor = mpos + len;
            const uint32_t covered = anchor - p;  // lanes below this were swallowed by the match
            mm = covered >= 32 ? 0u : mm & ~((1u << covered) - 1);
        }
        p = max(p + 32, anchor);
    }
    return anchor;
}

}  // namespace cj
