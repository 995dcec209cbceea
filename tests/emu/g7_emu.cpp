// g7_emu.cpp — host-side emulation of the generation-7 lane program (cramjam_b200/csrc/lz_decode7.cuh), test infrastructure.
//
// The lane program is written against an `Env` of memory operations.  On the device these are shared-memory loads/stores,
// cp.async and predicated global stores; here they act on a byte array standing in for the lane's shared-memory record,
// with every asynchronous copy either delivered at once (mode 0) or at the last moment the program's own wait_group allows
// (mode 1).  A program that is correct in both modes neither reads a copy before it is guaranteed to have landed nor
// overwrites ring bytes that a chunk in flight still needs.  Alignment of every access, the bounds of the record, and "a far
// fetch only reads output that has been stored" are checked on the way.  tests/test_g7_emu.py drives it against the oracle.
#include <stdint.h>
#include <string.h>

#include <deque>
#include <vector>

#include "../../cramjam_b200/csrc/lz_decode7.cuh"

namespace {
using cj::g7::u4;

struct HostEnv {
    std::vector<uint8_t> smem;
    uint32_t in_l, out_l, st_l, lut;
    int mode;
    struct Cp {
        uint32_t saddr;
        uint8_t data[16];
    };
    std::deque<std::vector<Cp>> groups;
    std::vector<Cp> cur;
    uint8_t* dst_base = nullptr;
    const uint8_t* src_base = nullptr;
    uint64_t src_n = 0, dst_cap = 0;
    uint64_t stored = 0;       // output bytes stored to "global memory" so far (granules arrive in order)
    long iters = 0, bubbles = 0, far_fetches = 0;
    int error = 0;             // first violated invariant
    bool redo_flag = false, done = false;
    uint32_t out_len = 0;
    int extra = 0;             // iterations the lane keeps step with its warp after it is done

    void fail(int code) {
        if (!error) error = code;
    }
    bool ok_range(uint32_t a, uint32_t n, uint32_t align) {
        if ((a & (align - 1)) != 0) { fail(1); return false; }
        if ((uint64_t)a + n > smem.size()) { fail(2); return false; }
        if (a < lut && (a % cj::g7::GROW) + n > 16) { fail(10); return false; }   // a lane stays inside its own 16-byte column of the granule rows
        return true;
    }
    uint32_t lds32(uint32_t a) {
        if (!ok_range(a, 4, 4)) return 0;
        uint32_t v;
        memcpy(&v, &smem[a], 4);
        return v;
    }
    uint32_t lds8(uint32_t a) {
        if (!ok_range(a, 1, 1)) return 0;
        return smem[a];
    }
    void sts32(uint32_t a, uint32_t v) {
        if (!ok_range(a, 4, 4)) return;
        memcpy(&smem[a], &v, 4);
    }
    u4 lds128(uint32_t a) {
        u4 v = {0, 0, 0, 0};
        if (!ok_range(a, 16, 16)) return v;
        memcpy(&v, &smem[a], 16);
        return v;
    }
    void tick() { iters++; }
    void sts128(uint32_t a, u4 v) { sts128_if(a, v, true); }
    void sts128_if(uint32_t a, u4 v, bool p) {
        if (!p) return;
        if (!ok_range(a, 16, 16)) return;
        memcpy(&smem[a], &v, 16);
    }
    void stg128_if(uint8_t* p, u4 v, bool pred) {
        if (!pred) return;
        if (done || redo_flag) { /* chunks in flight of a declined block still retire; they stay inside the capacity */ }
        const uint64_t o = (uint64_t)(p - dst_base);
        if (((uintptr_t)p & 15u) != 0 || o + 16 > dst_cap) { fail(3); return; }
        if (o != stored) fail(4);   // granules leave in order, each exactly once
        memcpy(p, &v, 16);
        stored = o + 16;
    }
    void stg256_if(uint8_t* p, u4 a, u4 c, bool pred) {
        if (!pred) return;
        const uint64_t o = (uint64_t)(p - dst_base);
        if (((uintptr_t)p & 31u) != 0 || o + 32 > dst_cap) { fail(3); return; }
        if (o != stored) fail(4);
        memcpy(p, &a, 16);
        memcpy(p + 16, &c, 16);
        stored = o + 32;
    }
    void stg8(uint8_t* p, uint32_t v) {
        const uint64_t o = (uint64_t)(p - dst_base);
        if (o >= dst_cap) { fail(5); return; }
        *p = (uint8_t)v;
    }
    uint32_t ldg8(const uint8_t* p) {
        const uint64_t o = (uint64_t)(p - src_base);
        if (o >= src_n) { fail(6); return 0; }
        return *p;
    }
    void push(uint32_t saddr, const uint8_t* data16) {
        Cp c;
        c.saddr = saddr;
        memcpy(c.data, data16, 16);
        if (mode == 0) memcpy(&smem[saddr], c.data, 16);
        else cur.push_back(c);
    }
    void cp16_far_if(uint32_t saddr, const uint8_t* g, bool pred) {
        if (!pred) return;
        if (!ok_range(saddr, 16, 16)) return;
        const uint64_t o = (uint64_t)(g - dst_base);
        if (((uintptr_t)g & 15u) != 0 || o + 16 > stored) { fail(7); return; }   // only output that is in global memory already
        far_fetches++;
        push(saddr, g);
    }
    void cp16_in_if(uint32_t saddr, const uint8_t* g, uint32_t ssz, bool pred) {
        if (!pred) return;
        if (!ok_range(saddr, 16, 16)) return;
        const uint64_t o = (uint64_t)(g - src_base);
        if (((uintptr_t)g & 15u) != 0 || ssz > 16 || o + ssz > src_n) { fail(8); return; }
        uint8_t tmp[16] = {0};
        memcpy(tmp, g, ssz);
        push(saddr, tmp);
    }
    void commit() {
        if (mode != 0) {
            groups.push_back(cur);
            cur.clear();
        }
    }
    template <int N>
    void wait() {
        while ((int)groups.size() > N) {
            for (const Cp& c : groups.front()) memcpy(&smem[c.saddr], c.data, 16);
            groups.pop_front();
        }
    }
    bool any(bool active) {
        if (iters > 64 + 8 * (long)(src_n + dst_cap)) { fail(9); return false; }   // a lane that stops making progress
        if (active) return true;
        return extra-- > 0;
    }
    void redo() { redo_flag = true; }
    void finish_ok(uint32_t len) {
        done = true;
        out_len = len;
    }
};

template <int CODEC, int D>
long run(const uint8_t* src, uint32_t n, uint8_t* dst, uint64_t cap, int mode, int extra, long* stats) {
    HostEnv env;
    const uint32_t rec = cj::g7::warp_bytes(D);   // the emulated lane is lane 0 of its warp; the other lanes' granules stay untouched
    env.smem.resize(rec + 1024);
    uint32_t seed = 0x1234567u + n;
    for (auto& b : env.smem) { seed = seed * 1664525u + 1013904223u; b = (uint8_t)(seed >> 24); }   // nothing may rely on initial contents
    env.in_l = 0;
    env.out_l = cj::g7::IN_G * cj::g7::GROW;
    env.st_l = (cj::g7::IN_G + cj::g7::OUT_G) * cj::g7::GROW;
    env.lut = rec;
    for (uint32_t t = 0; t < 256; t++) {
        const uint32_t e = cj::g7::tag_entry(t);
        memcpy(&env.smem[env.lut + 4 * t], &e, 4);
    }
    env.mode = mode;
    env.extra = extra;
    env.dst_base = dst;
    env.src_base = src;
    env.src_n = n;
    env.dst_cap = cap;
    cj::g7::decode_block<CODEC, D>(env, true, src, dst, n, cap);
    if (stats) { stats[0] = env.iters; stats[1] = env.far_fetches; stats[2] = env.error; }
    if (env.error) return -100 - env.error;
    if (env.redo_flag) return -1;
    if (!env.done) return -2;
    return (long)env.out_len;
}
}  // namespace

// Returns the decoded length, -1 if the lane handed the block to the redo list, < -100 if an invariant of the emulated
// machine was violated.  src and dst must be 16-byte aligned; dst must hold cap bytes.
extern "C" long g7_emu_decode(int codec, int depth, const uint8_t* src, uint32_t n, uint8_t* dst, uint64_t cap, int mode, int extra, long* stats) {
    if (codec == cj::g7::CODEC_SNAPPY) {
        if (depth == 2) return run<cj::g7::CODEC_SNAPPY, 2>(src, n, dst, cap, mode, extra, stats);
        if (depth == 3) return run<cj::g7::CODEC_SNAPPY, 3>(src, n, dst, cap, mode, extra, stats);
        return run<cj::g7::CODEC_SNAPPY, 4>(src, n, dst, cap, mode, extra, stats);
    }
    if (depth == 2) return run<cj::g7::CODEC_LZ4, 2>(src, n, dst, cap, mode, extra, stats);
    if (depth == 3) return run<cj::g7::CODEC_LZ4, 3>(src, n, dst, cap, mode, extra, stats);
    return run<cj::g7::CODEC_LZ4, 4>(src, n, dst, cap, mode, extra, stats);
}
