// tests/emu/g6_emu.cpp — host emulation of the generation-6 decoder (cramjam_b200/csrc/lz_decode6.cuh) for the CPU test
// suite: the walk (host restatement) produces the checkpoints, then T emulated lanes run range_step() in lock step over the
// block's window exactly as the lanes of the EXEC kernel do (same code, compiled for the host).  TEST INFRASTRUCTURE.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "lz_decode6.cuh"

using namespace cj::g6;

template <int CODEC>
static long run(const uint8_t* src, uint32_t n, uint8_t* dst, uint64_t cap, int T, long* iters, long* moved) {
    std::vector<uint32_t> ck(2 * NR + 2, 0xdeadbeefu);
    WalkOut w{ck.data()};
    const uint32_t ulen = CODEC == 0 ? snappy_walk_host(src, n, cap, 1u << 20, w) : lz4_walk_host(src, n, cap, 1u << 20, w);
    if (!ulen) return 0;
    std::vector<uint64_t> win((MAXU + 64) / 8, 0xEEEEEEEEEEEEEEEEull), in((n + 64) / 8 + 1, 0);
    memcpy(in.data(), src, n);
    std::vector<uint32_t> bits(MAXU / 32 + 2, 0);
    Block b{(uint8_t*)win.data(), (const uint8_t*)in.data(), bits.data(), ck.data(), n, ulen};
    const uint32_t nr = (ulen + R - 1) / R;
    std::vector<Range> st(T);
    std::vector<int> cur(T, -1);      // range a lane works on, -1 = needs one
    uint32_t next = 0, done = 0;
    long it = 0, mv = 0;
    while (done < nr) {
        if (++it > 50000000) return -1;   // livelock guard
        for (int l = 0; l < T; l++) {
            if (cur[l] < 0) {
                if (next >= nr) continue;
                cur[l] = (int)next++;
                range_start<CODEC>(st[l], b, (uint32_t)cur[l]);
            }
            const uint32_t before = st[l].o;
            if (range_step<CODEC>(st[l], b, (uint32_t)cur[l]) == STEP_DONE) { cur[l] = -1; done++; }
            else if (st[l].o != before) mv++;
        }
    }
    for (uint32_t p = 0; p < ulen; p++)
        if (!((bits[p >> 5] >> (p & 31)) & 1)) return -2;   // every byte must have been announced
    memcpy(dst, win.data(), ulen);
    if (iters) *iters = it;
    if (moved) *moved = mv;
    return ulen;
}

extern "C" long g6_emu_decode(int codec, const uint8_t* src, uint32_t n, uint8_t* dst, uint64_t cap, int lanes, long* iters, long* moved) {
    return codec == 0 ? run<0>(src, n, dst, cap, lanes, iters, moved) : run<2>(src, n, dst, cap, lanes, iters, moved);
}
