// synth.cuh — deterministic "Silesia-like" block generator shared by host and device code.
//
// an LZ77-style source model (skewed literal runs alternating with
// back-references into the block's own history), ~8 % incompressible and ~8 % highly repetitive
// blocks mixed in, every block a pure function of (seed, global block index) so any rank count
// and the host CPU baseline see identical bytes.  Tuned (tools/tune_synth.py) so that the CPU
// encoders land near the Silesia aggregates at 64 KiB: snappy/lz4 ratio ~2.0-2.1, ~8 K elements
// per block, mean literal run ~6 B, mean copy ~9 B.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CJ_HD __host__ __device__ __forceinline__
#else
#define CJ_HD static inline
#endif

namespace cj {

struct SynthRng {
    uint64_t s;
    CJ_HD uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

// Skewed literal symbol: text-like 96-symbol alphabet, roughly 5 bits/byte of entropy.
CJ_HD uint8_t synth_symbol(uint32_t r) {
    uint32_t x = r & 0xff, y = (r >> 8) & 0xff, z = (r >> 16) & 0xff;
    uint32_t v = (x * y * z) >> 16;  // 0..252, heavily skewed towards small values
    v = (v * 96) >> 8;               // 0..94
    return (uint8_t)(32 + v);
}

// Fills out[0..len) for the block with global index `index`.
CJ_HD void synth_block(uint8_t* out, size_t len, uint64_t seed, uint64_t index) {
    SynthRng g;
    g.s = seed ^ (index * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull);
    const uint32_t cls = (uint32_t)(g.next() % 100u);
    size_t pos = 0;
    if (cls < 8) {  // incompressible (x-ray / sao like)
        while (pos < len) {
            uint64_t r = g.next();
            for (int k = 0; k < 8 && pos < len; k++, r >>= 8) out[pos++] = (uint8_t)r;
        }
        return;
    }
    const bool rep = cls < 16;  // highly repetitive (nci / xml like)
    while (pos < len) {
        uint64_t r = g.next();
        // literal run
        uint32_t ll;
        if (rep) ll = 1 + (uint32_t)(r & 3);
        else if (((r >> 8) & 15) == 0) ll = 1 + (uint32_t)((r >> 12) % 48);
        else if (((r >> 8) & 15) < 10) ll = 0;  // match follows match directly (no literals)
        else ll = 1 + (uint32_t)((r >> 12) % 7);
        uint64_t lr = 0;
        for (uint32_t i = 0; i < ll && pos < len; i++) {
            if ((i & 1) == 0) lr = g.next();
            out[pos++] = synth_symbol((uint32_t)(lr >> ((i & 1) * 32)));
        }
        if (pos >= len) break;
        if (pos < 8) continue;
        // back-reference
        uint32_t ml;
        if (rep) ml = 16 + (uint32_t)((r >> 20) % 120);
        else if (((r >> 20) & 31) == 0) ml = 8 + (uint32_t)((r >> 26) % 160);
        else ml = 5 + (uint32_t)((r >> 26) % 13);
        // offset classes after the measured Silesia distribution: ~24 % within 256 B,
        // ~41 % within 4 KiB, ~35 % beyond; about 1 % short enough to overlap the copy itself
        const uint32_t oc = (uint32_t)((r >> 40) % 100);
        const uint32_t orr = (uint32_t)(r >> 47);
        uint32_t off;
        if (oc < 1) off = 1 + orr % 8;
        else if (oc < 24) off = 16 + orr % 240;
        else if (oc < 65) off = 256 + orr % 3840;
        else off = 4096 + orr % 61440;
        if (off > pos) off = 1 + (off - 1) % (uint32_t)pos;
        for (uint32_t i = 0; i < ml && pos < len; i++, pos++) out[pos] = out[pos - off];
    }
}

}  // namespace cj
