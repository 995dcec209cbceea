// lz_decode7.cu — generation-7 batch decode of Snappy raw blocks and LZ4 blocks (sm_100a): one THREAD per block, 16-byte chunks,
// linear per-lane shared-memory records.  The lane program is lz_decode7.cuh (the same source runs lane by lane on the host
// in tests/emu/g7_emu.cpp); this file supplies its memory operations as PTX, the kernel around it and the launcher.
//
// Same reference entry points as lz_decode4.cu: snap::raw::Decoder::decompress behind cramjam.snappy.decompress_raw /
// decompress_raw_into (src/snappy.rs:52-60,102-108); LZ4_decompress_safe behind lz4::block::decompress_into
// (src/lz4.rs:78-95,140-173).  Units the lane program declines go on the redo list of the warp-per-block kernel
// (lz_decode.cuh), which owns every status code.
#include "internal.h"
#include "lz_decode.cuh"
#include "lz_decode7.cuh"

namespace cj {

#ifndef CJ_G7_POL
#define CJ_G7_POL 3   // L2 eviction hints: bit 0 far fetches evict-first, bit 1 output stores evict-last (together 7.93 -> 7.74 ms on host-made Snappy streams, DRAM reads -6 %; LZ4 unchanged), bit 2 input evict-first (8.8-9.3 ms: a line of input is read over several passes)
#endif
constexpr int G7_MAX_WARPS = 20;   // most warps per CTA (640 threads x 102 registers); fewer if the lane records of 20 warps do not fit the SM's shared memory
constexpr int g7_max_warps(int D) { return (int)((232448u - 1024u) / g7::warp_bytes(D)) < G7_MAX_WARPS ? (int)((232448u - 1024u) / g7::warp_bytes(D)) : G7_MAX_WARPS; }

struct G7 {
    uint32_t* redo_list;   // units for the generation-2 kernel
    unsigned* ctr;         // [1] redo count  [2] redo work queue
};

// Memory operations of the lane program.  Shared-memory accesses are volatile asm on 32-bit shared addresses and keep their
// program order; predicated forms are predicated PTX (no branch around them).
struct G7Env {
    uint32_t in_l, out_l, st_l, lut;
    const Batch& b;
    const G7& g;
    uint32_t cur;
#if CJ_G7_POL
    uint64_t pol_first, pol_last;   // L2 eviction policies (createpolicy): far / input lines leave first, the block's own output stays
#endif
    __device__ __forceinline__ G7Env(const Batch& b_, const G7& g_) : b(b_), g(g_) {}
    __device__ __forceinline__ void tick() const {}
    __device__ __forceinline__ uint32_t lds32(uint32_t a) const { return cj::lds32(a); }
    __device__ __forceinline__ uint32_t lds8(uint32_t a) const { return cj::lds8(a); }
    __device__ __forceinline__ g7::u4 lds128(uint32_t a) const {
        g7::u4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
        return v;
    }
    __device__ __forceinline__ void sts128(uint32_t a, g7::u4 v) const {
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    __device__ __forceinline__ void sts128_if(uint32_t a, g7::u4 v, bool p) const {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n\t}" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((uint32_t)p) : "memory");
    }
    __device__ __forceinline__ void stg128_if(uint8_t* p, g7::u4 v, bool pred) const {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.global.v4.u32 [%0], {%1,%2,%3,%4};\n\t}" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((uint32_t)pred) : "memory");
    }
    // two finished granules that make up one 32-byte sector leave with one 256-bit store: half as many write transactions
    __device__ __forceinline__ void stg256_if(uint8_t* p, g7::u4 a, g7::u4 c, bool pred) const {
#if CJ_G7_POL & 2
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %9, 0;\n\t@p st.global.L2::cache_hint.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %10;\n\t}" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w), "r"((uint32_t)pred), "l"(pol_last) : "memory");
        return;
#endif
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %9, 0;\n\t@p st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n\t}" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w), "r"((uint32_t)pred) : "memory");
    }
    // Far source granule.  cp.async.ca (LDGSTS, L1-allocating) reads output bytes that the SAME lane stored earlier with
    // st.global.v4 (stg128_if) — never another thread's.  The store precedes the fetch in program order, both are volatile asm with
    // a memory clobber and are issued by one warp through one LSU queue; the L1 is write-through and a store updates or evicts the
    // line it hits, which is what makes an ordinary ld.global after st.global by the same thread return the stored value, and
    // LDGSTS performs the same L1 lookup as LDG.  (The argument, and its evidence, are those of lz_decode4.cu.)
    __device__ __forceinline__ void cp16_far_if(uint32_t saddr, const uint8_t* gptr, bool pred) const {
#if CJ_G7_POL & 1
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %3;\n\t}" ::"r"(saddr), "l"(gptr), "r"((uint32_t)pred), "l"(pol_first) : "memory");
        return;
#endif
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 16;\n\t}" ::"r"(saddr), "l"(gptr), "r"((uint32_t)pred) : "memory");
    }
    __device__ __forceinline__ void stg8(uint8_t* p, uint32_t v) const { *p = (uint8_t)v; }
    __device__ __forceinline__ uint32_t ldg8(const uint8_t* p) const { return ldg_u8(p); }
    __device__ __forceinline__ void cp16_in_if(uint32_t saddr, const uint8_t* gptr, uint32_t ssz, bool pred) const {
#if CJ_G7_POL & 4
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %4;\n\t}" ::"r"(saddr), "l"(gptr), "r"(ssz), "r"((uint32_t)pred), "l"(pol_first) : "memory");
        return;
#endif
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16, %2;\n\t}" ::"r"(saddr), "l"(gptr), "r"(ssz), "r"((uint32_t)pred) : "memory");
    }
    __device__ __forceinline__ void commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
    template <int N>
    __device__ __forceinline__ void wait() const { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
    __device__ __forceinline__ bool any(bool active) const { return __any_sync(FULL, active); }
    __device__ __forceinline__ void redo() const {
        const unsigned i = atomicAdd(&g.ctr[1], 1u);
        g.redo_list[i] = cur;
    }
    __device__ __forceinline__ void finish_ok(uint32_t len) const {
        b.dst_len[cur] = len;
        b.status[cur] = CJ_OK;
    }
};

template <int CODEC, int D>
__global__ void __launch_bounds__(G7_MAX_WARPS * 32, 1) g7_kernel(Batch b, G7 g) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warps = blockDim.x >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    G7Env env(b, g);
#if CJ_G7_POL
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(env.pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(env.pol_last));
#endif
    const uint32_t wbase = smem_addr(smem) + (uint32_t)warp * g7::warp_bytes(D);
    env.in_l = wbase + (uint32_t)lane * 16;                                       // granule q of the lane's input ring: in_l + q * 512
    env.out_l = wbase + g7::IN_G * g7::GROW + (uint32_t)lane * 16;                // ... of its output ring: out_l + q * 512
    env.st_l = wbase + (g7::IN_G + g7::OUT_G) * g7::GROW + (uint32_t)lane * 16;   // ... of the staging pair of slot u: st_l + (2u + q) * 512
    env.lut = smem_addr(smem) + (uint32_t)warps * g7::warp_bytes(D);
    if (CODEC == CJ_SNAPPY_RAW) {
        for (uint32_t t = threadIdx.x; t < 256; t += blockDim.x) sts32(env.lut + 4 * t, g7::tag_entry(t));
        __syncthreads();
    }
    const uint32_t nwarps = gridDim.x * warps;
    for (uint32_t first = (blockIdx.x * warps + warp) * 32; first < b.n; first += nwarps * 32) {
        const uint32_t cur = first + lane;
        const bool has = cur < b.n;
        env.cur = cur;
        const uint8_t* src = nullptr;
        uint8_t* dst = nullptr;
        uint64_t sl = 0, dcap = 0;
        if (has) {
            sl = b.src_len[cur];
            dcap = b.dst_cap[cur];
            src = b.src_base + b.src_off[cur];
            dst = b.dst_base + b.dst_off[cur];
        }
        g7::decode_block<CODEC, D>(env, has, src, dst, sl, dcap);
        __syncwarp();
    }
}

cudaError_t launch_lz_decode_list(int codec, const Batch& b, uint32_t* redo_list, unsigned* ctr, int sm_count, cudaStream_t stream);

template <int CODEC, int D>
static cudaError_t launch_g7(const Batch& b, const G7& g, int sm_count, cudaStream_t stream) {
    static cj_per_device_flag attr_flag;
    int& attr_done = attr_flag.here();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(g7_kernel<CODEC, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g7::cta_bytes(D, g7_max_warps(D)));
        if (e != cudaSuccess) return e;
        attr_done = 1;
    }
    // One CTA per SM with as many warps as the batch needs (a lane per block); batches beyond sm_count x G7_MAX_WARPS x 32 blocks
    // are decoded in several equal rounds by the same CTAs.
    const size_t warps = ((size_t)b.n + 31) / 32;
    const char* we = getenv("CJ_G7_WARPS");   // experiments: warps per CTA (read per launch)
    const int force_w = we ? atoi(we) : 0;
    constexpr int MAXW = g7_max_warps(D);
    const size_t rounds = std::max<size_t>(1, (warps + (size_t)sm_count * MAXW - 1) / ((size_t)sm_count * MAXW));
    int w = (int)std::min<size_t>(MAXW, std::max<size_t>(1, (warps + (size_t)sm_count * rounds - 1) / ((size_t)sm_count * rounds)));
    if (force_w >= 1 && force_w <= MAXW) w = force_w;
    const int grid = (int)std::min<size_t>((warps + w - 1) / w, (size_t)sm_count * (size_t)std::max(1, MAXW / w));
    const size_t smem = (size_t)g7::cta_bytes(D, w);
    // The L1 share of the SM's 256 KB matters (re-reads of back-reference sectors hit it): ask for the smallest
    // shared-memory carve-out that holds the CTAs of one SM.
    static cj_per_device_flag carve_flag;
    int& carve_for_w = carve_flag.here();
    if (carve_for_w != w) {
        const int per_sm = (grid + sm_count - 1) / sm_count;
        const int pct = (int)std::min<size_t>(100, ((smem + 1024) * per_sm * 100 + 228 * 1024 - 1) / (228 * 1024));
        cudaError_t e = cudaFuncSetAttribute(g7_kernel<CODEC, D>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        if (e != cudaSuccess) return e;
        carve_for_w = w;
    }
    g7_kernel<CODEC, D><<<grid, w * 32, smem, stream>>>(b, g);
    return cudaGetLastError();
}

cudaError_t launch_lz_decode7(int codec, const Batch& b, LzScratch& sc, int sm_count, cudaStream_t stream) {
    const char* de = getenv("CJ_G7_D");   // experiments: chunks in flight per lane (read per launch)
    const int depth = de ? atoi(de) : 3;   // measured at 65 536 x 64 KiB: Snappy 6.5 ms at 3 (2: 6.7, 4: 7.0) on GPU-made streams, 8.0 / 8.0 / 8.4 on host-made ones; LZ4 8.2 ms at 3 on host-made streams (2: 8.4, 4: 8.6), 8.1 vs 8.0 at 4 on GPU-made ones
    const size_t n = b.n;
    if (sc.ensure_fixed((n + 8) * 4 + 64) != 0) return cudaErrorMemoryAllocation;
    G7 g;
    g.ctr = (unsigned*)sc.fixed();
    g.redo_list = (uint32_t*)sc.fixed() + 8;
    cudaError_t e = cudaMemsetAsync(g.ctr, 0, 4 * sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    if (codec == CJ_LZ4_BLOCK) e = depth <= 2 ? launch_g7<CJ_LZ4_BLOCK, 2>(b, g, sm_count, stream) : (depth == 3 ? launch_g7<CJ_LZ4_BLOCK, 3>(b, g, sm_count, stream) : launch_g7<CJ_LZ4_BLOCK, 4>(b, g, sm_count, stream));
    else e = depth <= 2 ? launch_g7<CJ_SNAPPY_RAW, 2>(b, g, sm_count, stream) : (depth == 3 ? launch_g7<CJ_SNAPPY_RAW, 3>(b, g, sm_count, stream) : launch_g7<CJ_SNAPPY_RAW, 4>(b, g, sm_count, stream));
    if (e != cudaSuccess) return e;
    return launch_lz_decode_list(codec, b, g.redo_list, g.ctr, sm_count, stream);
}

}  // namespace cj
