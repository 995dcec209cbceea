// far_pair_probe.cu — what does a lane's SECOND 16-byte cp.async into a DRAM line cost?
//
// The thread-per-block decoder (lz_decode7.cuh) fetches the 1-2 aligned granules that hold a back-reference with one
// cp.async each.  ncu says every such request is an L2 miss with its own 64-byte DRAM read, also when the second granule
// lies in the line the first one just asked for.  This probe reproduces the access pattern (one lane per 64 KiB window,
// random lines inside it, D copies in flight per lane) and varies how the pair is requested.  Run under
// `ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct` for the traffic, plain for the times.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o variants/far_pair_probe tools/far_pair_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int D = 3;          // groups in flight per lane
constexpr int WARPS = 14;
constexpr int ITERS = 4096;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(uint32_t s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void cp16_128(uint32_t s, const void* g) { asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void cp16_cg(uint32_t s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void pf_l2(const void* g) { asm volatile("prefetch.global.L2 [%0];" ::"l"(g)); }

// MODE 0: one granule per iteration
//      1: granules g, g+1 in one 32-byte sector, back to back
//      2: granules in the two sectors of one 64-byte half line, back to back
//      3: granules in adjacent 64-byte halves of one 128-byte line, back to back
//      4: as 2, the second request one iteration after the first (retired one iteration later)
//      5: as 2, the second request two iterations after the first
//      6: as 2, first request carries .L2::128B
//      7: as 2, prefetch.global.L2 of the line one iteration ahead of both requests
//      8: as 2, both with .cg (no L1 allocation)
//      9: as 2, but the second request comes from the NEIGHBOUR lane in the same instruction as its own first (lane^1 swaps)
template <int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(const uint8_t* __restrict__ base, uint32_t* out, int spin, uint32_t seed) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t gid = (size_t)blockIdx.x * (WARPS * 32) + threadIdx.x;
    const uint8_t* win = base + gid * 65536;   // this lane's window
    // staging: (D + 2) slots x 2 granules, interleaved across lanes like the decoder's rings
    uint8_t* wsm = sm + (size_t)warp * ((D + 2) * 2 * 512);
    auto slot = [&](int s, int h) { return smem_u32(wsm + (size_t)(s * 2 + h) * 512 + lane * 16); };
    uint32_t x = seed ^ (uint32_t)(gid * 2654435761u);
    uint32_t acc = 0;
    const uint8_t* pend1 = nullptr; int pend1_slot = 0;   // second request delayed by one iteration
    const uint8_t* pend2 = nullptr; int pend2_slot = 0;   // ... by two
    const uint8_t* ahead = nullptr;
    constexpr int LAG = MODE == 4 ? 1 : (MODE == 5 ? 2 : 0);
    for (int it = 0; it < ITERS + D + LAG; it++) {
        wait<D - 1>();
        if (MODE == 9) __syncwarp();   // granules asked for by the neighbour lane
        // retire the piece issued D + LAG iterations ago
        if (it >= D + LAG) {
            const int s = (it - D - LAG) % (D + 2);
            uint4 a, b;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(slot(s, 0)));
            acc ^= a.x + a.y + a.z + a.w;
            if (MODE != 0) {
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(slot(s, 1)));
                acc ^= b.x + b.y + b.z + b.w;
            }
        }
        // "decode": a dependent chain standing in for the ~200 instructions of a decoder iteration
        for (int k = 0; k < spin; k++) x = x * 1664525u + 1013904223u + (acc & 1u);
        if (it < ITERS) {
            const int s = it % (D + 2);
            uint32_t line = (x >> 8) & 1023u;                  // a 64-byte line of the window
            if (MODE == 7) {                                   // address known one iteration ahead
                const uint8_t* nxt = win + (size_t)line * 64;
                const uint8_t* now = ahead ? ahead : nxt;
                pf_l2(nxt);
                ahead = nxt;
                cp16(slot(s, 0), now + 16);
                cp16(slot(s, 1), now + 32);
            } else {
                const uint8_t* p = win + (size_t)line * 64;
                if (MODE == 0) cp16(slot(s, 0), p);
                if (MODE == 1) { cp16(slot(s, 0), p); cp16(slot(s, 1), p + 16); }
                if (MODE == 2) { cp16(slot(s, 0), p + 16); cp16(slot(s, 1), p + 32); }
                if (MODE == 3) { const uint8_t* q = win + (size_t)(line & ~1u) * 64; cp16(slot(s, 0), q + 48); cp16(slot(s, 1), q + 64); }
                if (MODE == 4) {
                    if (pend1) cp16(slot(pend1_slot, 1), pend1);
                    cp16(slot(s, 0), p + 16);
                    pend1 = p + 32; pend1_slot = s;
                }
                if (MODE == 5) {
                    if (pend2) cp16(slot(pend2_slot, 1), pend2);
                    pend2 = pend1; pend2_slot = pend1_slot;
                    cp16(slot(s, 0), p + 16);
                    pend1 = p + 32; pend1_slot = s;
                }
                if (MODE == 6) { cp16_128(slot(s, 0), p + 16); cp16(slot(s, 1), p + 32); }
                if (MODE == 8) { cp16_cg(slot(s, 0), p + 16); cp16_cg(slot(s, 1), p + 32); }
                if (MODE == 9) {
                    // instruction 1: even lanes ask for their own first granule, odd lanes for the even neighbour's second;
                    // instruction 2: the other way round.  Each instruction then holds both sectors of a line.
                    const uint64_t pn = __shfl_xor_sync(0xffffffffu, (uint64_t)(uintptr_t)p, 1);
                    const uint8_t* q = (const uint8_t*)(uintptr_t)pn;
                    const uint32_t my0 = slot(s, 0), my1 = slot(s, 1);
                    const uint32_t nb0 = __shfl_xor_sync(0xffffffffu, my0, 1), nb1 = __shfl_xor_sync(0xffffffffu, my1, 1);
                    (void)nb0;
                    const bool even = (lane & 1) == 0;
                    cp16(even ? my0 : nb1, even ? p + 16 : q + 32);     // the even lane's pair
                    cp16(even ? nb1 : my0, even ? q + 32 : p + 16);     // the odd lane's pair
                }
            }
        } else if (MODE == 4) {
            if (pend1) { cp16(slot(pend1_slot, 1), pend1); pend1 = nullptr; }
        } else if (MODE == 5) {
            if (pend2) cp16(slot(pend2_slot, 1), pend2);
            pend2 = pend1; pend2_slot = pend1_slot; pend1 = nullptr;
        }
        commit();
    }
    wait<0>();
    out[gid] = acc ^ x;
}

template <int MODE>
static void run(const uint8_t* buf, uint32_t* out, int sms, int spin, const char* what) {
    const size_t smem = (size_t)WARPS * (D + 2) * 2 * 512;
    CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e9f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(a));
        probe<MODE><<<sms - 1, WARPS * 32, smem>>>(buf, out, spin, 12345u + r);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double pieces = (double)(sms - 1) * WARPS * 32 * ITERS;
    printf("mode %d spin %3d: %7.3f ms  %6.1f ns/iteration/lane-set  %7.1f M pieces  -> 64 B x pieces = %.2f GB   %s\n", MODE, spin, best,
           best * 1e6 / ITERS, pieces / 1e6, pieces * 64 / 1e9, what);
}

int main(int argc, char** argv) {
    const int spin = argc > 1 ? atoi(argv[1]) : 60;
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t lanes = (size_t)(sms - 1) * WARPS * 32;
    uint8_t* buf; uint32_t* out;
    CK(cudaMalloc(&buf, lanes * 65536));
    CK(cudaMemset(buf, 1, lanes * 65536));
    CK(cudaMalloc(&out, lanes * 4));
    run<0>(buf, out, sms, spin, "one granule");
    run<1>(buf, out, sms, spin, "pair in one 32 B sector");
    run<2>(buf, out, sms, spin, "pair in one 64 B half line");
    run<3>(buf, out, sms, spin, "pair across the halves of a 128 B line");
    run<4>(buf, out, sms, spin, "as 2, second request one iteration later");
    run<5>(buf, out, sms, spin, "as 2, second request two iterations later");
    run<6>(buf, out, sms, spin, "as 2, .L2::128B on the first");
    run<7>(buf, out, sms, spin, "as 2, prefetch.global.L2 one iteration ahead");
    run<8>(buf, out, sms, spin, "as 2, .cg");
    run<9>(buf, out, sms, spin, "as 2, both sectors of a line in ONE instruction (lane pairs swap)");
    CK(cudaDeviceSynchronize());
    return 0;
}
