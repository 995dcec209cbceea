// util_kernels.cu — small data-movement kernels around the codec kernels: batched unit copy
// (packing compressed blocks into a dense arena, building / splitting frame containers).
#include "common.cuh"

namespace cj {

// One warp per unit: dst_base[dst_off[i] .. +len[i]) = src_base[src_off[i] .. +len[i]).
// 16-byte vector body when source and destination are congruent modulo 16, else byte-wise.
__global__ void __launch_bounds__(256) copy_units_kernel(uint32_t n, const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                                                         const uint64_t* __restrict__ len, uint8_t* __restrict__ dst_base,
                                                         const uint64_t* __restrict__ dst_off) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n; u += warps) {
        const uint8_t* s = src_base + src_off[u];
        uint8_t* d = dst_base + dst_off[u];
        const uint64_t L = len[u];
        if ((((uintptr_t)s ^ (uintptr_t)d) & 15) == 0) {
            uint64_t head = (16 - ((uintptr_t)s & 15)) & 15;
            if (head > L) head = L;
            if ((uint64_t)lane < head) d[lane] = s[lane];
            const uint64_t body = (L - head) & ~(uint64_t)15;
            const uint4* s4 = reinterpret_cast<const uint4*>(s + head);
            uint4* d4 = reinterpret_cast<uint4*>(d + head);
            for (uint64_t i = lane; i < body / 16; i += 32) d4[i] = __ldg(s4 + i);
            const uint64_t t0 = head + body;
            if (t0 + lane < L) d[t0 + lane] = s[t0 + lane];
        } else {
            for (uint64_t i = lane; i < L; i += 32) d[i] = s[i];
        }
    }
}

cudaError_t launch_copy_units(uint32_t n, const uint8_t* src_base, const uint64_t* src_off, const uint64_t* len, uint8_t* dst_base,
                              const uint64_t* dst_off, int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    int grid = (int)((n + 7) / 8);
    if (grid > sm_count * 8) grid = sm_count * 8;
    copy_units_kernel<<<grid, 256, 0, stream>>>(n, src_base, src_off, len, dst_base, dst_off);
    return cudaGetLastError();
}

}  // namespace cj
