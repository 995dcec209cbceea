// common.cuh — shared device-side definitions for the block-codec kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cramjam_cuda.h"

// Kernel function attributes (dynamic shared-memory limit, carve-out) belong to the device the call is made on: launchers keep
// one flag per device, not one per process, so that a second context on another GPU sets them again.
struct cj_per_device_flag {
    int v[64] = {};
    int& here() {
        int d = 0;
        (void)cudaGetDevice(&d);
        return v[d & 63];
    }
};

namespace cj {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t MAX_UNIT = 0x7fffffffu;  // units are limited to 2^31-1 bytes on either side

// Device view of a cj_batch (all pointers are device pointers).
struct Batch {
    uint32_t n;
    const uint8_t* __restrict__ src_base;
    const uint64_t* __restrict__ src_off;
    const uint64_t* __restrict__ src_len;
    uint8_t* __restrict__ dst_base;
    const uint64_t* __restrict__ dst_off;
    const uint64_t* __restrict__ dst_cap;
    uint64_t* __restrict__ dst_len;
    int32_t* __restrict__ status;
};

__device__ __forceinline__ uint32_t ldg_u8(const uint8_t* p) { return __ldg(p); }

__device__ __forceinline__ uint32_t rd32g(const uint8_t* p) {  // unaligned little-endian u32 from read-only input
    return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24);
}

// Next unit index from the grid-wide work queue (one atomic per warp).
__device__ __forceinline__ uint32_t next_unit(unsigned* counter, int lane) {
    uint32_t v = 0;
    if (lane == 0) v = atomicAdd(counter, 1u);
    return __shfl_sync(FULL, v, 0);
}

}  // namespace cj
