// lz_decode.cu — batch kernels for independent LZ4 blocks / Snappy raw blocks (device code: lz_decode.cuh).
#include "lz_decode.cuh"

namespace cj {


template <int CODEC, bool FAST>
__global__ void __launch_bounds__(DEC_WARPS * 32, CJ_DEC_CTAS) lz_decode_kernel(Batch b, unsigned* __restrict__ counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* smem_warp = smem + (size_t)warp * DEC_SMEM_WARP;
    ring_barrier_init(smem_warp, lane);
    for (;;) {
        const uint32_t u = next_unit(counter, lane);
        if (u >= b.n) break;
        const uint64_t slen = b.src_len[u], dcap = b.dst_cap[u];
        const uint8_t* src = b.src_base + b.src_off[u];
        uint8_t* dst = b.dst_base + b.dst_off[u];
        uint32_t produced = 0;
        int32_t st;
        if (slen > MAX_UNIT) {
            st = CJ_ST_TOO_BIG;
        } else {
            const uint32_t cap = dcap > MAX_UNIT ? MAX_UNIT : (uint32_t)dcap;
            st = decode_block<CODEC, FAST>(src, (uint32_t)slen, dst, cap, smem_warp, lane, &produced);
        }
        if (lane == 0) {
            b.dst_len[u] = st == CJ_OK ? produced : 0;
            b.status[u] = st;
        }
        __syncwarp();
    }
}

template <int CODEC, bool FAST>
static cudaError_t launch_one(const Batch& b, unsigned* counter, int sm_count, cudaStream_t stream, bool reset_counter) {
    const size_t smem = (size_t)DEC_SMEM_WARP * DEC_WARPS;
    auto k = lz_decode_kernel<CODEC, FAST>;
    static cj_per_device_flag attr_flag;
    int& attr_done = attr_flag.here();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = 1;
    }
    int grid = sm_count * CJ_DEC_CTAS;  // CTAs per SM x 4 warps x (ORING + IRING + 256 B) of shared memory
    const int need = (int)((b.n + DEC_WARPS - 1) / DEC_WARPS);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (reset_counter) {  // the pinned-arena pipeline zeroes all of its counters once, up front, to keep copy engines off this stream
        cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned), stream);
        if (e != cudaSuccess) return e;
    }
    k<<<grid, DEC_WARPS * 32, smem, stream>>>(b, counter);
    return cudaGetLastError();
}

// The same decoder over a redo list filled by the thread-per-block path (lz_decode4.cu): ctr[1] = entries, ctr[2] = work queue.
template <int CODEC>
__global__ void __launch_bounds__(DEC_WARPS * 32, CJ_DEC_CTAS) lz_decode_list_kernel(Batch b, const uint32_t* __restrict__ redo_list, unsigned* ctr) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* smem_warp = smem + (size_t)warp * DEC_SMEM_WARP;
    ring_barrier_init(smem_warp, lane);
    const uint32_t count = ctr[1];
    for (;;) {
        const uint32_t i = next_unit(&ctr[2], lane);
        if (i >= count) break;
        const uint32_t u = redo_list[i];
        const uint64_t slen = b.src_len[u], dcap = b.dst_cap[u];
        uint32_t produced = 0;
        int32_t st;
        if (slen > MAX_UNIT) st = CJ_ST_TOO_BIG;
        else st = decode_block<CODEC, true>(b.src_base + b.src_off[u], (uint32_t)slen, b.dst_base + b.dst_off[u], dcap > MAX_UNIT ? MAX_UNIT : (uint32_t)dcap, smem_warp, lane, &produced);
        if (lane == 0) {
            b.dst_len[u] = st == CJ_OK ? produced : 0;
            b.status[u] = st;
        }
        __syncwarp();
    }
}

cudaError_t launch_lz_decode_list(int codec, const Batch& b, uint32_t* redo_list, unsigned* ctr, int sm_count, cudaStream_t stream) {
    const size_t smem = (size_t)DEC_SMEM_WARP * DEC_WARPS;
    static cj_per_device_flag attr_flag;
    int& attr_done = attr_flag.here();
    if (!attr_done) {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(lz_decode_list_kernel<CJ_SNAPPY_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(lz_decode_list_kernel<CJ_LZ4_BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        attr_done = 1;
    }
    const int grid = (int)std::min<size_t>(((size_t)b.n + DEC_WARPS - 1) / DEC_WARPS, (size_t)sm_count * CJ_DEC_CTAS);
    if (codec == CJ_LZ4_BLOCK) lz_decode_list_kernel<CJ_LZ4_BLOCK><<<grid, DEC_WARPS * 32, smem, stream>>>(b, redo_list, ctr);
    else lz_decode_list_kernel<CJ_SNAPPY_RAW><<<grid, DEC_WARPS * 32, smem, stream>>>(b, redo_list, ctr);
    return cudaGetLastError();
}

cudaError_t launch_lz_decode(int codec, const Batch& b, unsigned* counter, int sm_count, cudaStream_t stream, bool reset_counter) {
    // CJ_DECODE_SERIAL=1 selects the generation-1 warp-serial kernel (kept for A/B measurements).
    static const bool serial = [] { const char* e = getenv("CJ_DECODE_SERIAL"); return e && e[0] == '1'; }();
    if (codec == CJ_LZ4_BLOCK)
        return serial ? launch_one<CJ_LZ4_BLOCK, false>(b, counter, sm_count, stream, reset_counter) : launch_one<CJ_LZ4_BLOCK, true>(b, counter, sm_count, stream, reset_counter);
    return serial ? launch_one<CJ_SNAPPY_RAW, false>(b, counter, sm_count, stream, reset_counter) : launch_one<CJ_SNAPPY_RAW, true>(b, counter, sm_count, stream, reset_counter);
}

}  // namespace cj
