// synth.cu — device side of the synthetic corpus generator (see synth.cuh).  One thread per block:
// the source model is serial within a block, and generation is set-up work outside every timed
// region, so simplicity wins over speed here.
#include "common.cuh"
#include "synth.cuh"

namespace cj {

__global__ void synth_kernel(uint8_t* __restrict__ dst, size_t n_blocks, size_t block_len, uint64_t seed, uint64_t first_index) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_blocks) synth_block(dst + i * block_len, block_len, seed, first_index + i);
}

cudaError_t launch_synth(uint8_t* dst, size_t n_blocks, size_t block_len, uint64_t seed, uint64_t first_index, cudaStream_t stream) {
    if (n_blocks == 0) return cudaSuccess;
    const int threads = 32;
    const unsigned grid = (unsigned)((n_blocks + threads - 1) / threads);
    synth_kernel<<<grid, threads, 0, stream>>>(dst, n_blocks, block_len, seed, first_index);
    return cudaGetLastError();
}

}  // namespace cj
