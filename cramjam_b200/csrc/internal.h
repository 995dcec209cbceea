// internal.h — definitions shared by the host-side translation units of libcramjam_cuda.so.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

void cj_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            cj_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,    \
                         cudaGetErrorString(_e));                                                   \
            return CJ_E_CUDA;                                                                       \
        }                                                                                           \
    } while (0)

namespace cj {
cudaError_t launch_lz_decode(int codec, const Batch& b, unsigned* counter, int sm_count, cudaStream_t stream, bool reset_counter = true);
cudaError_t launch_lz_encode(int codec, const Batch& b, unsigned* counter, int sm_count, int effort, cudaStream_t stream, bool reset_counter = true);
cudaError_t launch_lz4f_decode(const Batch& b, unsigned* counter, int sm_count, cudaStream_t stream);
cudaError_t launch_zstd_decode(const Batch& b, unsigned* counter, uint8_t* lit_scratch, int sm_count, cudaStream_t stream);
size_t zstd_scratch_bytes(int sm_count, uint32_t n);
cudaError_t launch_zstd_encode(const Batch& b, unsigned* counter, uint8_t* scratch, int sm_count, int level, cudaStream_t stream);
size_t zstd_enc_scratch_bytes(int sm_count, uint32_t n);
cudaError_t launch_copy_units(uint32_t n, const uint8_t* src_base, const uint64_t* src_off, const uint64_t* len, uint8_t* dst_base,
                              const uint64_t* dst_off, int sm_count, cudaStream_t stream);
cudaError_t launch_synth(uint8_t* dst, size_t n_blocks, size_t block_len, uint64_t seed, uint64_t first_index, cudaStream_t stream);
int frames_decompress(cj_ctx* ctx, int codec, int where, const cj_batch* batch);
int frames_compress(cj_ctx* ctx, int codec, int where, const cj_batch* batch, const cj_params* params);
}  // namespace cj

// Growable device / pinned scratch buffer.
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    int ensure(size_t bytes) {
        if (bytes <= cap) return CJ_OK;
        release();
        size_t want = std::max(bytes + bytes / 8, (size_t)1 << 20);
        cudaError_t e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            cap = 0;
            cj_set_error("%s of %zu bytes failed: %s", pinned ? "cudaMallocHost" : "cudaMalloc", want, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return CJ_E_NOMEM;
        }
        cap = want;
        return CJ_OK;
    }
    void release() {
        if (p) {
            if (pinned) cudaFreeHost(p);
            else cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
};

// Device scratch of the thread-per-block decode path (lz_decode4.cu): work counters + the redo list.
struct LzScratch {
    Scratch fixed_;
    int ensure_fixed(size_t bytes) { return fixed_.ensure(bytes); }
    void* fixed() const { return fixed_.p; }
};
namespace cj {
// One thread per block (lz_decode4.cu, Snappy raw and LZ4 block) + the generation-2 kernel over whatever it declines.
cudaError_t launch_lz_decode4(int codec, const Batch& b, LzScratch& sc, int sm_count, cudaStream_t stream);
// The same, generation 7 (lz_decode7.cu): 16-byte chunks, linear per-lane records.
cudaError_t launch_lz_decode7(int codec, const Batch& b, LzScratch& sc, int sm_count, cudaStream_t stream);
}

struct cj_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    unsigned* counters = nullptr;  // device work-queue counters
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    static constexpr int PIPE = 32;                // most chunks the pinned-arena pipeline may use (api.cu; CJ_PIPE_CHUNKS selects; default 32 from 1 GiB of payload, else 16; sizes ramp up, CJ_PIPE_RAMP=0 for equal chunks)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;  // copy streams of that pipeline (created on first use)
    cudaEvent_t ev_in[PIPE] = {}, ev_k[PIPE] = {};
    uint64_t launches = 0;
    int decode_gen = 7;            // LZ4/Snappy block decode path for large batches: 2 = one warp per block (lz_decode.cuh), 4 = one thread per block with
                                   // 8-byte chunks (lz_decode4.cu), 7 = one thread per block with 16-byte chunks and granule rings (lz_decode7.cu), the default
    long g4_min_units = 16384;     // smallest Snappy batch that leaves generation 2; LZ4 batches need half as many again (a lane needs ~3.4 ms / ~4.5 ms for a
                                   // 64 KiB Snappy / LZ4 block however small the batch: measured crossovers ~14 000 and ~22 000 blocks, profiles/README.md)
    const unsigned* redo_ctr = nullptr;   // device counters of the most recent generation-4 launch ([1] = units handed to generation 2)
    bool redo_valid = false;
    std::mutex mu;
    std::mutex mu_one;             // single device-resident unit calls (cj_*_ex with CJ_DEVICE)
    Scratch d_src, d_dst, d_desc, h_src, h_dst, h_desc;   // block-codec staging (run_host)
    Scratch f_dsrc, f_ddst, f_dtmp, f_ddesc, f_hsrc, f_hdst, f_hdesc;  // frame-container staging (frames.cu)
    Scratch d_one;                                                     // descriptor words of a single device-resident unit (cj_*_ex)
    Scratch z_lit, z_enc;                                              // zstd per-warp literal buffers / encoder scratch (device)
    LzScratch g4;                                                      // thread-per-block LZ decode: counters + redo list (device)
    cj_ctx() {
        h_src.pinned = h_dst.pinned = h_desc.pinned = true;
        f_hsrc.pinned = f_hdst.pinned = f_hdesc.pinned = true;
    }
    void release_all() {
        Scratch* all[] = {&d_src, &d_dst, &d_desc, &h_src, &h_dst, &h_desc, &f_dsrc, &f_ddst, &f_dtmp, &f_ddesc, &f_hsrc, &f_hdst, &f_hdesc, &z_lit, &z_enc, &g4.fixed_, &d_one};
        for (Scratch* s : all) s->release();
    }
};

// One frame of a (possibly concatenated) LZ4F / zstd stream, found by the host-side header walk.
struct cj_frame_info {
    size_t offset, size;   // position and length of the frame inside the stream
    size_t content;        // decompressed size (exact if `known`, else an upper bound); 0 for skippable frames
    bool known;
};
int cj_lz4f_walk_host(const uint8_t* s, size_t n, size_t* out, bool* exact, std::vector<cj_frame_info>* frames);
int cj_zstd_walk_host(const uint8_t* s, size_t n, size_t* out, bool* exact, std::vector<cj_frame_info>* frames);

// api.cu: run a block-codec batch whose descriptors and payload already live on the device (caller holds ctx->mu).
int cj_run_device_batch(cj_ctx* c, int codec, bool compress, const cj::Batch& b, const cj_params* params);

static inline size_t cj_align16(size_t v) { return (v + 15) & ~(size_t)15; }

// memcpy of one large range on several host threads (first-touch page faults of a fresh 256 MiB output cost ~130 ms on one
// thread; they scale with threads).
static inline void cj_parallel_copy(void* dst, const void* src, size_t bytes) {
    const size_t CH = (size_t)8 << 20;
    unsigned hw = std::thread::hardware_concurrency();
    const size_t nt = bytes < 4 * CH ? 1 : std::min<size_t>({(size_t)(hw ? hw : 1), (size_t)16, bytes / CH});
    if (nt <= 1) { if (bytes) memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([=]() {
            const size_t a = bytes * t / nt, b = bytes * (t + 1) / nt;
            memcpy((uint8_t*)dst + a, (const uint8_t*)src + a, b - a);
        });
    for (auto& t : th) t.join();
}

// Parallel loop over units on host threads (gather into / scatter out of pinned staging).
template <class F>
static void cj_parallel_units(size_t n, size_t bytes, F&& f) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = bytes < ((size_t)8 << 20) ? 1 : std::min<size_t>({(size_t)(hw ? hw : 1), (size_t)16, n});
    if (nt <= 1) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([=, &f]() {
            for (size_t i = n * t / nt, e = n * (t + 1) / nt; i < e; i++) f(i);
        });
    for (auto& t : th) t.join();
}
