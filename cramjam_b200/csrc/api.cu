// api.cu — the extern "C" boundary of libcramjam_cuda.so (include/cramjam_cuda.h).
//
// Host-side plumbing only: context (device, stream, scratch arenas, pinned staging), batch
// staging for CJ_HOST / CJ_PINNED callers, dispatch to the codec kernels.  No codec arithmetic
// runs on the CPU here; if the CUDA device is missing every compute entry point fails loudly.
#include <chrono>

#include "internal.h"
#include "synth.cuh"

static thread_local char g_err[512] = "";

void cj_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" {

int cj_abi_version(void) { return CJ_ABI_VERSION; }

int cj_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

const char* cj_last_error(void) { return g_err; }

const char* cj_status_string(int32_t st) {
    switch (st) {
    case CJ_OK: return "ok";
    case CJ_ST_EMPTY: return "corrupt input (empty)";
    case CJ_ST_HEADER: return "corrupt input (invalid header)";
    case CJ_ST_TRUNCATED: return "corrupt input (unexpected end of input)";
    case CJ_ST_OFFSET: return "corrupt input (back-reference offset is zero or beyond the produced output)";
    case CJ_ST_DST_SMALL: return "output buffer is too small";
    case CJ_ST_LEN_MISMATCH: return "corrupt input (decompressed length does not match the header)";
    case CJ_ST_CHECKSUM: return "corrupt input (checksum mismatch)";
    case CJ_ST_CORRUPT: return "corrupt input";
    case CJ_ST_UNSUPPORTED: return "unsupported feature in input";
    case CJ_ST_TOO_BIG: return "input is too big";
    default: return "unknown status";
    }
}

int cj_ctx_create(int device, cj_ctx** out) {
    if (!out) return CJ_E_INVALID_ARG;
    *out = nullptr;
    int n = cj_device_count();
    if (n <= 0) {
        cj_set_error("no CUDA device is available: libcramjam_cuda has no CPU fallback");
        return CJ_E_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        cj_set_error("device %d out of range (have %d)", device, n);
        return CJ_E_INVALID_ARG;
    }
    CUDA_TRY(cudaSetDevice(device));
    cj_ctx* c = new (std::nothrow) cj_ctx();
    if (!c) return CJ_E_NOMEM;
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (const char* v = getenv("CJ_DECODE_GEN")) c->decode_gen = atoi(v);
    if (c->decode_gen != 2 && c->decode_gen != 4 && c->decode_gen != 7) c->decode_gen = 7;
    if (const char* v = getenv("CJ_G4_MIN_UNITS")) c->g4_min_units = atol(v);
    // The thread-per-block decoder's far back-reference fetches are 16-byte reads scattered over 4 GiB of live windows:
    // with the default 64-byte L2 fetch granularity every miss drags a second, unused sector in from DRAM.
    if (const char* v = getenv("CJ_L2_FETCH")) {
        const int g = atoi(v);
        if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g);
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CUDA_TRY(cudaMalloc(&c->counters, 64 * sizeof(unsigned)));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    *out = c;
    return CJ_OK;
}

void cj_ctx_destroy(cj_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->release_all();
    if (c->counters) cudaFree(c->counters);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    for (int i = 0; i < cj_ctx::PIPE; i++) {
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int cj_ctx_set_stream(cj_ctx* c, void* s) {
    if (!c) return CJ_E_INVALID_ARG;
    std::lock_guard<std::mutex> g(c->mu);
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return CJ_OK;
}

int cj_ctx_synchronize(cj_ctx* c) {
    if (!c) return CJ_E_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return CJ_OK;
}

uint64_t cj_ctx_launch_count(const cj_ctx* c) { return c ? c->launches : 0; }

int cj_ctx_set_decode_path(cj_ctx* c, int generation, long min_units) {
    if (!c || (generation != 2 && generation != 4 && generation != 7) || min_units < 1) return CJ_E_INVALID_ARG;
    std::lock_guard<std::mutex> g(c->mu);
    c->decode_gen = generation;
    c->g4_min_units = min_units;
    return CJ_OK;
}

int cj_ctx_get_decode_path(const cj_ctx* c, int* generation, long* min_units) {
    if (!c) return CJ_E_INVALID_ARG;
    if (generation) *generation = c->decode_gen;
    if (min_units) *min_units = c->g4_min_units;
    return CJ_OK;
}

int cj_ctx_last_redo_count(cj_ctx* c, unsigned* out) {
    if (!c || !out) return CJ_E_INVALID_ARG;
    std::lock_guard<std::mutex> g(c->mu);
    *out = 0;
    if (!c->g4.fixed() || !c->redo_valid) return CJ_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    unsigned ctr[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpy(ctr, c->redo_ctr, sizeof ctr, cudaMemcpyDeviceToHost));
    *out = ctr[1];
    return CJ_OK;
}

int cj_ctx_last_kernel_ms(cj_ctx* c, float* ms) {
    if (!c || !ms) return CJ_E_INVALID_ARG;
    if (!c->ev_valid) {
        *ms = 0.f;
        return CJ_OK;
    }
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return CJ_OK;
}

size_t cj_compress_bound(cj_codec codec, size_t n) {
    switch (codec) {
    case CJ_SNAPPY_RAW: return 32 + n + n / 6;
    case CJ_LZ4_BLOCK: return n > 0x7E000000u ? 0 : n + n / 255 + 16;
    case CJ_SNAPPY_FRAMED: {
        size_t chunks = (n + 65535) / 65536;
        return 10 + chunks * (8 + 32 + 65536 + 65536 / 6) + 16 + 64 * 8;   // + headers of the finer chunks of small inputs (frames.cu frame_piece)
    }
    case CJ_LZ4_FRAME: return 19 + (n / 65536 + 1) * (4 + 65536 + 4) + 8 + 64 * 4;   // likewise: up to 64 blocks below 4 MiB
    // a block that does not shrink is stored raw: header + 3 bytes per block; inputs above 64 KiB are written as several
    // frames of >= 64 KiB each (frames.cu zstd_compress_split): one more frame header + block header per piece
    case CJ_ZSTD: return n + 3 * (n / (128 * 1024) + 1) + 18 + (n / 65536) * 24;
    default: return 0;
    }
}

}  // extern "C"

// ---- kernel dispatch on a device-resident batch ------------------------------------------------
static int run_device(cj_ctx* c, int codec, bool compress, const cj::Batch& b, const cj_params* params, int slot = 0, bool reset_counter = true) {
    if (b.n == 0) return CJ_OK;
    cudaError_t e;
    unsigned* counters = c->counters + (slot & 63);
    cudaEventRecord(c->ev0, c->stream);
    if (!compress) {
        if (codec == CJ_SNAPPY_RAW || codec == CJ_LZ4_BLOCK) {
            // Large batches take a thread-per-block kernel (generation 7; generation 4, DESIGN.md 4.7, on request);
            // everything else, and whatever it declines, the warp-per-block kernel (generation 2).
            // cj_ctx_set_decode_path() / CJ_DECODE_GEN select.
            const long tpb_min = codec == CJ_LZ4_BLOCK ? c->g4_min_units + c->g4_min_units / 2 : c->g4_min_units;
            if (c->decode_gen >= 4 && reset_counter && (long)b.n >= tpb_min) {
                e = c->decode_gen == 7 ? cj::launch_lz_decode7(codec, b, c->g4, c->sm_count, c->stream)
                                       : cj::launch_lz_decode4(codec, b, c->g4, c->sm_count, c->stream);
                c->redo_ctr = (const unsigned*)c->g4.fixed();   // lz_decode4.cu keeps its counters at the start of the scratch
                c->redo_valid = true;
                c->launches += 1;
            } else {
                e = cj::launch_lz_decode(codec, b, counters, c->sm_count, c->stream, reset_counter);
            }
        }
        else if (codec == CJ_LZ4_FRAME) e = cj::launch_lz4f_decode(b, counters, c->sm_count, c->stream);
        else if (codec == CJ_ZSTD) {
            int rc = c->z_lit.ensure(cj::zstd_scratch_bytes(c->sm_count, b.n));
            if (rc) return rc;
            e = cj::launch_zstd_decode(b, counters, (uint8_t*)c->z_lit.p, c->sm_count, c->stream);
        }
        else { cj_set_error("codec %d has no device-resident batch decoder", codec); return CJ_E_INVALID_ARG; }
    } else {
        // speed / ratio knob of the block encoders: lz4 `acceleration` > 1 shrinks the match table (faster, lower ratio),
        // an HC-class level (lz4 block `compression=Some(n)`, lz4 frame level >= 3) grows it.  Snappy has no knob.
        int effort = 2;
        if (codec == CJ_LZ4_BLOCK && params) {
            if (params->level >= 3) effort = 3;
            else if (params->acceleration >= 4) effort = 0;
            else if (params->acceleration >= 2) effort = 1;
        }
        if (codec == CJ_SNAPPY_RAW || codec == CJ_LZ4_BLOCK) e = cj::launch_lz_encode(codec, b, counters, c->sm_count, effort, c->stream, reset_counter);
        else if (codec == CJ_ZSTD) {
            int rc = c->z_enc.ensure(cj::zstd_enc_scratch_bytes(c->sm_count, b.n));
            if (rc) return rc;
            e = cj::launch_zstd_encode(b, counters, (uint8_t*)c->z_enc.p, c->sm_count, params ? params->level : 0, c->stream);
        }
        else { cj_set_error("codec %d has no device-resident batch encoder", codec); return CJ_E_INVALID_ARG; }
    }
    cudaEventRecord(c->ev1, c->stream);
    c->ev_valid = true;
    c->launches += 1;
    if (e != cudaSuccess) {
        cj_set_error("kernel launch failed: %s", cudaGetErrorString(e));
        return CJ_E_CUDA;
    }
    return CJ_OK;
}

// Sends the produced bytes of units [a, b) from the device arena (same offsets, rebased by d_lo) to the caller's pinned
// memory.  Units that filled their slot and whose slots touch exactly are merged into one copy, so a dense batch of
// exact-size slots is a single DMA; a gap between slots (padding, an interleaved header) or a short unit ends the run,
// so no byte of the caller's memory outside [dst_off[i], dst_off[i] + dst_len[i]) is ever written.
static int copy_runs_home(cj_ctx* c, const cj_batch* bt, const uint64_t* dl, size_t a, size_t b, uint64_t d_lo, uint8_t* hd, cudaStream_t s) {
    uint64_t run_lo = 0, run_hi = 0;
    bool open = false;
    auto flush = [&]() -> int {
        if (open && run_hi > run_lo)
            CUDA_TRY(cudaMemcpyAsync(hd + run_lo, (uint8_t*)c->d_dst.p + (run_lo - d_lo), (size_t)(run_hi - run_lo), cudaMemcpyDeviceToHost, s));
        open = false;
        return CJ_OK;
    };
    int rc;
    for (size_t i = a; i < b; i++) {
        if (!dl[i]) continue;
        if (!open || bt->dst_off[i] != run_hi) {
            if ((rc = flush())) return rc;
            run_lo = bt->dst_off[i];
            open = true;
        }
        run_hi = bt->dst_off[i] + dl[i];
        if (dl[i] != bt->dst_cap[i] && (rc = flush())) return rc;   // a short unit: the next slot does not start where this one's bytes end
    }
    return flush();
}

// Dense pinned arenas, units in ascending order on both sides: split the batch into chunks and run
// H2D(c+1) | kernel(c) | D2H(c-1) on three streams so PCIe moves in both directions while the SMs decode.
// A chunk's payload is written straight into the caller's memory only if every unit of the chunk filled
// its slot exactly; any other chunk takes the per-unit path so no byte past dst_len[i] is touched.
static int run_pinned_pipelined(cj_ctx* c, int codec, bool compress, const cj_batch* bt, const cj_params* params, uint64_t s_lo, uint64_t s_hi,
                                uint64_t d_lo, uint64_t d_hi) {
    const size_t n = bt->n;
    const uint8_t* hs = (const uint8_t*)bt->src_base;
    uint8_t* hd = (uint8_t*)bt->dst_base;
    const size_t s_bytes = (size_t)(s_hi - s_lo), d_bytes = (size_t)(d_hi - d_lo);
    int rc;
    if ((rc = c->d_src.ensure(s_bytes + 16))) return rc;
    if ((rc = c->d_dst.ensure(d_bytes + 16))) return rc;
    const size_t desc_bytes = n * (5 * sizeof(uint64_t) + sizeof(int32_t)) + 64;
    if ((rc = c->d_desc.ensure(desc_bytes))) return rc;
    if ((rc = c->h_desc.ensure(desc_bytes))) return rc;
    if (!c->s_h2d) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < cj_ctx::PIPE; i++) {
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
        }
    }
    uint64_t* hq = (uint64_t*)c->h_desc.p;
    for (size_t i = 0; i < n; i++) {
        hq[i] = bt->src_off[i] - s_lo;
        hq[2 * n + i] = bt->dst_off[i] - d_lo;
    }
    memcpy(hq + n, bt->src_len, n * 8);
    memcpy(hq + 3 * n, bt->dst_cap, n * 8);
    uint64_t* dq = (uint64_t*)c->d_desc.p;
    CUDA_TRY(cudaMemcpyAsync(dq, hq, n * 32, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->counters, 0, 64 * sizeof(unsigned), c->stream));
    // Nothing but kernels may follow on the compute stream: a copy-engine operation here (counter memset, a small
    // result copy) would queue behind the bulk transfers on that engine and stall the kernels with it.  The kernels
    // therefore write dst_len / status straight into the pinned, device-mapped descriptor buffer.
    static const int chunks_env = [] {
        const char* e = getenv("CJ_PIPE_CHUNKS");
        const int v = e ? atoi(e) : 0;
        return v < 0 ? 0 : (v > cj_ctx::PIPE ? cj_ctx::PIPE : v);
    }();
    // measured at 65 536 x 64 KiB (tools/e2e_ab.py): 16 equal chunks 92.4 ms, 16 ramped 90.2-91.1, 32 ramped 89.4 (the link alone: 86.8)
    const int chunks = chunks_env ? chunks_env : (s_bytes + d_bytes >= ((size_t)1 << 30) ? 32 : 16);
    static const bool trace_ev = getenv("CJ_TRACE") != nullptr;
    cudaEvent_t tr0 = nullptr, tr_in[cj_ctx::PIPE] = {}, tr_k0[cj_ctx::PIPE] = {}, tr_k1[cj_ctx::PIPE] = {}, tr_out[cj_ctx::PIPE] = {};
    if (trace_ev) {
        cudaEventCreate(&tr0);
        for (int i = 0; i < chunks; i++) { cudaEventCreate(&tr_in[i]); cudaEventCreate(&tr_k0[i]); cudaEventCreate(&tr_k1[i]); cudaEventCreate(&tr_out[i]); }
        cudaEventRecord(tr0, c->stream);
    }
    // Chunk sizes ramp up 1 : 2 : 4 : 8 : 16 : 16 ...: the output link is the bottleneck of a decode batch, and it idles until
    // the first chunk has been copied in and decoded, so the first chunk is small (1/207 of the batch with 16 chunks, not 1/16).
    size_t first[cj_ctx::PIPE + 1];
    {
        static const bool ramp = [] { const char* e = getenv("CJ_PIPE_RAMP"); return !e || atoi(e) != 0; }();   // 0: equal chunks (A/B)
        auto weight = [&](int k) { return (size_t)1 << (ramp && k < 4 ? k : 4); };
        size_t total = 0, acc = 0;
        for (int k = 0; k < chunks; k++) total += weight(k);
        first[0] = 0;
        for (int k = 0; k < chunks; k++) {
            acc += weight(k);
            first[k + 1] = (size_t)((unsigned __int128)n * acc / total);
        }
    }
    // enqueue all input copies and kernels
    for (int k = 0; k < chunks; k++) {
        const size_t a = first[k], b = first[k + 1];
        if (a == b) continue;
        const uint64_t lo = bt->src_off[a] - s_lo, hi = bt->src_off[b - 1] + bt->src_len[b - 1] - s_lo;
        CUDA_TRY(cudaMemcpyAsync((uint8_t*)c->d_src.p + lo, hs + s_lo + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, c->s_h2d));
        CUDA_TRY(cudaEventRecord(c->ev_in[k], c->s_h2d));
        if (trace_ev) cudaEventRecord(tr_in[k], c->s_h2d);
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_in[k], 0));
        if (trace_ev) cudaEventRecord(tr_k0[k], c->stream);
        cj::Batch bb;
        bb.n = (uint32_t)(b - a);
        bb.src_base = (const uint8_t*)c->d_src.p; bb.src_off = dq + a; bb.src_len = dq + n + a;
        bb.dst_base = (uint8_t*)c->d_dst.p; bb.dst_off = dq + 2 * n + a; bb.dst_cap = dq + 3 * n + a;
        bb.dst_len = hq + 4 * n + a; bb.status = (int32_t*)(hq + 5 * n) + a;
        if ((rc = run_device(c, codec, compress, bb, params, k, false))) return rc;
        if (trace_ev) cudaEventRecord(tr_k1[k], c->stream);
        CUDA_TRY(cudaEventRecord(c->ev_k[k], c->stream));
    }
    // drain: as each chunk's kernel finishes, send its output home
    static const bool trace = getenv("CJ_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_now = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    const uint64_t* dl = hq + 4 * n;
    for (int k = 0; k < chunks; k++) {
        const size_t a = first[k], b = first[k + 1];
        if (a == b) continue;
        CUDA_TRY(cudaEventSynchronize(c->ev_k[k]));
        if (trace) fprintf(stderr, "[cj] chunk %d kernel done at %.2f ms\n", k, ms_now());
        if ((rc = copy_runs_home(c, bt, dl, a, b, d_lo, hd, c->s_d2h))) return rc;
        if (trace_ev) cudaEventRecord(tr_out[k], c->s_d2h);
    }
    CUDA_TRY(cudaStreamSynchronize(c->s_d2h));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (trace) fprintf(stderr, "[cj] all output home at %.2f ms\n", ms_now());
    if (trace_ev) {
        for (int i = 0; i < chunks; i++) {
            float a = 0, b = 0, d = 0, o = 0;
            cudaEventElapsedTime(&a, tr0, tr_in[i]); cudaEventElapsedTime(&b, tr0, tr_k0[i]); cudaEventElapsedTime(&d, tr0, tr_k1[i]);
            cudaEventElapsedTime(&o, tr0, tr_out[i]);
            fprintf(stderr, "[cj] gpu timeline chunk %d (units %zu..%zu): h2d done %.2f  kernel start %.2f  kernel end %.2f  d2h done %.2f ms\n", i, first[i], first[i + 1], a, b, d, o);
            cudaEventDestroy(tr_in[i]); cudaEventDestroy(tr_k0[i]); cudaEventDestroy(tr_k1[i]); cudaEventDestroy(tr_out[i]);
        }
        cudaEventDestroy(tr0);
    }
    memcpy(bt->dst_len, hq + 4 * n, n * 8);
    memcpy(bt->status, hq + 5 * n, n * 4);
    return CJ_OK;
}

// Batch whose payload lives in host memory: stage -> H2D -> kernel -> D2H -> unstage.
static int run_host(cj_ctx* c, int codec, bool compress, int where, const cj_batch* bt, const cj_params* params) {
    const size_t n = bt->n;
    if (n == 0) return CJ_OK;
    if (n > 0xffffffffull) { cj_set_error("too many units"); return CJ_E_INVALID_ARG; }
    const uint8_t* hs = (const uint8_t*)bt->src_base;
    uint8_t* hd = (uint8_t*)bt->dst_base;

    // Are the units laid out as (near-)contiguous arenas already?  Then copy the whole span once
    // and keep the caller's offsets (no per-unit packing).  Always true for engine-made arenas.
    uint64_t s_lo = ~0ull, s_hi = 0, d_lo = ~0ull, d_hi = 0, s_sum = 0, d_sum = 0;
    for (size_t i = 0; i < n; i++) {
        s_lo = std::min(s_lo, bt->src_off[i]); s_hi = std::max(s_hi, bt->src_off[i] + bt->src_len[i]); s_sum += bt->src_len[i];
        d_lo = std::min(d_lo, bt->dst_off[i]); d_hi = std::max(d_hi, bt->dst_off[i] + bt->dst_cap[i]); d_sum += bt->dst_cap[i];
    }
    const bool direct = where == CJ_PINNED;
    const bool s_span = direct && (s_hi - s_lo) <= s_sum + s_sum / 4 + 64 * n;
    const bool d_span = direct && (d_hi - d_lo) <= d_sum + 16 * n;
    if (s_span && d_span && n >= 256 && s_sum + d_sum >= ((uint64_t)64 << 20) && (codec == CJ_SNAPPY_RAW || codec == CJ_LZ4_BLOCK)) {
        bool ascending = true;
        for (size_t i = 1; i < n && ascending; i++)
            ascending = bt->src_off[i] >= bt->src_off[i - 1] + bt->src_len[i - 1] && bt->dst_off[i] >= bt->dst_off[i - 1] + bt->dst_cap[i - 1];
        if (ascending) return run_pinned_pipelined(c, codec, compress, bt, params, s_lo, s_hi, d_lo, d_hi);
    }

    // device-side layout
    std::vector<uint64_t> so(n), doff(n);
    size_t s_bytes, d_bytes;
    if (s_span) {
        for (size_t i = 0; i < n; i++) so[i] = bt->src_off[i] - s_lo;
        s_bytes = (size_t)(s_hi - s_lo);
    } else {
        size_t acc = 0;
        for (size_t i = 0; i < n; i++) { so[i] = acc; acc += cj_align16((size_t)bt->src_len[i]); }
        s_bytes = acc;
    }
    if (d_span) {
        for (size_t i = 0; i < n; i++) doff[i] = bt->dst_off[i] - d_lo;
        d_bytes = (size_t)(d_hi - d_lo);
    } else {
        size_t acc = 0;
        for (size_t i = 0; i < n; i++) { doff[i] = acc; acc += cj_align16((size_t)bt->dst_cap[i]); }
        d_bytes = acc;
    }
    int rc;
    if ((rc = c->d_src.ensure(s_bytes + 16))) return rc;
    if ((rc = c->d_dst.ensure(d_bytes + 16))) return rc;
    const size_t desc_bytes = n * (5 * sizeof(uint64_t) + sizeof(int32_t)) + 64;
    if ((rc = c->d_desc.ensure(desc_bytes))) return rc;
    if ((rc = c->h_desc.ensure(desc_bytes))) return rc;

    // descriptors: [src_off | src_len | dst_off | dst_cap | dst_len(out) | status(out)]
    uint64_t* hq = (uint64_t*)c->h_desc.p;
    memcpy(hq, so.data(), n * 8);
    memcpy(hq + n, bt->src_len, n * 8);
    memcpy(hq + 2 * n, doff.data(), n * 8);
    memcpy(hq + 3 * n, bt->dst_cap, n * 8);
    uint64_t* dq = (uint64_t*)c->d_desc.p;
    CUDA_TRY(cudaMemcpyAsync(dq, hq, n * 32, cudaMemcpyHostToDevice, c->stream));

    // payload in
    if (s_span) {
        CUDA_TRY(cudaMemcpyAsync(c->d_src.p, hs + s_lo, s_bytes, cudaMemcpyHostToDevice, c->stream));
    } else {
        if ((rc = c->h_src.ensure(s_bytes + 16))) return rc;
        uint8_t* stage = (uint8_t*)c->h_src.p;
        if (n <= 4) { for (size_t i = 0; i < n; i++) cj_parallel_copy(stage + so[i], hs + bt->src_off[i], (size_t)bt->src_len[i]); }
        else cj_parallel_units(n, s_bytes, [&](size_t i) { memcpy(stage + so[i], hs + bt->src_off[i], (size_t)bt->src_len[i]); });
        CUDA_TRY(cudaMemcpyAsync(c->d_src.p, stage, s_bytes, cudaMemcpyHostToDevice, c->stream));
    }

    cj::Batch b;
    b.n = (uint32_t)n;
    b.src_base = (const uint8_t*)c->d_src.p;
    b.src_off = dq;
    b.src_len = dq + n;
    b.dst_base = (uint8_t*)c->d_dst.p;
    b.dst_off = dq + 2 * n;
    b.dst_cap = dq + 3 * n;
    b.dst_len = dq + 4 * n;
    b.status = (int32_t*)(dq + 5 * n);
    if ((rc = run_device(c, codec, compress, b, params))) return rc;

    // results
    CUDA_TRY(cudaMemcpyAsync(hq + 4 * n, dq + 4 * n, n * 8 + n * 4, cudaMemcpyDeviceToHost, c->stream));
    bool tight = false;
    if (d_span) {
        // decompress into exact-size slots: the span is exactly the produced bytes when every unit
        // filled its capacity; otherwise fall through to the per-unit path so that no byte past
        // dst_len[i] of the caller's memory is touched.
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        // (ascending, exactly touching slots only: a gap between slots belongs to the caller and stays untouched)
        tight = true;
        const uint64_t* dl = hq + 4 * n;
        for (size_t i = 0; i < n && tight; i++)
            tight = dl[i] == bt->dst_cap[i] && (i == 0 || bt->dst_off[i] == bt->dst_off[i - 1] + bt->dst_cap[i - 1]);
        if (tight) CUDA_TRY(cudaMemcpyAsync(hd + d_lo, c->d_dst.p, d_bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    if (!tight) {
        if ((rc = c->h_dst.ensure(d_bytes + 16))) return rc;
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const uint64_t* dl = hq + 4 * n;
        // copy back only the prefix of the arena that holds produced bytes
        size_t last_end = 0;
        for (size_t i = 0; i < n; i++) last_end = std::max(last_end, (size_t)(doff[i] + dl[i]));
        uint8_t* stage = (uint8_t*)c->h_dst.p;
        if (last_end) CUDA_TRY(cudaMemcpyAsync(stage, c->d_dst.p, last_end, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (n <= 4) { for (size_t i = 0; i < n; i++) if (dl[i]) cj_parallel_copy(hd + bt->dst_off[i], stage + doff[i], (size_t)dl[i]); }
        else cj_parallel_units(n, last_end, [&](size_t i) { if (dl[i]) memcpy(hd + bt->dst_off[i], stage + doff[i], (size_t)dl[i]); });
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(bt->dst_len, hq + 4 * n, n * 8);
    memcpy(bt->status, hq + 5 * n, n * 4);
    return CJ_OK;
}

static int run_batch(cj_ctx* c, int codec, bool compress, int where, const cj_batch* bt, const cj_params* params) {
    if (!c || !bt) { cj_set_error("null context or batch"); return CJ_E_INVALID_ARG; }
    if (bt->n && (!bt->src_off || !bt->src_len || !bt->dst_off || !bt->dst_cap || !bt->dst_len || !bt->status)) {
        cj_set_error("null descriptor array");
        return CJ_E_INVALID_ARG;
    }
    std::lock_guard<std::mutex> g(c->mu);
    CUDA_TRY(cudaSetDevice(c->device));
    if (codec == CJ_SNAPPY_FRAMED || (codec == CJ_LZ4_FRAME && (compress || where != CJ_DEVICE)))
        return compress ? cj::frames_compress(c, codec, where, bt, params) : cj::frames_decompress(c, codec, where, bt);
    // Zstandard with host-visible buffers: large inputs are compressed as several frames, multi-frame streams are decoded one
    // frame per warp (frames.cu); batches of small single-frame units keep the plain one-warp-per-unit plumbing below
    if (codec == CJ_ZSTD && where != CJ_DEVICE && bt->n) {
        bool big = false;
        for (size_t i = 0; i < bt->n && !big; i++) big = bt->src_len[i] > (compress ? (64u << 10) : (128u << 10));
        if (big) return compress ? cj::frames_compress(c, codec, where, bt, params) : cj::frames_decompress(c, codec, where, bt);
    }
    // a raw Snappy block larger than 128 KiB is compressed piecewise, one warp per 64 KiB (frames.cu)
    if (codec == CJ_SNAPPY_RAW && compress && where != CJ_DEVICE && bt->n) {
        bool big = false;
        for (size_t i = 0; i < bt->n && !big; i++) big = bt->src_len[i] > (128u << 10) && bt->src_len[i] <= 0xFFFFFFFFull;
        if (big) return cj::frames_compress(c, codec, where, bt, params);
    }
    // LZ4 frame decode is a per-unit kernel (one warp walks a frame's blocks), so it shares the block-codec plumbing
    if (codec != CJ_SNAPPY_RAW && codec != CJ_LZ4_BLOCK && codec != CJ_LZ4_FRAME && codec != CJ_ZSTD) { cj_set_error("unknown codec %d", codec); return CJ_E_INVALID_ARG; }
    if (where == CJ_DEVICE) {
        if (bt->n > 0xffffffffull) { cj_set_error("too many units"); return CJ_E_INVALID_ARG; }
        cj::Batch b;
        b.n = (uint32_t)bt->n;
        b.src_base = (const uint8_t*)bt->src_base; b.src_off = bt->src_off; b.src_len = bt->src_len;
        b.dst_base = (uint8_t*)bt->dst_base; b.dst_off = bt->dst_off; b.dst_cap = bt->dst_cap;
        b.dst_len = bt->dst_len; b.status = bt->status;
        return run_device(c, codec, compress, b, params);
    }
    return run_host(c, codec, compress, where, bt, params);
}

// Entry used by frames.cu for the block payloads it has already placed in device memory.
int cj_run_device_batch(cj_ctx* c, int codec, bool compress, const cj::Batch& b, const cj_params* params) {
    return run_device(c, codec, compress, b, params);
}

extern "C" {

int cj_decompress_batch(cj_ctx* c, cj_codec codec, cj_mem where, const cj_batch* batch) {
    return run_batch(c, (int)codec, false, (int)where, batch, nullptr);
}

int cj_compress_batch(cj_ctx* c, cj_codec codec, cj_mem where, const cj_batch* batch, const cj_params* params) {
    return run_batch(c, (int)codec, true, (int)where, batch, params);
}

static int single(cj_ctx* c, cj_codec codec, bool compress, const void* src, size_t n, void* dst, size_t cap, size_t* written,
                  const cj_params* params, int where = CJ_HOST) {
    uint64_t so = 0, sl = n, dof = 0, dc = cap, dl = 0;
    int32_t st = 0;
    static uint8_t dummy[16];
    cj_batch b;
    b.n = 1;
    b.src_base = src ? src : dummy; b.src_off = &so; b.src_len = &sl;
    b.dst_base = dst ? dst : dummy; b.dst_off = &dof; b.dst_cap = &dc;
    b.dst_len = &dl; b.status = &st;
    int rc;
    if (where == CJ_DEVICE) {
        // device-resident pair: the six descriptor words travel through a small device scratch of their own
        if (!c) { cj_set_error("null context"); return CJ_E_INVALID_ARG; }
        if (!src || !dst) { cj_set_error("null device pointer"); return CJ_E_INVALID_ARG; }
        uint64_t h[6] = {so, sl, dof, dc, 0, 0};
        std::lock_guard<std::mutex> one(c->mu_one);   // the scratch is this call's from upload to read-back
        {
            std::lock_guard<std::mutex> g(c->mu);
            CUDA_TRY(cudaSetDevice(c->device));
            if ((rc = c->d_one.ensure(64))) return rc;
            CUDA_TRY(cudaMemcpyAsync(c->d_one.p, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
        }
        uint64_t* d = (uint64_t*)c->d_one.p;
        b.src_off = d; b.src_len = d + 1; b.dst_off = d + 2; b.dst_cap = d + 3; b.dst_len = d + 4; b.status = (int32_t*)(d + 5);
        rc = run_batch(c, (int)codec, compress, CJ_DEVICE, &b, params);
        if (rc) return rc;
        std::lock_guard<std::mutex> g(c->mu);
        CUDA_TRY(cudaMemcpyAsync(h + 4, d + 4, 16, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        dl = h[4];
        st = (int32_t)(h[5] & 0xFFFFFFFFu);
    } else {
        rc = run_batch(c, (int)codec, compress, where, &b, params);
        if (rc) return rc;
    }
    if (written) *written = (size_t)dl;
    if (st != CJ_OK) {
        cj_set_error("%s", cj_status_string(st));
        return CJ_E_UNIT_FAILED;
    }
    return CJ_OK;
}

int cj_decompress(cj_ctx* c, cj_codec codec, const void* src, size_t n, void* dst, size_t cap, size_t* written) {
    return single(c, codec, false, src, n, dst, cap, written, nullptr);
}

int cj_compress(cj_ctx* c, cj_codec codec, const void* src, size_t n, void* dst, size_t cap, size_t* written, const cj_params* params) {
    return single(c, codec, true, src, n, dst, cap, written, params);
}

int cj_decompress_ex(cj_ctx* c, cj_codec codec, cj_mem where, const void* src, size_t n, void* dst, size_t cap, size_t* written) {
    return single(c, codec, false, src, n, dst, cap, written, nullptr, (int)where);
}

int cj_compress_ex(cj_ctx* c, cj_codec codec, cj_mem where, const void* src, size_t n, void* dst, size_t cap, size_t* written,
                   const cj_params* params) {
    return single(c, codec, true, src, n, dst, cap, written, params, (int)where);
}

int cj_synth_blocks(cj_ctx* c, cj_mem where, void* dst, size_t n_blocks, size_t block_len, uint64_t seed, uint64_t first_index) {
    if (!dst && n_blocks != 0 && block_len != 0) return CJ_E_INVALID_ARG;
    if (where == CJ_DEVICE) {
        if (!c) return CJ_E_INVALID_ARG;
        std::lock_guard<std::mutex> g(c->mu);
        CUDA_TRY(cudaSetDevice(c->device));
        cudaError_t e = cj::launch_synth((uint8_t*)dst, n_blocks, block_len, seed, first_index, c->stream);
        c->launches += 1;
        if (e != cudaSuccess) { cj_set_error("synth launch failed: %s", cudaGetErrorString(e)); return CJ_E_CUDA; }
        return CJ_OK;
    }
    // host generator: same function, used for CPU baselines and parity fixtures (no ctx needed)
    uint8_t* out = (uint8_t*)dst;
    cj_parallel_units(n_blocks, n_blocks * block_len, [&](size_t i) { cj::synth_block(out + i * block_len, block_len, seed, first_index + i); });
    return CJ_OK;
}

int cj_copy_units(cj_ctx* c, size_t n, const void* src_base, const uint64_t* src_off, const uint64_t* len, void* dst_base,
                  const uint64_t* dst_off) {
    if (!c || n > 0xffffffffull) return CJ_E_INVALID_ARG;
    std::lock_guard<std::mutex> g(c->mu);
    CUDA_TRY(cudaSetDevice(c->device));
    cudaError_t e = cj::launch_copy_units((uint32_t)n, (const uint8_t*)src_base, src_off, len, (uint8_t*)dst_base, dst_off, c->sm_count, c->stream);
    c->launches += 1;
    if (e != cudaSuccess) { cj_set_error("copy_units launch failed: %s", cudaGetErrorString(e)); return CJ_E_CUDA; }
    return CJ_OK;
}

int cj_device_alloc(cj_ctx* c, size_t bytes, void** out) {
    if (!c || !out) return CJ_E_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
    if (e != cudaSuccess) { (void)cudaGetLastError(); cj_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return CJ_E_NOMEM; }
    return CJ_OK;
}
int cj_device_free(cj_ctx* c, void* p) {
    if (!c) return CJ_E_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaFree(p));
    return CJ_OK;
}
int cj_pinned_alloc(cj_ctx* c, size_t bytes, void** out) {
    if (!c || !out) return CJ_E_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 16);
    if (e != cudaSuccess) { (void)cudaGetLastError(); cj_set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); return CJ_E_NOMEM; }
    return CJ_OK;
}
int cj_pinned_free(cj_ctx* c, void* p) {
    if (!c) return CJ_E_INVALID_ARG;
    CUDA_TRY(cudaFreeHost(p));
    return CJ_OK;
}
int cj_host_register(cj_ctx* c, void* p, size_t bytes) {
    if (!c || !p || !bytes) return CJ_E_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { (void)cudaGetLastError(); cj_set_error("cudaHostRegister(%zu) failed: %s", bytes, cudaGetErrorString(e)); return CJ_E_CUDA; }
    return CJ_OK;
}
int cj_host_unregister(cj_ctx* c, void* p) {
    if (!c || !p) return CJ_E_INVALID_ARG;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); cj_set_error("cudaHostUnregister failed: %s", cudaGetErrorString(e)); return CJ_E_CUDA; }
    return CJ_OK;
}
int cj_memcpy_h2d(cj_ctx* c, void* d, const void* s, size_t bytes) {
    if (!c) return CJ_E_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, c->stream));
    return CJ_OK;
}
int cj_memcpy_d2h(cj_ctx* c, void* d, const void* s, size_t bytes) {
    if (!c) return CJ_E_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, c->stream));
    return CJ_OK;
}

}  // extern "C"
