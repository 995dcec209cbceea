/*
 * cramjam_cuda.h — C ABI of libcramjam_cuda.so, the B200 (sm_100a) batched block-codec engine that
 * stands where the reference calls `libcramjam::{snappy,lz4,zstd}::*` today.
 *
 * This header is the drop-in boundary: a host binding (the C++ `cramjam` module in this repo, or
 * upstream's Rust/pyo3 shim through `extern "C"`) calls ONLY these symbols.
 * Plain pointers and sizes, no C++/torch/Python types, never throws, never takes ownership of
 * caller memory, synchronous from the caller's view unless `where == CJ_DEVICE` (then the work is
 * enqueued on the context's stream and cj_ctx_synchronize() / the caller's own stream sync ends it).
 * Re-entrant: each cj_ctx serialises its own calls with an internal lock (the reference releases
 * the GIL around every codec call, src/lib.rs:225-288, so concurrent callers are expected).
 *
 * There is NO CPU fallback behind this interface.  If no CUDA device is usable every entry point
 * that would compute returns CJ_E_NO_DEVICE and cj_last_error() says why.
 *
 * Reference interface each entry point replaces (file:line in milesgranger/cramjam v2.12.0):
 *
 *   cj_compress_bound(CJ_SNAPPY_RAW)       snap::raw::max_compress_len          src/snappy.rs:112-115
 *   cj_compress_bound(CJ_LZ4_BLOCK)        lz4::block::compress_bound           src/lz4.rs:226-229
 *   cj_decompressed_len(CJ_SNAPPY_RAW)     snap::raw::decompress_len            src/snappy.rs:119-122
 *   cj_decompress[_batch](CJ_SNAPPY_RAW)   snappy::raw::decompress / _vec       src/snappy.rs:55-60,103-108
 *   cj_compress[_batch](CJ_SNAPPY_RAW)     snappy::raw::compress / _vec         src/snappy.rs:73-78,94-99
 *   cj_decompress[_batch](CJ_SNAPPY_FRAMED) snappy::decompress                  src/snappy.rs:22-27,87-90
 *   cj_compress[_batch](CJ_SNAPPY_FRAMED)  snappy::compress                     src/snappy.rs:37-42,81-84
 *   cj_decompress[_batch](CJ_LZ4_BLOCK)    lz4::block::decompress_into / _vec   src/lz4.rs:78-95,140-173
 *   cj_compress[_batch](CJ_LZ4_BLOCK)      lz4::block::compress_into / _vec     src/lz4.rs:113-131,191-216
 *   cj_decompress[_batch](CJ_LZ4_FRAME)    lz4::decompress                      src/lz4.rs:27-32,62-65
 *   cj_compress[_batch](CJ_LZ4_FRAME)      lz4::compress                        src/lz4.rs:42-59
 *   cj_decompress[_batch](CJ_ZSTD)         zstd::decompress                     src/zstd.rs:23-28,67-70
 *   cj_last_error()                        io::Error / snap::Error -> to_string src/exceptions.rs:9-20
 */
#ifndef CRAMJAM_CUDA_H
#define CRAMJAM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CJ_ABI_VERSION 1

typedef struct cj_ctx cj_ctx;

typedef enum {
    CJ_SNAPPY_RAW = 0,    /* snappy raw block (varint length + elements)                      */
    CJ_SNAPPY_FRAMED = 1, /* snappy framing format (stream id + CRC32C'd 64 KiB chunks)        */
    CJ_LZ4_BLOCK = 2,     /* bare LZ4 block; the 4-byte size prefix (store_size) is host framing */
    CJ_LZ4_FRAME = 3,     /* LZ4F frame                                                         */
    CJ_ZSTD = 4           /* Zstandard frame(s)                                                 */
} cj_codec;

typedef enum {
    CJ_HOST = 0,   /* every pointer is pageable host memory; staged through pinned buffers      */
    CJ_PINNED = 1, /* payload pointers are page-locked host memory (DMA'd directly); descriptor
                      arrays are ordinary host memory                                           */
    CJ_DEVICE = 2  /* every pointer (payload AND descriptor/result arrays) is device memory on the
                      context's device; nothing is copied; work is asynchronous on the stream   */
} cj_mem;

/* Return codes of the entry points (0 = the call itself succeeded; per-unit results live in
 * status[]).  Per-unit status codes share the numbering 1..10. */
enum {
    CJ_OK = 0,
    CJ_ST_EMPTY = 1,        /* empty input where the format needs at least a header             */
    CJ_ST_HEADER = 2,       /* bad magic / varint / descriptor                                   */
    CJ_ST_TRUNCATED = 3,    /* input ends inside an element                                      */
    CJ_ST_OFFSET = 4,       /* back-reference offset 0 or beyond produced output                 */
    CJ_ST_DST_SMALL = 5,    /* output capacity too small                                         */
    CJ_ST_LEN_MISMATCH = 6, /* produced length != length announced by the header                 */
    CJ_ST_CHECKSUM = 7,     /* CRC32C / XXH32 / XXH64 mismatch                                   */
    CJ_ST_CORRUPT = 8,      /* any other format violation                                        */
    CJ_ST_UNSUPPORTED = 9,  /* legal but unsupported (e.g. zstd dictionary)                      */
    CJ_ST_TOO_BIG = 10,     /* unit larger than the engine accepts (2^31-1 bytes)                */
    CJ_E_INVALID_ARG = -1,
    CJ_E_NO_DEVICE = -2,
    CJ_E_CUDA = -3,
    CJ_E_NOMEM = -4,
    CJ_E_UNIT_FAILED = -5   /* single-buffer conveniences: the unit's status was non-zero        */
};

/* A batch of n independent units in structure-of-arrays form.  Unit i reads
 * src_base[src_off[i] .. +src_len[i]) and writes dst_base[dst_off[i] .. +dst_cap[i]); the engine
 * stores the bytes produced in dst_len[i] and a CJ_ST_* code in status[i].  Units must not overlap
 * in dst.  With CJ_DEVICE, 16-byte aligned src/dst unit starts take the vectorised paths.      */
typedef struct {
    size_t n;
    const void* src_base;
    const uint64_t* src_off;
    const uint64_t* src_len;
    void* dst_base;
    const uint64_t* dst_off;
    const uint64_t* dst_cap;
    uint64_t* dst_len; /* out */
    int32_t* status;   /* out */
} cj_batch;

typedef struct {
    int32_t level;        /* lz4 frame / zstd level; <0 = codec default                          */
    int32_t acceleration; /* lz4 block `acceleration` (src/lz4.rs:113-131); <=0 = 1              */
    int32_t flags;        /* reserved, 0                                                         */
} cj_params;

/* ---- context ------------------------------------------------------------------------------ */
int cj_abi_version(void);
int cj_device_count(void);                       /* 0 when no usable CUDA device               */
int cj_ctx_create(int device, cj_ctx** out);
void cj_ctx_destroy(cj_ctx* ctx);
int cj_ctx_set_stream(cj_ctx* ctx, void* cuda_stream); /* borrow a cudaStream_t (NULL = own stream) */
int cj_ctx_synchronize(cj_ctx* ctx);
const char* cj_last_error(void);                 /* thread-local, NUL-terminated               */
const char* cj_status_string(int32_t status);    /* text used for the Python exception message */
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t cj_ctx_launch_count(const cj_ctx* ctx);
/* Tuning knob, not part of the reference surface: which kernel family decodes LZ4 / Snappy block batches of at least
 * min_units units (smaller batches always take generation 2).  2 = one warp per block; 4 = one thread per
 * block with 8-byte chunks; 7 (default) = one thread per block with 16-byte chunks and
 * granule rings.  Snappy raw and LZ4 block (LZ4 batches need 1.5 x min_units); results are identical on all of
 * them.  The default min_units is 16384. */
int cj_ctx_set_decode_path(cj_ctx* ctx, int generation, long min_units);
int cj_ctx_get_decode_path(const cj_ctx* ctx, int* generation, long* min_units);
/* Diagnostics: how many units of the most recent thread-per-block batch were handed to the generation-2 kernel (waits for the stream). */
int cj_ctx_last_redo_count(cj_ctx* ctx, unsigned* out);
/* Device-side duration in milliseconds of the codec kernels of the most recent batch call
 * (CUDA events on the context's stream; waits for them). */
int cj_ctx_last_kernel_ms(cj_ctx* ctx, float* ms);

/* ---- size helpers (pure host arithmetic / header parsing on HOST memory) ------------------- */
size_t cj_compress_bound(cj_codec codec, size_t src_len);
int cj_decompressed_len(cj_codec codec, const void* src, size_t src_len, size_t* out);
/* Upper bound of the decompressed size from headers alone (exact whenever the format records it):
 * what a host uses to size the output Vec when the caller gave no output_len. */
int cj_decompress_bound(cj_codec codec, const void* src, size_t src_len, size_t* out);

/* ---- batched core -------------------------------------------------------------------------- */
int cj_decompress_batch(cj_ctx* ctx, cj_codec codec, cj_mem where, const cj_batch* batch);
int cj_compress_batch(cj_ctx* ctx, cj_codec codec, cj_mem where, const cj_batch* batch, const cj_params* params);

/* ---- single-buffer conveniences mirroring libcramjam's slice functions (host memory) ------- */
int cj_decompress(cj_ctx* ctx, cj_codec codec, const void* src, size_t src_len, void* dst, size_t dst_cap, size_t* written);
int cj_compress(cj_ctx* ctx, cj_codec codec, const void* src, size_t src_len, void* dst, size_t dst_cap, size_t* written,
                const cj_params* params);

/* The same for a buffer pair that lives in pinned host memory (CJ_PINNED: payload DMA'd straight from / to the caller's
 * pages, no staging copy) or in device memory (CJ_DEVICE: nothing is copied; the unit descriptors are staged by the
 * engine).  This is what `Buffer`'s pinned / device staging path of the host binding calls (reference src/io.rs:370-375:
 * RustyBuffer gains a pinned-host / device backing; north_star).  Frame codecs whose container walk needs host-visible
 * bytes (snappy framing, LZ4 frame compress) reject CJ_DEVICE. */
int cj_decompress_ex(cj_ctx* ctx, cj_codec codec, cj_mem where, const void* src, size_t src_len, void* dst, size_t dst_cap,
                     size_t* written);
int cj_compress_ex(cj_ctx* ctx, cj_codec codec, cj_mem where, const void* src, size_t src_len, void* dst, size_t dst_cap,
                   size_t* written, const cj_params* params);

/* ---- synthetic "Silesia-like" corpus: identical bytes from the host and the
 *      device generator for the same (seed, first_index); used by tests and bench.py ---------- */
int cj_synth_blocks(cj_ctx* ctx, cj_mem where, void* dst, size_t n_blocks, size_t block_len, uint64_t seed,
                    uint64_t first_index);

/* ---- batched unit copy on the device (all pointers device memory; asynchronous on the stream):
 *      dst_base[dst_off[i] .. +len[i]) = src_base[src_off[i] .. +len[i]).  Packs the ragged output of
 *      cj_compress_batch into a dense arena, splices block payloads into frame containers. ------- */
int cj_copy_units(cj_ctx* ctx, size_t n, const void* src_base, const uint64_t* src_off, const uint64_t* len, void* dst_base,
                  const uint64_t* dst_off);

/* ---- device memory helpers for hosts that have no CUDA runtime binding of their own -------- */
int cj_device_alloc(cj_ctx* ctx, size_t bytes, void** out);
int cj_device_free(cj_ctx* ctx, void* p);
int cj_pinned_alloc(cj_ctx* ctx, size_t bytes, void** out);
int cj_pinned_free(cj_ctx* ctx, void* p);
/* Page-locks / releases an existing host range so that CJ_PINNED calls may DMA from / to it (the host binding's
 * Buffer(pinned=True) keeps its storage registered). */
int cj_host_register(cj_ctx* ctx, void* p, size_t bytes);
int cj_host_unregister(cj_ctx* ctx, void* p);
int cj_memcpy_h2d(cj_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes); /* async on stream */
int cj_memcpy_d2h(cj_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes); /* async on stream */

#ifdef __cplusplus
}
#endif
#endif
