/*
 * cj_oracle.h — CPU restatement ("oracle") of the cramjam snappy / lz4 / zstd hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the timed CPU baseline.  The product path
 * (cramjam_b200/ + libcramjam_cuda.so) never links, imports or calls it and fails loudly when
 * its CUDA library is missing.
 *
 * Why a restatement: the reference (milesgranger/cramjam @ v2.12.0) holds no codec arithmetic.
 * Every codec body is a call into un-vendored crates —
 *     libcramjam 0.8.0  (Cargo.toml:101, Cargo.lock:407-424)
 *     snap 1.1.1        (Cargo.lock:744-746)
 *     lz4 1.28.1 / lz4-sys 1.11.1+lz4-1.10.0   (Cargo.lock:457-470)
 *     zstd 0.13.3 / zstd-sys 2.0.14+zstd.1.5.7 (Cargo.lock:1025-1050)
 * none of which is under /root/reference, and no Rust toolchain exists in the build image, so
 * oracle/_ref cannot be produced.  The functions below restate the published formats
 * (google/snappy format_description.txt + framing_format.txt, lz4_Block_format.md +
 * lz4_Frame_format.md, RFC 8878) and are anchored on the reference's call sites:
 *     src/snappy.rs:22-122   src/lz4.rs:27-229   src/zstd.rs:23-70
 *
 * PARITY PINNING (tests/test_oracle_*.py, all run without a GPU):
 *   - the reference's golden fixtures tests/data/integration/plaintext.txt.{snappy,lz4,zst}
 *     (tests/test_integration.py:32-50) decode byte-exact to plaintext.txt;
 *   - tests/test_variants.py:329-334 LZ4 block known answer b"\xe0howdy neighbor";
 *   - README.md:96-97 snappy.compress_into(15 B) == 33 bytes;
 *   - cross-decode both ways against the same C libraries the reference wraps, as present in
 *     this image: liblz4.so.1 (1.9.4), libzstd.so.1 (1.5.5), Google snappy (pyarrow), xxhash.
 */
#ifndef CJ_ORACLE_H
#define CJ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-unit status codes.  Shared numbering with include/cramjam_cuda.h (CJ_ST_*) so a parity
 * test can compare the status arrays of both implementations directly. */
enum {
    CJO_OK = 0,
    CJO_E_EMPTY = 1,          /* empty input where the format needs at least a header      */
    CJO_E_HEADER = 2,         /* bad magic / varint / descriptor                            */
    CJO_E_TRUNCATED = 3,      /* input ends inside an element                               */
    CJO_E_OFFSET = 4,         /* back-reference offset 0 or beyond produced output          */
    CJO_E_DST_SMALL = 5,      /* output capacity too small                                  */
    CJO_E_LEN_MISMATCH = 6,   /* produced length != length announced by the header          */
    CJO_E_CHECKSUM = 7,       /* CRC32C / XXH32 / XXH64 mismatch                            */
    CJO_E_CORRUPT = 8,        /* any other format violation                                 */
    CJO_E_UNSUPPORTED = 9,    /* legal but unsupported (e.g. zstd dictionary id)            */
    CJO_E_TOO_BIG = 10        /* input larger than the format allows                        */
};

/* All codec functions return bytes written (>= 0) or -(status) on failure. */

/* ---- checksums ------------------------------------------------------------------------ */
uint32_t cjo_crc32c(const void* p, size_t n);                 /* Castagnoli, reflected      */
uint32_t cjo_crc32c_masked(const void* p, size_t n);          /* snappy framing mask        */
uint32_t cjo_xxh32(const void* p, size_t n, uint32_t seed);
uint64_t cjo_xxh64(const void* p, size_t n, uint64_t seed);

/* ---- snappy raw block: src/snappy.rs:52-122 -> snap::raw ------------------------------- */
size_t  cjo_snappy_max_compressed_len(size_t n);              /* 32 + n + n/6               */
int64_t cjo_snappy_raw_decompressed_len(const uint8_t* src, size_t n);
int64_t cjo_snappy_raw_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap);
int64_t cjo_snappy_raw_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap);

/* ---- snappy framed stream: src/snappy.rs:22-42,81-90 -> snap::{read,write}::Frame* ----- */
size_t  cjo_snappy_frame_max_compressed_len(size_t n);
int64_t cjo_snappy_frame_decompressed_len(const uint8_t* src, size_t n);
int64_t cjo_snappy_frame_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap);
int64_t cjo_snappy_frame_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap);

/* ---- lz4 block: src/lz4.rs:78-229 -> lz4::block -> LZ4_compress_default/_fast,
 *      LZ4_decompress_safe.  These work on the bare block (no 4-byte size prefix); the
 *      prefix (store_size) is host-side framing handled by the callers. ------------------- */
size_t  cjo_lz4_compress_bound(size_t n);                     /* n + n/255 + 16             */
int64_t cjo_lz4_block_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int acceleration);
int64_t cjo_lz4_block_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap);

/* ---- lz4 frame: src/lz4.rs:27-65 -> lz4::{Encoder,Decoder} (LZ4F) ----------------------- */
size_t  cjo_lz4f_max_compressed_len(size_t n);
int64_t cjo_lz4f_decompressed_len(const uint8_t* src, size_t n);  /* walks blocks if no C.Size */
/* flags: bit0 = independent blocks (else linked), bit1 = content checksum, bit2 = content size */
int64_t cjo_lz4f_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, int flags);
int64_t cjo_lz4f_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap);

/* ---- zstd frame decode: src/zstd.rs:23-28,67-70 -> libzstd streaming decoder ----------- */
int64_t cjo_zstd_decompressed_len(const uint8_t* src, size_t n); /* sum of FCS over all frames; -E if any lacks it */
int64_t cjo_zstd_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap);

/* ---- batch driver used by the CPU-baseline timing legs (pthreads over independent units) */
enum { CJO_SNAPPY_RAW = 0, CJO_SNAPPY_FRAMED = 1, CJO_LZ4_BLOCK = 2, CJO_LZ4_FRAME = 3, CJO_ZSTD = 4 };
/* dir: 0 = decompress, 1 = compress.  Units are (base + off[i], len[i]).  Returns 0 and fills
 * out_len[i] (bytes or -(status)); elapsed wall seconds of the parallel region in *seconds. */
int cjo_batch(int codec, int dir, size_t n,
              const uint8_t* src_base, const uint64_t* src_off, const uint64_t* src_len,
              uint8_t* dst_base, const uint64_t* dst_off, const uint64_t* dst_cap,
              int64_t* out_len, int nthreads, double* seconds);

/* ---- the synthetic corpus of SURVEY.md 8(d) for the CPU arms (synth.c; identical bytes to the device generator) */
void cjo_synth_block(uint8_t* out, size_t len, uint64_t seed, uint64_t index);
void cjo_synth_blocks(uint8_t* out, size_t n_blocks, size_t block_len, uint64_t seed, uint64_t first_index, int nthreads);

void cjo_pack_units(const uint8_t* src, const uint64_t* so, const uint64_t* len, uint8_t* dst, const uint64_t* dof, size_t n, int nthreads);

const char* cjo_version(void);

#ifdef __cplusplus
}
#endif
#endif
