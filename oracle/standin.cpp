// oracle/standin.cpp — the fastest CPU implementations of the path that exist in this image, behind the batch interface of
// batch.c.  TEST / BASELINE INFRASTRUCTURE ONLY (see cj_oracle.h): bench.py's cpu_baseline leg and `--impl reference` time
// it; the product never loads it.
//
// The reference's arithmetic lives in Rust crates that cannot be built here (snap 1.1.1, lz4-sys 1.11.1+lz4-1.10.0,
// zstd-sys 2.0.14+zstd.1.5.7; Cargo.lock:457-470,744-746,1025-1050).  The closest compiled stand-ins in the image are
// the C/C++ libraries Apache Arrow bundles in pyarrow's libarrow.so — Google snappy (the library snap is a port of),
// lz4 and zstd — reached through arrow::util::Codec (arrow/util/compression.h), one Codec per worker thread, output
// written into caller-provided buffers (no allocation inside the timed region), units handed out by an atomic
// counter exactly as batch.c does for the oracle port.
#include <arrow/util/compression.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

namespace {
arrow::Compression::type kind(int codec) {
    switch (codec) {
    case 0: return arrow::Compression::SNAPPY;     // raw snappy block (CJO_SNAPPY_RAW)
    case 2: return arrow::Compression::LZ4;        // raw LZ4 block, no size prefix (CJO_LZ4_BLOCK)
    case 3: return arrow::Compression::LZ4_FRAME;  // CJO_LZ4_FRAME
    case 4: return arrow::Compression::ZSTD;       // CJO_ZSTD
    default: return arrow::Compression::UNCOMPRESSED;
    }
}
}  // namespace

extern "C" {

const char* cjs_describe(void) { return "Apache Arrow bundled codecs (Google snappy / lz4 / zstd) via arrow::util::Codec, pyarrow libarrow.so"; }

// dir: 0 = decompress, 1 = compress (level: zstd level, ignored otherwise).  Returns 0, fills out_len[i] (bytes, or -1 on
// failure) and the wall seconds of the parallel region; -1 if the codec is not available.
int cjs_batch(int codec, int dir, int level, size_t n, const uint8_t* src_base, const uint64_t* src_off, const uint64_t* src_len,
              uint8_t* dst_base, const uint64_t* dst_off, const uint64_t* dst_cap, int64_t* out_len, int nthreads, double* seconds) {
    const arrow::Compression::type k = kind(codec);
    if (k == arrow::Compression::UNCOMPRESSED || !arrow::util::Codec::IsAvailable(k)) return -1;
    if (nthreads < 1) nthreads = 1;
    std::vector<std::unique_ptr<arrow::util::Codec>> codecs;
    for (int t = 0; t < nthreads; t++) {
        auto r = (codec == 4 && dir == 1) ? arrow::util::Codec::Create(k, level) : arrow::util::Codec::Create(k);
        if (!r.ok()) return -1;
        codecs.push_back(std::move(r).ValueOrDie());
    }
    std::atomic<size_t> next{0};
    auto work = [&](int t) {
        arrow::util::Codec* c = codecs[t].get();
        for (;;) {
            size_t i = next.fetch_add(8, std::memory_order_relaxed);
            if (i >= n) break;
            const size_t e = i + 8 < n ? i + 8 : n;
            for (; i < e; i++) {
                auto r = dir == 0 ? c->Decompress((int64_t)src_len[i], src_base + src_off[i], (int64_t)dst_cap[i], dst_base + dst_off[i])
                                  : c->Compress((int64_t)src_len[i], src_base + src_off[i], (int64_t)dst_cap[i], dst_base + dst_off[i]);
                out_len[i] = r.ok() ? *r : -1;
            }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

}  // extern "C"
