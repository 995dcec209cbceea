// stubs.cu — entry points whose kernels are not built yet.  They fail loudly (no CPU fallback).
#include "internal.h"
namespace cj {
int zstd_decompress_host(cj_ctx*, int, const cj_batch*) { cj_set_error("zstd decode is not built yet"); return CJ_E_INVALID_ARG; }
}
int cj_zstd_bound_host(const uint8_t*, size_t, size_t*, bool*) { return CJ_ST_UNSUPPORTED; }
