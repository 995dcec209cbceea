// stubs.cu — entry points whose kernels are not built yet.  They fail loudly (no CPU fallback).
#include "common.cuh"
void cj_set_error(const char* fmt, ...);
namespace cj {
int frames_decompress(cj_ctx*, int codec, int, const cj_batch*) { cj_set_error("codec %d decode not built yet", codec); return CJ_E_INVALID_ARG; }
int frames_compress(cj_ctx*, int codec, int, const cj_batch*, const cj_params*) { cj_set_error("codec %d encode not built yet", codec); return CJ_E_INVALID_ARG; }
}
extern "C" int cj_decompressed_len(cj_codec, const void*, size_t, size_t*) { return CJ_E_INVALID_ARG; }
