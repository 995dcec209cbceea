"""ctypes access to the system codec libraries used as independent cross-checks of the oracle:
liblz4.so.1, libzstd.so.1 (the C libraries the reference wraps through lz4-sys / zstd-sys, at
the versions present in this image) and Google snappy via pyarrow.  Test infrastructure only."""
import ctypes as C

import numpy as np


def _load(name):
    try:
        return C.CDLL(name)
    except OSError:
        return None


_lz4 = _load("liblz4.so.1")
_zstd = _load("libzstd.so.1")

if _lz4 is not None:
    _lz4.LZ4_compress_default.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    _lz4.LZ4_compress_fast.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    _lz4.LZ4_compress_HC.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    _lz4.LZ4_decompress_safe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    _lz4.LZ4_compressBound.argtypes = [C.c_int]
    _lz4.LZ4F_compressFrameBound.argtypes = [C.c_size_t, C.c_void_p]
    _lz4.LZ4F_compressFrameBound.restype = C.c_size_t
    _lz4.LZ4F_compressFrame.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p]
    _lz4.LZ4F_compressFrame.restype = C.c_size_t
    _lz4.LZ4F_isError.argtypes = [C.c_size_t]

have_lz4 = _lz4 is not None
have_zstd = _zstd is not None


def lz4_compress(b, accel=1, hc=None):
    cap = _lz4.LZ4_compressBound(len(b))
    out = C.create_string_buffer(max(cap, 1))
    if hc is not None:
        n = _lz4.LZ4_compress_HC(b, out, len(b), cap, hc)
    else:
        n = _lz4.LZ4_compress_fast(b, out, len(b), cap, accel)
    assert n > 0 or len(b) == 0
    return out.raw[:n]


def lz4_decompress(b, cap):
    """Returns bytes or None on error (LZ4_decompress_safe < 0)."""
    out = C.create_string_buffer(max(cap, 1))
    n = _lz4.LZ4_decompress_safe(b, out, len(b), cap)
    return None if n < 0 else out.raw[:n]


class _LZ4F_frameInfo(C.Structure):
    _fields_ = [("blockSizeID", C.c_int), ("blockMode", C.c_int), ("contentChecksumFlag", C.c_int),
                ("frameType", C.c_int), ("contentSize", C.c_ulonglong), ("dictID", C.c_uint),
                ("blockChecksumFlag", C.c_int)]


class _LZ4F_prefs(C.Structure):
    _fields_ = [("frameInfo", _LZ4F_frameInfo), ("compressionLevel", C.c_int), ("autoFlush", C.c_uint),
                ("favorDecSpeed", C.c_uint), ("reserved", C.c_uint * 3)]


def lz4f_compress(b, level=0, independent=False, content_checksum=True, block_checksum=False,
                  block_size_id=4, content_size=False):
    p = _LZ4F_prefs()
    p.frameInfo.blockSizeID = block_size_id
    p.frameInfo.blockMode = 1 if independent else 0
    p.frameInfo.contentChecksumFlag = 1 if content_checksum else 0
    p.frameInfo.blockChecksumFlag = 1 if block_checksum else 0
    p.frameInfo.contentSize = len(b) if content_size else 0
    p.compressionLevel = level
    cap = _lz4.LZ4F_compressFrameBound(len(b), C.byref(p))
    out = C.create_string_buffer(cap)
    n = _lz4.LZ4F_compressFrame(out, cap, b, len(b), C.byref(p))
    assert not _lz4.LZ4F_isError(n)
    return out.raw[:n]


if _zstd is not None:
    _zstd.ZSTD_compressBound.argtypes = [C.c_size_t]
    _zstd.ZSTD_compressBound.restype = C.c_size_t
    _zstd.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    _zstd.ZSTD_compress.restype = C.c_size_t
    _zstd.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    _zstd.ZSTD_decompress.restype = C.c_size_t
    _zstd.ZSTD_isError.argtypes = [C.c_size_t]
    _zstd.ZSTD_createCCtx.restype = C.c_void_p
    _zstd.ZSTD_freeCCtx.argtypes = [C.c_void_p]
    _zstd.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
    _zstd.ZSTD_CCtx_setParameter.restype = C.c_size_t
    _zstd.ZSTD_compress2.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    _zstd.ZSTD_compress2.restype = C.c_size_t


def zstd_compress(b, level=3, checksum=False, window_log=None, content_size=True):
    cap = _zstd.ZSTD_compressBound(len(b))
    out = C.create_string_buffer(max(cap, 64))
    cctx = _zstd.ZSTD_createCCtx()
    try:
        _zstd.ZSTD_CCtx_setParameter(cctx, 100, level)       # ZSTD_c_compressionLevel
        _zstd.ZSTD_CCtx_setParameter(cctx, 201, 1 if checksum else 0)  # ZSTD_c_checksumFlag
        _zstd.ZSTD_CCtx_setParameter(cctx, 200, 1 if content_size else 0)  # ZSTD_c_contentSizeFlag
        if window_log is not None:
            _zstd.ZSTD_CCtx_setParameter(cctx, 101, window_log)  # ZSTD_c_windowLog
        n = _zstd.ZSTD_compress2(cctx, out, len(out), b, len(b))
    finally:
        _zstd.ZSTD_freeCCtx(cctx)
    assert not _zstd.ZSTD_isError(n)
    return out.raw[:n]


def zstd_decompress(b, cap):
    out = C.create_string_buffer(max(cap, 1))
    n = _zstd.ZSTD_decompress(out, cap, b, len(b))
    return None if _zstd.ZSTD_isError(n) else out.raw[:n]


def snappy_compress(b):
    import pyarrow as pa
    return pa.compress(b, codec="snappy", asbytes=True)


def snappy_decompress(b, n):
    import pyarrow as pa
    return pa.decompress(b, decompressed_size=n, codec="snappy", asbytes=True)
