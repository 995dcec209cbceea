"""ctypes loader for the CPU oracle (oracle/libcj_oracle.so); `import oracle` from the repo root.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libcj_oracle.so")

SNAPPY_RAW, SNAPPY_FRAMED, LZ4_BLOCK, LZ4_FRAME, ZSTD = range(5)

STATUS = {
    0: "OK", 1: "EMPTY", 2: "HEADER", 3: "TRUNCATED", 4: "OFFSET", 5: "DST_SMALL",
    6: "LEN_MISMATCH", 7: "CHECKSUM", 8: "CORRUPT", 9: "UNSUPPORTED", 10: "TOO_BIG",
}


def build(force=False):
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _DIR, "-s"])
    return _SO


_STANDIN_SO = os.path.join(_DIR, "libcj_standin.so")
_standin = None


def build_standin(force=False):
    """Builds the stand-in harness (standin.cpp over pyarrow's bundled Google snappy / lz4 / zstd).  Returns the path,
    or None when pyarrow (its headers or libarrow) is not there."""
    src = os.path.join(_DIR, "standin.cpp")
    if not force and os.path.exists(_STANDIN_SO) and os.path.getmtime(_STANDIN_SO) >= os.path.getmtime(src):
        return _STANDIN_SO
    try:
        import glob
        import pyarrow
        inc = pyarrow.get_include()
        libs = sorted(glob.glob(os.path.join(os.path.dirname(pyarrow.__file__), "libarrow.so*")))
        if not libs or not os.path.exists(os.path.join(inc, "arrow", "util", "compression.h")):
            return None
        subprocess.check_call(["make", "-C", _DIR, "-s", "libcj_standin.so", f"ARROW_INC={inc}", f"ARROW_LIB={libs[0]}"])
        return _STANDIN_SO
    except Exception:
        return None


def standin():
    """The stand-in library or None.  Only bench.py's CPU legs and tests use it."""
    global _standin
    if _standin is None:
        so = build_standin()
        if so is None:
            _standin = False
        else:
            try:
                L = C.CDLL(so)
                u8p, sz = C.c_void_p, C.c_size_t
                L.cjs_batch.argtypes = [C.c_int, C.c_int, C.c_int, sz, u8p, u8p, u8p, u8p, u8p, u8p, u8p, C.c_int, C.POINTER(C.c_double)]
                L.cjs_batch.restype = C.c_int
                L.cjs_describe.restype = C.c_char_p
                _standin = L
            except OSError:
                _standin = False
    return _standin or None


def standin_batch(codec, direction, src_base, src_off, src_len, dst_base, dst_off, dst_cap, nthreads=1, level=3):
    """Same contract as batch() over the stand-in libraries; out_len[i] = -1 where a unit failed.  None if unavailable."""
    L = standin()
    if L is None:
        return None
    n = len(src_off)
    so = np.ascontiguousarray(src_off, dtype=np.uint64); sl = np.ascontiguousarray(src_len, dtype=np.uint64)
    do = np.ascontiguousarray(dst_off, dtype=np.uint64); dc = np.ascontiguousarray(dst_cap, dtype=np.uint64)
    out = np.empty(n, dtype=np.int64)
    sec = C.c_double(0)
    rc = L.cjs_batch(codec, direction, level, n, src_base.ctypes.data, so.ctypes.data, sl.ctypes.data,
                     dst_base.ctypes.data, do.ctypes.data, dc.ctypes.data, out.ctypes.data, nthreads, C.byref(sec))
    if rc != 0:
        return None
    return out, sec.value


def pack_units(src, src_off, lens, dst, dst_off, nthreads=1):
    """dst[dst_off[i] : +lens[i]] = src[src_off[i] : +lens[i]] for every unit (benchmark set-up helper)."""
    so = np.ascontiguousarray(src_off, dtype=np.uint64); ln = np.ascontiguousarray(lens, dtype=np.uint64)
    do = np.ascontiguousarray(dst_off, dtype=np.uint64)
    lib().cjo_pack_units(src.ctypes.data, so.ctypes.data, ln.ctypes.data, dst.ctypes.data, do.ctypes.data, len(ln), nthreads)


def synth(n_blocks, block_len, seed=0xC0FFEE, first_index=0, nthreads=None):
    """The synthetic corpus (oracle/synth.c), identical bytes to cramjam_b200's device / host generator."""
    out = np.empty(n_blocks * block_len, dtype=np.uint8)
    lib().cjo_synth_blocks(out.ctypes.data, n_blocks, block_len, seed, first_index, nthreads or (os.cpu_count() or 1))
    return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, sz, i64 = C.c_void_p, C.c_size_t, C.c_int64
        for name in ("cjo_snappy_raw_compress", "cjo_snappy_raw_decompress", "cjo_snappy_frame_compress",
                     "cjo_snappy_frame_decompress", "cjo_lz4_block_decompress", "cjo_lz4f_decompress",
                     "cjo_zstd_decompress"):
            f = getattr(L, name)
            f.argtypes = [u8p, sz, u8p, sz]
            f.restype = i64
        L.cjo_lz4_block_compress.argtypes = [u8p, sz, u8p, sz, C.c_int]
        L.cjo_lz4_block_compress.restype = i64
        L.cjo_lz4f_compress.argtypes = [u8p, sz, u8p, sz, C.c_int]
        L.cjo_lz4f_compress.restype = i64
        for name in ("cjo_snappy_raw_decompressed_len", "cjo_snappy_frame_decompressed_len",
                     "cjo_lz4f_decompressed_len", "cjo_zstd_decompressed_len"):
            f = getattr(L, name)
            f.argtypes = [u8p, sz]
            f.restype = i64
        for name in ("cjo_snappy_max_compressed_len", "cjo_snappy_frame_max_compressed_len",
                     "cjo_lz4_compress_bound", "cjo_lz4f_max_compressed_len"):
            f = getattr(L, name)
            f.argtypes = [sz]
            f.restype = sz
        L.cjo_crc32c.argtypes = [u8p, sz]; L.cjo_crc32c.restype = C.c_uint32
        L.cjo_crc32c_masked.argtypes = [u8p, sz]; L.cjo_crc32c_masked.restype = C.c_uint32
        L.cjo_xxh32.argtypes = [u8p, sz, C.c_uint32]; L.cjo_xxh32.restype = C.c_uint32
        L.cjo_xxh64.argtypes = [u8p, sz, C.c_uint64]; L.cjo_xxh64.restype = C.c_uint64
        L.cjo_batch.argtypes = [C.c_int, C.c_int, sz, u8p, u8p, u8p, u8p, u8p, u8p, u8p, C.c_int, C.POINTER(C.c_double)]
        L.cjo_batch.restype = C.c_int
        L.cjo_synth_blocks.argtypes = [u8p, sz, sz, C.c_uint64, C.c_uint64, C.c_int]
        L.cjo_synth_blocks.restype = None
        L.cjo_pack_units.argtypes = [u8p, u8p, u8p, u8p, u8p, sz, C.c_int]
        L.cjo_pack_units.restype = None
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, status):
        super().__init__(STATUS.get(status, str(status)))
        self.status = status


def _buf(b):
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    return a, a.ctypes.data if a.size else 0


def _call(fn, src, cap, *extra):
    a, p = _buf(src)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    r = fn(p, a.size, out.ctypes.data, cap, *extra)
    if r < 0:
        raise OracleError(int(-r))
    return out[:r].tobytes()


def crc32c(b):
    a, p = _buf(b); return lib().cjo_crc32c(p, a.size)


def crc32c_masked(b):
    a, p = _buf(b); return lib().cjo_crc32c_masked(p, a.size)


def xxh32(b, seed=0):
    a, p = _buf(b); return lib().cjo_xxh32(p, a.size, seed)


def xxh64(b, seed=0):
    a, p = _buf(b); return lib().cjo_xxh64(p, a.size, seed)


def _len(fn, src):
    a, p = _buf(src)
    r = fn(p, a.size)
    if r < 0:
        raise OracleError(int(-r))
    return int(r)


def snappy_raw_compress(b):
    return _call(lib().cjo_snappy_raw_compress, b, lib().cjo_snappy_max_compressed_len(len(b)))


def snappy_raw_len(b):
    return _len(lib().cjo_snappy_raw_decompressed_len, b)


def snappy_raw_decompress(b, cap=None):
    return _call(lib().cjo_snappy_raw_decompress, b, snappy_raw_len(b) if cap is None else cap)


def snappy_frame_compress(b):
    return _call(lib().cjo_snappy_frame_compress, b, lib().cjo_snappy_frame_max_compressed_len(len(b)))


def snappy_frame_decompress(b, cap=None):
    return _call(lib().cjo_snappy_frame_decompress, b, _len(lib().cjo_snappy_frame_decompressed_len, b) if cap is None else cap)


def lz4_block_compress(b, acceleration=1):
    return _call(lib().cjo_lz4_block_compress, b, lib().cjo_lz4_compress_bound(len(b)), acceleration)


def lz4_block_decompress(b, cap):
    return _call(lib().cjo_lz4_block_decompress, b, cap)


def lz4f_compress(b, flags=3):
    return _call(lib().cjo_lz4f_compress, b, lib().cjo_lz4f_max_compressed_len(len(b)), flags)


def lz4f_decompress(b, cap=None):
    return _call(lib().cjo_lz4f_decompress, b, _len(lib().cjo_lz4f_decompressed_len, b) if cap is None else cap)


def zstd_len(b):
    return _len(lib().cjo_zstd_decompressed_len, b)


def zstd_decompress(b, cap=None):
    return _call(lib().cjo_zstd_decompress, b, zstd_len(b) if cap is None else cap)


def batch(codec, direction, src_base, src_off, src_len, dst_base, dst_off, dst_cap, nthreads=1):
    """Runs n units; returns (out_len int64[n], seconds)."""
    n = len(src_off)
    so = np.ascontiguousarray(src_off, dtype=np.uint64); sl = np.ascontiguousarray(src_len, dtype=np.uint64)
    do = np.ascontiguousarray(dst_off, dtype=np.uint64); dc = np.ascontiguousarray(dst_cap, dtype=np.uint64)
    out = np.empty(n, dtype=np.int64)
    sec = C.c_double(0)
    rc = lib().cjo_batch(codec, direction, n, src_base.ctypes.data, so.ctypes.data, sl.ctypes.data,
                         dst_base.ctypes.data, do.ctypes.data, dc.ctypes.data, out.ctypes.data, nthreads, C.byref(sec))
    assert rc == 0
    return out, sec.value
