"""ctypes binding of include/cramjam_cuda.h (libcramjam_cuda.so).

This is the only way Python code in this repo reaches the engine: tests, bench.py and the host
mirror all call the same extern "C" symbols a Rust/pyo3 or C++ host would bind.  There is no CPU
fallback: a missing library or a missing CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CJ_LIB_PATH") or os.path.join(HERE, "libcramjam_cuda.so")  # override: kernel A/B builds only

SNAPPY_RAW, SNAPPY_FRAMED, LZ4_BLOCK, LZ4_FRAME, ZSTD = range(5)
HOST, PINNED, DEVICE = range(3)

E_INVALID_ARG, E_NO_DEVICE, E_CUDA, E_NOMEM, E_UNIT_FAILED = -1, -2, -3, -4, -5


class Batch(C.Structure):
    _fields_ = [("n", C.c_size_t), ("src_base", C.c_void_p), ("src_off", C.c_void_p), ("src_len", C.c_void_p),
                ("dst_base", C.c_void_p), ("dst_off", C.c_void_p), ("dst_cap", C.c_void_p),
                ("dst_len", C.c_void_p), ("status", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("level", C.c_int32), ("acceleration", C.c_int32), ("flags", C.c_int32)]


class EngineError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"libcramjam_cuda error {rc}: {msg}")
        self.rc = rc


_lib = None


def lib():
    """Loads libcramjam_cuda.so (built in-tree by cramjam_b200/build.py).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: run `python -m cramjam_b200.build` (nvcc, sm_100a). "
                          "cramjam_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, sz, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int32
    sig = {
        "cj_abi_version": ([], C.c_int),
        "cj_device_count": ([], C.c_int),
        "cj_ctx_create": ([C.c_int, C.POINTER(vp)], C.c_int),
        "cj_ctx_destroy": ([vp], None),
        "cj_ctx_set_stream": ([vp, vp], C.c_int),
        "cj_ctx_synchronize": ([vp], C.c_int),
        "cj_last_error": ([], C.c_char_p),
        "cj_status_string": ([i32], C.c_char_p),
        "cj_ctx_launch_count": ([vp], u64),
        "cj_ctx_set_decode_path": ([vp, C.c_int, C.c_long], C.c_int),
        "cj_ctx_get_decode_path": ([vp, C.POINTER(C.c_int), C.POINTER(C.c_long)], C.c_int),
        "cj_ctx_last_redo_count": ([vp, C.POINTER(C.c_uint)], C.c_int),
        "cj_ctx_last_kernel_ms": ([vp, C.POINTER(C.c_float)], C.c_int),
        "cj_compress_bound": ([C.c_int, sz], sz),
        "cj_decompressed_len": ([C.c_int, vp, sz, C.POINTER(sz)], C.c_int),
        "cj_decompress_bound": ([C.c_int, vp, sz, C.POINTER(sz)], C.c_int),
        "cj_decompress_batch": ([vp, C.c_int, C.c_int, C.POINTER(Batch)], C.c_int),
        "cj_compress_batch": ([vp, C.c_int, C.c_int, C.POINTER(Batch), C.POINTER(Params)], C.c_int),
        "cj_decompress": ([vp, C.c_int, vp, sz, vp, sz, C.POINTER(sz)], C.c_int),
        "cj_compress": ([vp, C.c_int, vp, sz, vp, sz, C.POINTER(sz), C.POINTER(Params)], C.c_int),
        "cj_decompress_ex": ([vp, C.c_int, C.c_int, vp, sz, vp, sz, C.POINTER(sz)], C.c_int),
        "cj_compress_ex": ([vp, C.c_int, C.c_int, vp, sz, vp, sz, C.POINTER(sz), C.POINTER(Params)], C.c_int),
        "cj_host_register": ([vp, vp, sz], C.c_int),
        "cj_host_unregister": ([vp, vp], C.c_int),
        "cj_synth_blocks": ([vp, C.c_int, vp, sz, sz, u64, u64], C.c_int),
        "cj_copy_units": ([vp, sz, vp, vp, vp, vp, vp], C.c_int),
        "cj_device_alloc": ([vp, sz, C.POINTER(vp)], C.c_int),
        "cj_device_free": ([vp, vp], C.c_int),
        "cj_pinned_alloc": ([vp, sz, C.POINTER(vp)], C.c_int),
        "cj_pinned_free": ([vp, vp], C.c_int),
        "cj_memcpy_h2d": ([vp, vp, vp, sz], C.c_int),
        "cj_memcpy_d2h": ([vp, vp, vp, sz], C.c_int),
    }
    for name, (args, res) in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = res
    _lib = L
    return L


def last_error():
    return lib().cj_last_error().decode("utf-8", "replace")


def status_string(st):
    return lib().cj_status_string(int(st)).decode()


def _check(rc):
    if rc != 0:
        raise EngineError(rc, last_error())


def _ptr(a):
    """Address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


def synth_host(n_blocks, block_len, seed=0xC0FFEE, first_index=0):
    """Host side of the synthetic corpus generator (needs no GPU)."""
    out = np.empty(n_blocks * block_len, dtype=np.uint8)
    _check(lib().cj_synth_blocks(None, HOST, out.ctypes.data, n_blocks, block_len, seed, first_index))
    return out


class Context:
    """One engine context (device, stream, scratch arenas).  Thread-safe (internal lock)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().cj_ctx_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().cj_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle):
        _check(lib().cj_ctx_set_stream(self._h, cuda_stream_handle))

    def synchronize(self):
        _check(lib().cj_ctx_synchronize(self._h))

    def set_decode_path(self, generation, min_units=16384):
        """LZ4 / Snappy block decode kernels for batches of >= min_units units: 2 = one warp per block, 4 = one thread per
        block with 8-byte chunks (lz_decode4.cu), 7 = one thread per block with 16-byte chunks and linear per-lane records
        (lz_decode7.cu)."""
        _check(lib().cj_ctx_set_decode_path(self._h, generation, min_units))

    def last_redo_count(self):
        """Units of the most recent generation-4 batch that were handed to the generation-2 kernel."""
        v = C.c_uint()
        _check(lib().cj_ctx_last_redo_count(self._h, C.byref(v)))
        return v.value

    def decode_path(self):
        g, m = C.c_int(), C.c_long()
        _check(lib().cj_ctx_get_decode_path(self._h, C.byref(g), C.byref(m)))
        return g.value, m.value

    @property
    def launch_count(self):
        return int(lib().cj_ctx_launch_count(self._h))

    def last_kernel_ms(self):
        ms = C.c_float()
        _check(lib().cj_ctx_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def _batch(self, n, src_base, src_off, src_len, dst_base, dst_off, dst_cap, dst_len, status):
        return Batch(n, _ptr(src_base), _ptr(src_off), _ptr(src_len), _ptr(dst_base), _ptr(dst_off), _ptr(dst_cap),
                     _ptr(dst_len), _ptr(status))

    def decompress_batch(self, codec, where, n, src_base, src_off, src_len, dst_base, dst_off, dst_cap, dst_len, status):
        b = self._batch(n, src_base, src_off, src_len, dst_base, dst_off, dst_cap, dst_len, status)
        _check(lib().cj_decompress_batch(self._h, codec, where, C.byref(b)))

    def compress_batch(self, codec, where, n, src_base, src_off, src_len, dst_base, dst_off, dst_cap, dst_len, status,
                       level=-1, acceleration=1):
        b = self._batch(n, src_base, src_off, src_len, dst_base, dst_off, dst_cap, dst_len, status)
        p = Params(level, acceleration, 0)
        _check(lib().cj_compress_batch(self._h, codec, where, C.byref(b), C.byref(p)))

    def copy_units(self, n, src_base, src_off, lens, dst_base, dst_off):
        _check(lib().cj_copy_units(self._h, n, _ptr(src_base), _ptr(src_off), _ptr(lens), _ptr(dst_base), _ptr(dst_off)))

    def synth_device(self, dst, n_blocks, block_len, seed=0xC0FFEE, first_index=0):
        _check(lib().cj_synth_blocks(self._h, DEVICE, _ptr(dst), n_blocks, block_len, seed, first_index))

    # -- host conveniences (numpy in / numpy out) used by tests and the Python host mirror -----
    def run_host_units(self, codec, compress, units, caps, where=HOST, **kw):
        """units: list of bytes-like; caps: list of output capacities.
        Returns (list of bytes|None, status int32[n])."""
        n = len(units)
        lens = np.array([len(u) for u in units], dtype=np.uint64)
        so = np.zeros(n, dtype=np.uint64)
        if n:
            so[1:] = np.cumsum((lens[:-1] + 15) & ~np.uint64(15))
        src = np.zeros(int(so[-1] + lens[-1]) + 16 if n else 16, dtype=np.uint8)
        for i, u in enumerate(units):
            src[int(so[i]):int(so[i]) + len(u)] = np.frombuffer(u, dtype=np.uint8)
        dc = np.array(caps, dtype=np.uint64)
        do = np.zeros(n, dtype=np.uint64)
        if n:
            do[1:] = np.cumsum((dc[:-1] + 15) & ~np.uint64(15))
        dst = np.zeros(int(do[-1] + dc[-1]) + 16 if n else 16, dtype=np.uint8)
        dl = np.zeros(n, dtype=np.uint64)
        st = np.zeros(n, dtype=np.int32)
        if compress:
            self.compress_batch(codec, where, n, src, so, lens, dst, do, dc, dl, st, **kw)
        else:
            self.decompress_batch(codec, where, n, src, so, lens, dst, do, dc, dl, st)
        outs = [dst[int(do[i]):int(do[i] + dl[i])].tobytes() if st[i] == 0 else None for i in range(n)]
        return outs, st
