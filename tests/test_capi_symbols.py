"""No-GPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/cramjam_cuda.h declares; pure-host helpers behave; compute entry points fail loudly
(never fall back to a CPU path) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from cramjam_b200 import build
    build.build()
    from cramjam_b200 import _capi
    return _capi


def test_header_symbols_exported():
    capi = _lib()
    hdr = open(os.path.join(ROOT, "include", "cramjam_cuda.h")).read()
    names = sorted(set(re.findall(r"\b(cj_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = C.CDLL(capi.SO_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert capi.lib().cj_abi_version() == 1


def test_compress_bounds():
    capi = _lib()
    L = capi.lib()
    assert L.cj_compress_bound(capi.SNAPPY_RAW, 65536) == 76490      # 32 + n + n/6   
    assert L.cj_compress_bound(capi.LZ4_BLOCK, 65536) == 65809       # LZ4_compressBound
    assert L.cj_compress_bound(capi.LZ4_BLOCK, 0x7E000001) == 0


def _frame_piece(n):
    """Chunk / block size the frame encoders cut an input of n bytes into (frames.cu frame_piece)."""
    if n >= 4 << 20:
        return 65536
    return min(65536, max(16384, (n // 64 + 4095) & ~4095))


def test_frame_bounds_cover_the_finer_chunks_of_small_inputs():
    """Inputs under 4 MiB are cut into 16-64 KiB chunks (more warps per call): the bounds must hold an incompressible input,
    every chunk stored raw with its own header (the case a 65 535-byte random input once missed by 3 bytes)."""
    capi = _lib()
    L = capi.lib()
    for n in (0, 1, 15, 16383, 16384, 16385, 65535, 65536, 65537, 200000, (1 << 20) - 1, 1 << 20, (4 << 20) - 1, 4 << 20, (4 << 20) + 1, 100 << 20):
        piece = _frame_piece(n)
        chunks = (n + piece - 1) // piece
        assert chunks <= max(64, n // 65536 + 1)
        assert L.cj_compress_bound(capi.SNAPPY_FRAMED, n) >= 10 + chunks * 8 + n          # stream id + (type, len, crc) per chunk
        assert L.cj_compress_bound(capi.LZ4_FRAME, n) >= 15 + chunks * 4 + n + 8          # header + block sizes + end mark + content checksum


def test_synth_host_is_deterministic_and_shaped():
    capi = _lib()
    a = capi.synth_host(8, 65536, seed=1, first_index=5)
    b = capi.synth_host(4, 65536, seed=1, first_index=7)
    assert a.size == 8 * 65536
    assert np.array_equal(a[2 * 65536:6 * 65536], b)                 # function of (seed, global index) only


def test_no_device_fails_loudly():
    capi = _lib()
    if capi.lib().cj_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.EngineError) as e:
        capi.Context(0)
    assert e.value.rc == capi.E_NO_DEVICE
    assert "no CPU fallback" in str(e.value)
