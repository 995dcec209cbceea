"""CPU tests of the generation-7 lane program (cramjam_b200/csrc/lz_decode7.cuh) through its host-side emulation
(tests/emu/g7_emu.cpp): the same source the device kernel compiles, run lane by lane on the host with every asynchronous
copy delivered either at once or at the last moment the program's wait_group allows.  Accepted blocks must be bit-exact with
the oracle; a block the lane declines (-> redo list of the warp-per-block kernel on the device) is fine, a block the oracle
rejects must never be accepted with different bytes, and no invariant of the emulated machine may be violated."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import corpus
import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "g7_emu.cpp")
HDR = os.path.join(HERE, "..", "cramjam_b200", "csrc", "lz_decode7.cuh")
SO = os.path.join(HERE, "emu", "_g7_emu.so")
SNAPPY, LZ4 = 0, 2


def _lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.g7_emu_decode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_long)]
    L.g7_emu_decode.restype = C.c_long
    return L


def _aligned(n, shift=0):
    """n bytes at an address that is `shift` (0 or 16) past a 32-byte boundary."""
    raw = np.zeros(n + 96, dtype=np.uint8)
    o = (-raw.ctypes.data) % 32 + shift
    return raw[o:o + n]


def emu(codec, comp, cap, depth=2, mode=1, extra=7, dst_shift=None):
    """-> (result, bytes, stats): result = decoded length, -1 = declined (redo list)."""
    L = _lib()
    src = _aligned(max(len(comp), 1))
    src[:len(comp)] = np.frombuffer(comp, dtype=np.uint8)
    if dst_shift is None:
        dst_shift = 16 * ((len(comp) + cap) & 1)   # finished granules leave in 32-byte pairs: both phases of the output address
    dst = _aligned(cap + 16, dst_shift)
    dst[:] = 0xEE
    stats = (C.c_long * 3)()
    r = L.g7_emu_decode(codec, depth, src.ctypes.data, len(comp), dst.ctypes.data, cap, mode, extra, stats)
    assert r > -100, f"emulated machine invariant {-(r + 100)} violated"
    assert r != -2, "lane neither finished nor declined"
    assert bytes(dst[cap:cap + 16]) == b"\xEE" * 16, "wrote beyond the capacity"
    return r, dst[:max(r, 0)].tobytes(), list(stats)


def oracle_decode(codec, comp, cap):
    try:
        return O.snappy_raw_decompress(comp, cap) if codec == SNAPPY else O.lz4_block_decompress(comp, cap)
    except O.OracleError:
        return None


def check(codec, comp, cap, **kw):
    want = oracle_decode(codec, comp, cap)
    r, got, stats = emu(codec, comp, cap, **kw)
    if r >= 0:
        assert want is not None, "accepted a block the oracle rejects"
        assert got == want
    return r, stats


COMP = {SNAPPY: O.snappy_raw_compress, LZ4: O.lz4_block_compress}


@pytest.mark.parametrize("codec", [SNAPPY, LZ4])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("depth", [2, 3, 4])
def test_edge_cases(codec, mode, depth):
    accepted = 0
    for d in corpus.edge_cases():
        if len(d) > 70000 and depth != 2:
            continue
        for shift in (0, 16):
            r, _ = check(codec, COMP[codec](d), len(d), depth=depth, mode=mode, dst_shift=shift)
        accepted += r >= 0
        if len(d) >= 1:
            assert r == len(d), f"well-formed aligned block of {len(d)} bytes declined"
    assert accepted >= 30


@pytest.mark.parametrize("codec", [SNAPPY, LZ4])
@pytest.mark.parametrize("mode", [0, 1])
def test_synthetic_blocks(codec, mode):
    U = 65536
    data = O.synth(48, U, seed=0xC0FFEE, first_index=100)
    for i in range(48):
        blk = data[i * U:(i + 1) * U].tobytes()
        r, _ = check(codec, COMP[codec](blk), U, mode=mode)
        assert r == U


@pytest.mark.parametrize("codec", [SNAPPY, LZ4])
def test_short_offsets_and_periods(codec):
    rng = np.random.default_rng(11)
    for period in list(range(1, 40)) + [47, 48, 49, 63, 64, 65, 79, 80, 81, 100, 127, 128, 129]:
        base = rng.integers(0, 256, size=period, dtype=np.uint8).tobytes()
        for total in (period + 1, period + 17, 300, 5000):
            d = (base * (total // period + 2))[:total]
            for mode in (0, 1):
                r, _ = check(codec, COMP[codec](d), len(d), mode=mode)
                assert r == len(d)
    for seed in range(30):   # dense mixture of short and long offsets
        d = corpus.lz_model(3000 + 97 * seed, seed, lit_mean=2.0, match_mean=12.0)
        for mode in (0, 1):
            r, _ = check(codec, COMP[codec](d), len(d), mode=mode)
            assert r == len(d)


@pytest.mark.parametrize("codec", [SNAPPY, LZ4])
def test_capacity_variants(codec):
    for d in (corpus.text(3000, 1), corpus.lz_model(5000, 2), b"a" * 700, corpus.random_bytes(300, 3)):
        c = COMP[codec](d)
        for cap in (len(d), len(d) + 1, len(d) + 100, len(d) - 1, len(d) // 2, 1):
            check(codec, c, cap)


@pytest.mark.parametrize("codec", [SNAPPY, LZ4])
def test_hostile_streams(codec):
    from test_gpu_lz_decode import _mutations
    rng = np.random.default_rng(7)
    n = acc = 0
    for d in (corpus.text(3000, 1), corpus.lz_model(5000, 2), corpus.random_bytes(300, 3), b"a" * 700):
        c = COMP[codec](d)
        for m in _mutations(c, rng, 250):
            if not m:
                continue
            cap = max(1, len(d) + int(rng.integers(-8, 64)))
            r, _ = check(codec, m, cap, mode=int(rng.integers(0, 2)))
            n += 1
            acc += r >= 0
    assert acc > n // 20   # mutations that leave the stream valid are still decoded by the lane itself


@pytest.mark.parametrize("codec", [SNAPPY, LZ4])
def test_real_corpus_blocks(codec):
    """64 KiB blocks of the six Silesia files the reference's benchmarks ship (tests/golden/corpus), encoded by the oracle and, where
    they are installed, by Google snappy / liblz4 (fast and HC): text, XML, a chemical database with very long matches, an
    executable, a medical image — element mixes the synthetic generator does not produce."""
    import bz2
    import syslibs as S
    U = 65536
    cdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "corpus")
    encoders = [COMP[codec]]
    if codec == SNAPPY:
        try:
            S.snappy_compress(b"probe")
            encoders.append(S.snappy_compress)   # Google snappy (pyarrow's copy)
        except Exception:
            pass
    if codec == LZ4 and S.have_lz4:
        encoders += [lambda d: S.lz4_compress(d, accel=1), lambda d: S.lz4_compress(d, hc=9)]
    n = 0
    for f in sorted(os.listdir(cdir)):
        if not f.endswith(".bz2"):
            continue
        raw = bz2.decompress(open(os.path.join(cdir, f), "rb").read())
        nb = len(raw) // U
        for b in (0, nb // 2, nb - 1):
            blk = raw[b * U:(b + 1) * U]
            for e, enc in enumerate(encoders):
                r, _ = check(codec, enc(blk), U, depth=3, mode=(b + e) & 1)
                assert r == U, (f, b, e)
                n += 1
    assert n >= 18
