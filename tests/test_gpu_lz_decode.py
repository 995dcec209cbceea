"""-m gpu parity tests: LZ4 block and Snappy raw block decode kernels vs the CPU oracle, through
the C ABI (cj_decompress_batch), in host-staged and device-resident modes."""
import numpy as np
import pytest

import corpus
import oracle as O
import syslibs as S
from cramjam_b200 import _capi as capi
from gpu_util import arena, assert_same_as_oracle, ctx, gpu_decode_device, gpu_decode_host, oracle_batch

pytestmark = pytest.mark.gpu

CASES = corpus.edge_cases()


def test_lz4_howdy_golden():
    # reference tests/test_variants.py:329-334
    outs, st = gpu_decode_host(capi.LZ4_BLOCK, [b"\xe0howdy neighbor"], [14])
    assert st[0] == 0 and outs[0] == b"howdy neighbor"


@pytest.mark.parametrize("via", ["host", "device"])
def test_lz4_edge_cases_exact_capacity(via):
    units = [O.lz4_block_compress(d) for d in CASES]
    assert_same_as_oracle(capi.LZ4_BLOCK, units, [len(d) for d in CASES], via)


def test_lz4_liblz4_streams_and_capacity_variants():
    if not S.have_lz4:
        pytest.skip("no liblz4")
    units, caps = [], []
    for d in CASES:
        if not d:
            continue
        for kw in (dict(accel=1), dict(hc=9)):
            c = S.lz4_compress(d, **kw)
            for cap in (len(d), len(d) + 1, len(d) + 100, max(0, len(d) - 1), len(d) // 2, 0):
                units.append(c)
                caps.append(cap)
    assert_same_as_oracle(capi.LZ4_BLOCK, units, caps, "host")


def _mutations(stream, rng, count):
    out = []
    for _ in range(count):
        m = bytearray(stream)
        k = int(rng.integers(0, 4))
        if k == 0:
            m[int(rng.integers(0, len(m)))] = int(rng.integers(0, 256))
        elif k == 1:
            m = m[: int(rng.integers(0, len(m)))]
        elif k == 2:
            i = int(rng.integers(0, len(m)))
            m[i:i] = bytes([int(rng.integers(0, 256))])
        else:
            m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
        out.append(bytes(m))
    return out


def test_lz4_hostile_streams_match_oracle_status():
    rng = np.random.default_rng(42)
    units, caps = [], []
    for d in (corpus.text(3000, 1), corpus.lz_model(5000, 2), corpus.random_bytes(300, 3), b"a" * 700):
        c = O.lz4_block_compress(d)
        for m in _mutations(c, rng, 300):
            units.append(m)
            caps.append(len(d) + int(rng.integers(-8, 64)))
    units += [b"", b"\x00", b"sknow", b"\xff" * 40]
    caps += [10, 0, 100, 100]
    assert_same_as_oracle(capi.LZ4_BLOCK, units, caps, "device")


@pytest.mark.parametrize("via", ["host", "device"])
def test_snappy_edge_cases(via):
    units = [O.snappy_raw_compress(d) for d in CASES]
    assert_same_as_oracle(capi.SNAPPY_RAW, units, [len(d) for d in CASES], via)


def test_snappy_google_streams_and_capacity_variants():
    units, caps = [], []
    for d in CASES:
        if not d:
            continue
        c = S.snappy_compress(d)
        for cap in (len(d), len(d) + 77, max(0, len(d) - 1), 0):
            units.append(c)
            caps.append(cap)
    assert_same_as_oracle(capi.SNAPPY_RAW, units, caps, "host")


def test_snappy_hostile_streams_match_oracle_status():
    rng = np.random.default_rng(43)
    units, caps = [], []
    for d in (corpus.text(3000, 1), corpus.lz_model(5000, 2), corpus.random_bytes(300, 3), b"a" * 700):
        c = O.snappy_raw_compress(d)
        for m in _mutations(c, rng, 300):
            units.append(m)
            caps.append(len(d) + int(rng.integers(-8, 64)))
    units += [b"", b"\x00", b"sknow", b"\xff" * 40, b"\x05\xfc\xff\xff\xff\xff"]
    caps += [10, 0, 100, 100, 100]
    assert_same_as_oracle(capi.SNAPPY_RAW, units, caps, "device")


def test_snappy_copy4_and_long_literal_tags():
    # hand-built stream: literal "abcdefgh", copy-4 (tag 3) of len 8 offset 8, long literal via tag 61
    lit = bytes(range(97, 105))
    long_lit = corpus.random_bytes(300, 9)
    body = bytes([(8 - 1) << 2]) + lit + bytes([((8 - 1) << 2) | 3]) + (8).to_bytes(4, "little")
    body += bytes([61 << 2]) + (300 - 1).to_bytes(2, "little") + long_lit
    total = 8 + 8 + 300
    stream = bytes([total & 0x7f | 0x80, total >> 7]) + body
    assert O.snappy_raw_decompress(stream) == lit + lit + long_lit
    outs, st = gpu_decode_host(capi.SNAPPY_RAW, [stream], [total])
    assert st[0] == 0 and outs[0] == lit + lit + long_lit


@pytest.mark.parametrize("codec", [capi.LZ4_BLOCK, capi.SNAPPY_RAW])
def test_synthetic_64k_blocks_device_resident(codec):
    """BASELINE config shape at a size the oracle finishes in seconds: 1024 x 64 KiB synthetic blocks,
    unaligned unit starts included."""
    n, U = 1024, 65536
    data = capi.synth_host(n, U, seed=0xC0FFEE, first_index=0)
    comp = O.lz4_block_compress if codec == capi.LZ4_BLOCK else O.snappy_raw_compress
    units = [comp(data[i * U:(i + 1) * U].tobytes()) for i in range(n)]
    for align, lead in ((16, 0), (1, 3)):
        src, so, sl = arena(units, align=align, lead=lead)
        caps = np.full(n, U, dtype=np.uint64)
        gdst, gdo, glen, st = gpu_decode_device(codec, src, so, sl, caps, dst_align=align, dst_lead=lead)
        assert (st == 0).all() and (glen == U).all()
        for i in range(n):
            assert np.array_equal(gdst[int(gdo[i]):int(gdo[i]) + U], data[i * U:(i + 1) * U]), (i, align)


@pytest.mark.parametrize("codec", [capi.LZ4_BLOCK, capi.SNAPPY_RAW])
def test_large_single_blocks(codec):
    """Blocks far larger than the shared-memory ring (far back-references served from L2/HBM)."""
    big = corpus.lz_model(3_000_000, 5) + corpus.text(2_000_000, 6) + b"z" * 1_000_000 + corpus.lz_model(3_000_000, 5)[:500_000]
    comp = O.lz4_block_compress if codec == capi.LZ4_BLOCK else O.snappy_raw_compress
    outs, st = gpu_decode_host(codec, [comp(big)], [len(big)])
    assert st[0] == 0 and outs[0] == big


def test_pinned_path_matches_host_path():
    n, U = 64, 65536
    data = capi.synth_host(n, U, seed=5)
    units = [O.lz4_block_compress(data[i * U:(i + 1) * U].tobytes()) for i in range(n)]
    a, _ = gpu_decode_host(capi.LZ4_BLOCK, units, [U] * n, where=capi.HOST)
    # CJ_PINNED with ordinary numpy memory is still legal for cudaMemcpyAsync (it just is not DMA-direct)
    b, _ = gpu_decode_host(capi.LZ4_BLOCK, units, [U] * n, where=capi.PINNED)
    assert a == b == [data[i * U:(i + 1) * U].tobytes() for i in range(n)]


@pytest.mark.parametrize("codec", [capi.LZ4_BLOCK, capi.SNAPPY_RAW])
def test_pinned_pipelined_path(codec):
    """Dense arenas >= 64 MiB through CJ_PINNED take the chunked H2D | kernel | D2H pipeline; one chunk holds a
    corrupt unit and must fall back to per-unit copies without disturbing its neighbours."""
    n, U = 1536, 65536
    data = capi.synth_host(n, U, seed=77)
    comp = O.lz4_block_compress if codec == capi.LZ4_BLOCK else O.snappy_raw_compress
    units = [comp(data[i * U:(i + 1) * U].tobytes()) for i in range(n)]
    bad = 700
    units[bad] = units[bad][:100]                      # truncated stream -> error status
    src, so, sl = arena(units)
    do = np.arange(n, dtype=np.uint64) * U
    dc = np.full(n, U, dtype=np.uint64)
    dst = np.full(n * U, 0xAB, dtype=np.uint8)
    dl = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=np.int32)
    ctx().decompress_batch(codec, capi.PINNED, n, src, so, sl, dst, do, dc, dl, st)
    assert st[bad] != 0 and (np.delete(st, bad) == 0).all()
    assert (dst[bad * U:(bad + 1) * U] == 0xAB).all()   # nothing written for the failed unit
    ok = np.ones(n * U, dtype=bool)
    ok[bad * U:(bad + 1) * U] = False
    assert np.array_equal(dst[ok], data[ok])
