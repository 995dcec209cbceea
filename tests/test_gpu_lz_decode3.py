"""-m gpu parity tests of the alternative block-decode paths — generation 3 (lz_decode3.cu: index walk + lane state
machines), generation 4 (lz_decode4.cu: one thread per block) and 5 (4 and 2 side by side on a split batch), each with
the generation-2 redo list — forced on through cj_ctx_set_decode_path() for every batch size.
Same oracle, same status-code expectations as test_gpu_lz_decode.py."""
import numpy as np
import pytest

import corpus
import oracle as O
import syslibs as S
from cramjam_b200 import _capi as capi
from gpu_util import arena, assert_same_as_oracle, ctx, gpu_decode_device, oracle_batch
from test_gpu_lz_decode import _mutations

pytestmark = pytest.mark.gpu

CASES = corpus.edge_cases()


@pytest.fixture(autouse=True, params=[3, 4, 5], ids=["gen3", "gen4", "gen5"])
def gen3(request):
    default = ctx().decode_path()
    ctx().set_decode_path(request.param, 1)   # generations 4 and 5 take Snappy only; LZ4 batches stay on generation 2
    yield
    ctx().set_decode_path(*default)


@pytest.mark.parametrize("codec", [capi.SNAPPY_RAW, capi.LZ4_BLOCK])
def test_edge_cases(codec):
    comp = O.snappy_raw_compress if codec == capi.SNAPPY_RAW else O.lz4_block_compress
    units = [comp(d) for d in CASES]
    assert_same_as_oracle(codec, units, [len(d) for d in CASES], "device")


@pytest.mark.parametrize("codec", [capi.SNAPPY_RAW, capi.LZ4_BLOCK])
def test_synthetic_blocks_bit_exact(codec):
    n, U = 600, 65536
    data = capi.synth_host(n, U, seed=0xC0FFEE, first_index=100)
    comp = O.snappy_raw_compress if codec == capi.SNAPPY_RAW else O.lz4_block_compress
    units = [comp(data[i * U:(i + 1) * U].tobytes()) for i in range(n)]
    src, so, sl = arena(units)
    dst, do, dl, st = gpu_decode_device(codec, src, so, sl, [U] * n)
    assert (st == 0).all() and (dl == U).all()
    assert np.array_equal(dst[:n * U], data)
    if ctx().decode_path()[0] == 4:   # well-formed, aligned blocks are decoded by the thread-per-block kernel itself, not by its fallback
        assert ctx().last_redo_count() == 0


@pytest.mark.parametrize("codec", [capi.SNAPPY_RAW, capi.LZ4_BLOCK])
def test_hostile_streams_match_oracle_status(codec):
    rng = np.random.default_rng(7)
    comp = O.snappy_raw_compress if codec == capi.SNAPPY_RAW else O.lz4_block_compress
    units, caps = [], []
    for d in (corpus.text(3000, 1), corpus.lz_model(5000, 2), corpus.random_bytes(300, 3), b"a" * 700, corpus.lz_model(70000, 5)):
        c = comp(d)
        for m in _mutations(c, rng, 200):
            units.append(m)
            caps.append(len(d) + int(rng.integers(-8, 64)))
    units += [b"", b"\x00", b"sknow", b"\xff" * 40]
    caps += [10, 0, 100, 100]
    assert_same_as_oracle(codec, units, caps, "device")


def test_capacity_variants_and_system_encoders():
    units, caps, codecs = [], [], []
    for d in CASES:
        if not d:
            continue
        cs = [O.lz4_block_compress(d)]
        if S.have_lz4:
            cs += [S.lz4_compress(d, accel=1), S.lz4_compress(d, hc=9)]
        for c in cs:
            for cap in (len(d), len(d) + 1, len(d) + 100, max(0, len(d) - 1), len(d) // 2, 0):
                units.append(c)
                caps.append(cap)
    assert_same_as_oracle(capi.LZ4_BLOCK, units, caps, "device")
    units, caps = [], []
    for d in CASES:
        c = O.snappy_raw_compress(d)
        for cap in (len(d), len(d) + 7, max(0, len(d) - 1)):
            units.append(c)
            caps.append(cap)
    assert_same_as_oracle(capi.SNAPPY_RAW, units, caps, "device")


@pytest.mark.parametrize("codec", [capi.SNAPPY_RAW, capi.LZ4_BLOCK])
def test_unaligned_units_take_the_redo_path(codec):
    # unit starts that are not 16-byte aligned are not indexed; the generation-2 kernel decodes them
    comp = O.snappy_raw_compress if codec == capi.SNAPPY_RAW else O.lz4_block_compress
    datas = [corpus.lz_model(4000 + 37 * i, i) for i in range(40)]
    units = [comp(d) for d in datas]
    src, so, sl = arena(units, align=1, lead=3)
    odst, odo, olen = oracle_batch(codec, 0, src, so, sl, [len(d) for d in datas])
    dst, do, dl, st = gpu_decode_device(codec, src, so, sl, [len(d) for d in datas], dst_align=1, dst_lead=5)
    assert (st == 0).all()
    for i, d in enumerate(datas):
        assert dst[int(do[i]):int(do[i]) + len(d)].tobytes() == d


def test_long_literals_and_long_matches():
    # incompressible data (one long literal), long runs (matches far beyond 64 bytes, offset 1), and a mix
    rng = np.random.default_rng(3)
    datas = [rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(), b"\x00" * 100000, b"ab" * 40000,
             bytes(rng.integers(0, 256, 5000, dtype=np.uint8)) + b"z" * 3000 + bytes(rng.integers(0, 4, 9000, dtype=np.uint8))]
    for codec, comp in ((capi.SNAPPY_RAW, O.snappy_raw_compress), (capi.LZ4_BLOCK, O.lz4_block_compress)):
        units = [comp(d) for d in datas]
        assert_same_as_oracle(codec, units, [len(d) for d in datas], "device")
