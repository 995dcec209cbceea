"""-m gpu parity tests of the thread-per-block decode paths (generation 4, lz_decode4.cu, and generation 7, lz_decode7.cu:
Snappy raw and LZ4 block) with their generation-2 redo list, each forced on through cj_ctx_set_decode_path() for every batch size, and of batches beyond one
wave of the kernel (the multi-round path).  Same oracle, same status-code expectations as test_gpu_lz_decode.py."""
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


@pytest.fixture(autouse=True, params=[4, 7], ids=["gen4", "gen7"])
def gen4(request):
    default = ctx().decode_path()
    ctx().set_decode_path(request.param, 1)
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
    if ctx().decode_path()[0] in (4, 7):   # well-formed, aligned blocks are decoded by the thread-per-block kernel itself, not by its fallback
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


@pytest.mark.parametrize("codec", [capi.SNAPPY_RAW, capi.LZ4_BLOCK])
def test_multi_round_batch_configs4_size(codec):
    """131 072 x 64 KiB blocks per GPU (BASELINE.json configs[4] per-GPU size) is more than one wave of the
    thread-per-block kernel (148 SMs x 20 warps x 32 lanes = 94 720 blocks): the batch is decoded in several rounds.
    A 4 096-block sample is encoded by the ORACLE encoder (not the GPU encoder); the batch tiles the sample's
    descriptors 32 times (indices, not bytes) into 131 072 distinct output slots, and every slot must equal the
    generator's original block."""
    import torch
    S_, U, REP = 4096, 65536, 32
    n = S_ * REP
    data = capi.synth_host(S_, U, seed=0xC0FFEE, first_index=7000)
    bound = lambda k: (32 + k + k // 6) if codec == capi.SNAPPY_RAW else (k + k // 255 + 16)
    slot = (bound(U) + 15) // 16 * 16
    src = np.zeros(S_ * slot + 64, dtype=np.uint8)
    so1 = np.arange(S_, dtype=np.uint64) * np.uint64(U)
    do1 = np.arange(S_, dtype=np.uint64) * np.uint64(slot)
    lens = O.batch(O.SNAPPY_RAW if codec == capi.SNAPPY_RAW else O.LZ4_BLOCK, 1, data, so1, np.full(S_, U, dtype=np.uint64), src, do1,
                   np.full(S_, slot, dtype=np.uint64), nthreads=16)[0]
    assert (lens > 0).all()
    dev = torch.device("cuda:0")
    as_i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
    t_src = torch.from_numpy(src).to(dev)
    t_so = as_i64(np.tile(do1, REP))
    t_sl = as_i64(np.tile(lens.astype(np.uint64), REP))
    t_do = as_i64(np.arange(n, dtype=np.uint64) * np.uint64(U))
    t_dc = as_i64(np.full(n, U, dtype=np.uint64))
    t_dl = torch.zeros(n, dtype=torch.int64, device=dev)
    t_st = torch.full((n,), -99, dtype=torch.int32, device=dev)
    t_dst = torch.zeros(n * U + 64, dtype=torch.uint8, device=dev)
    c = ctx()
    torch.cuda.synchronize()   # the context runs on its own stream: the tensors above must exist before it starts
    c.decompress_batch(codec, capi.DEVICE, n, t_src, t_so, t_sl, t_dst, t_do, t_dc, t_dl, t_st)
    c.synchronize()
    assert int((t_st != 0).sum()) == 0 and int((t_dl != U).sum()) == 0
    assert c.last_redo_count() == 0
    want = torch.from_numpy(data).to(dev)
    got = t_dst[:n * U].view(REP, S_ * U)
    for r in range(REP):
        assert torch.equal(got[r], want), r
