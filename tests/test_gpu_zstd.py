"""-m gpu tests: Zstandard frame decode kernel vs the CPU oracle / libzstd, through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

import corpus
import oracle as O
import syslibs as S
from cramjam_b200 import _capi as capi
from gpu_util import assert_same_as_oracle, ctx

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PLAINTEXT = open(os.path.join(G, "plaintext.txt"), "rb").read()
CASES = corpus.edge_cases()


def test_golden_fixture():
    # reference tests/test_integration.py:32-50, row zstd/zst
    f = open(os.path.join(G, "plaintext.txt.zst"), "rb").read()
    outs, st = ctx().run_host_units(capi.ZSTD, False, [f], [len(PLAINTEXT)])
    assert st[0] == 0 and outs[0] == PLAINTEXT
    out = C.c_size_t()
    assert capi.lib().cj_decompressed_len(capi.ZSTD, f, len(f), C.byref(out)) == 0 and out.value == 857


def test_sknow_is_an_error():
    outs, st = ctx().run_host_units(capi.ZSTD, False, [b"sknow"], [100])
    assert st[0] != 0


@pytest.mark.skipif(not S.have_zstd, reason="libzstd.so.1 not present")
@pytest.mark.parametrize("kw", [dict(level=1), dict(level=3), dict(level=3, checksum=True), dict(level=9), dict(level=19),
                                dict(level=3, window_log=17), dict(level=-5), dict(level=3, content_size=False)])
def test_libzstd_frames_decode_bit_exact(kw):
    units = [S.zstd_compress(d, **kw) for d in CASES]
    outs, st = ctx().run_host_units(capi.ZSTD, False, units, [len(d) for d in CASES])
    assert (st == 0).all(), [(i, int(s)) for i, s in enumerate(st) if s]
    assert outs == CASES


@pytest.mark.skipif(not S.have_zstd, reason="libzstd.so.1 not present")
def test_concatenated_and_skippable_frames_and_capacity():
    a, b = S.zstd_compress(b"one " * 1000), S.zstd_compress(corpus.text(5000, 9), checksum=True)
    skip = (0x184D2A50).to_bytes(4, "little") + (3).to_bytes(4, "little") + b"abc"
    want = b"one " * 1000 + corpus.text(5000, 9)
    outs, st = ctx().run_host_units(capi.ZSTD, False, [a + skip + b, a + skip + b, b""], [len(want), len(want) - 1, 0])
    assert st[0] == 0 and outs[0] == want
    assert st[1] == 5                                      # DST_SMALL
    assert st[2] == 0 and outs[2] == b""                   # empty input decodes to nothing (streaming decoder semantics)


@pytest.mark.skipif(not S.have_zstd, reason="libzstd.so.1 not present")
def test_hostile_frames_match_oracle():
    rng = np.random.default_rng(11)
    units, caps = [], []
    for d, kw in ((corpus.lz_model(20000, 3), dict(level=3, checksum=True)), (corpus.text(30000, 4), dict(level=1)), (b"q" * 5000, dict(level=3))):
        f = S.zstd_compress(d, **kw)
        for _ in range(250):
            m = bytearray(f)
            k = int(rng.integers(0, 3))
            if k == 0:
                m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
            elif k == 1:
                m = m[: int(rng.integers(0, len(m)))]
            else:
                m[int(rng.integers(0, len(m)))] = int(rng.integers(0, 256))
            units.append(bytes(m))
            caps.append(len(d))
    assert_same_as_oracle(capi.ZSTD, units, caps, "host")


@pytest.mark.skipif(not S.have_zstd, reason="libzstd.so.1 not present")
def test_synthetic_256k_frames_level3():
    """BASELINE configs[3] shape at a size the oracle finishes in seconds: 128 x 256 KiB level-3 frames."""
    n, U = 128, 262144
    data = capi.synth_host(n * 4, 65536, seed=0xC0FFEE)
    blocks = [data[i * U:(i + 1) * U].tobytes() for i in range(n)]
    units = [S.zstd_compress(b, 3) for b in blocks]
    outs, st = ctx().run_host_units(capi.ZSTD, False, units, [U] * n)
    assert (st == 0).all() and outs == blocks


# ---- encoder (reference src/zstd.rs:37-64; SURVEY 8f "next" row, first real encoder) ----
def _zcompress(units):
    return ctx().run_host_units(capi.ZSTD, True, units, [capi.lib().cj_compress_bound(capi.ZSTD, len(u)) for u in units])


@pytest.mark.skipif(not S.have_zstd, reason="libzstd.so.1 not present")
def test_zstd_encode_is_valid_for_libzstd_oracle_and_own_decoder():
    outs, st = _zcompress(CASES)
    assert (st == 0).all()
    for d, c in zip(CASES, outs):
        assert c[:4] == b"\x28\xb5\x2f\xfd"
        assert S.zstd_decompress(c, len(d)) == d, len(d)       # libzstd accepts and reproduces
        assert O.zstd_len(c) == len(d)                          # pledged content size in the header
        assert O.zstd_decompress(c) == d
    back, st2 = ctx().run_host_units(capi.ZSTD, False, outs, [len(d) for d in CASES])
    assert (st2 == 0).all() and back == CASES


def test_zstd_encode_compresses_and_is_deterministic():
    n, U = 64, 262144
    data = capi.synth_host(n * 4, 65536, seed=3)
    units = [data[i * U:(i + 1) * U].tobytes() for i in range(n)]
    a, st = _zcompress(units)
    b, _ = _zcompress(units)
    assert (st == 0).all() and a == b
    ratio = n * U / sum(len(c) for c in a)
    assert ratio > 2.0, ratio                                   # greedy LZ77 + predefined FSE tables + Huffman literals
    for u, c in zip(units[::8], a[::8]):
        assert O.zstd_decompress(c) == u
    outs, st = _zcompress([corpus.text(5000, 1)])
    d = corpus.text(5000, 1)
    small, st = ctx().run_host_units(capi.ZSTD, True, [d], [10])
    assert st[0] == 5                                           # DST_SMALL


def test_large_single_buffer_round_trip_is_multi_frame_and_libzstd_readable():
    """Inputs above 512 KiB are written as concatenated independent frames (one warp each) and read back
    frame-parallel; libzstd and the oracle must read the stream, and a libzstd-made multi-frame stream must decode."""
    data = capi.synth_host(48, 65536, seed=7, first_index=3).tobytes() + b"tail" * 1000   # 3 MiB + 4000 B
    bound = capi.lib().cj_compress_bound(capi.ZSTD, len(data))
    enc, st = ctx().run_host_units(capi.ZSTD, True, [data, data[:700000], b""], [bound, bound, 64])
    assert (st == 0).all()
    assert enc[0].count(b"\x28\xb5\x2f\xfd") >= 7          # 3 MiB + tail -> 7 frames
    for c, d in zip(enc, (data, data[:700000], b"")):
        assert O.zstd_decompress(c) == d
        if S.have_zstd:
            assert S.zstd_decompress(c, len(d)) == d
    outs, st = ctx().run_host_units(capi.ZSTD, False, enc, [len(data), 700000, 0])
    assert (st == 0).all() and outs[0] == data and outs[1] == data[:700000] and outs[2] == b""
    if S.have_zstd:
        multi = b"".join(S.zstd_compress(data[i:i + 300000], 3) for i in range(0, len(data), 300000))
        skippable = b"\x50\x2a\x4d\x18" + (5).to_bytes(4, "little") + b"hello"
        stream = skippable + multi
        rng = np.random.default_rng(5)
        units, caps = [stream, stream, stream[:-3], multi + b"\x00"], [len(data), len(data) - 1, len(data), len(data)]
        for _ in range(24):   # corrupt one byte somewhere: the frame-parallel path must hand over to the whole-stream path
            m = bytearray(stream)
            m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
            units.append(bytes(m)); caps.append(len(data))
        assert_same_as_oracle(capi.ZSTD, units, caps, "host")


@pytest.mark.skipif(not S.have_zstd, reason="libzstd.so.1 not present")
def test_zstd_encode_levels_and_huffman_literals():
    """level 1-2 (and negative levels) store literals raw, every other level — the reference default included —
    Huffman-codes them (zstd_huf.cuh): both kinds must be valid for libzstd / the oracle / our decoder, the Huffman
    ones smaller; binary data whose alphabet does not fit the direct tree description falls back to raw literals."""
    texts = [corpus.text(n, n) for n in (300, 5000, 70000, 262144)] + [capi.synth_host(4, 65536, seed=5).tobytes(), corpus.lz_model(150000, 9)]
    binary = [corpus.random_bytes(40000, 3), bytes(np.random.default_rng(1).integers(0, 200, 90000, dtype=np.uint8))]
    units = texts + binary
    bound = [capi.lib().cj_compress_bound(capi.ZSTD, len(u)) for u in units]
    sizes = {}
    for level in (1, 3, 0, 9, -3):
        outs, st = ctx().run_host_units(capi.ZSTD, True, units, bound, level=level)
        assert (st == 0).all(), level
        for d, c in zip(units, outs):
            assert S.zstd_decompress(c, len(d)) == d, (level, len(d))
            assert O.zstd_decompress(c) == d
        back, st2 = ctx().run_host_units(capi.ZSTD, False, outs, [len(d) for d in units])
        assert (st2 == 0).all() and back == units
        sizes[level] = [len(c) for c in outs]
    assert sizes[3] == sizes[0] == sizes[9] and sizes[1] == sizes[-3]
    for i in range(len(units)):
        assert sizes[3][i] <= sizes[1][i], (i, sizes[3][i], sizes[1][i])
    assert sizes[3][4] < sizes[1][4] * 0.9, (sizes[3][4], sizes[1][4])   # the synthetic corpus: literals are ~45 % of the frame
    assert sum(sizes[3]) < sum(sizes[1]) * 0.97


def _block_modes(f):
    """(literal section types, sequence table modes) of the compressed blocks of one zstd frame — header walk only."""
    p = 4
    fhd = f[p]; p += 1
    single, fcs, did = (fhd >> 5) & 1, fhd >> 6, fhd & 3
    p += (0 if single else 1) + [0, 1, 2, 4][did] + [1 if single else 0, 2, 4, 8][fcs]
    lit, seq = [], []
    while True:
        h = f[p] | f[p + 1] << 8 | f[p + 2] << 16; p += 3
        last, bt, bs = h & 1, (h >> 1) & 3, h >> 3
        if bt == 2:
            b0 = f[p]; lt, sf = b0 & 3, (b0 >> 2) & 3
            lit.append(lt)
            if lt < 2:
                hdr, regen = (1, b0 >> 3) if sf in (0, 2) else ((2, (b0 >> 4) | (f[p + 1] << 4)) if sf == 1 else (3, (b0 >> 4) | (f[p + 1] << 4) | (f[p + 2] << 12)))
                q = p + hdr + (regen if lt == 0 else 1)
            else:
                v = int.from_bytes(f[p:p + 5], "little")
                hdr, comp = (3, (v >> 14) & 0x3ff) if sf in (0, 1) else ((4, (v >> 18) & 0x3fff) if sf == 2 else (5, (v >> 22) & 0x3ffff))
                q = p + hdr + comp
            n0 = f[q]
            if n0:
                m = f[q + (1 if n0 < 128 else (2 if n0 < 255 else 3))]
                seq.append(((m >> 6) & 3, (m >> 4) & 3, (m >> 2) & 3))
        p += bs if bt != 1 else 1
        if last:
            return lit, seq


@pytest.mark.skipif(not S.have_zstd, reason="libzstd.so.1 not present")
def test_treeless_literals_and_repeat_mode_tables():
    """Multi-block frames whose later blocks reuse the previous Huffman table (treeless literals) and the previous FSE tables
    (Repeat_Mode): the decode kernel keeps both kinds of table in ONE shared-memory region and rebuilds what a block reuses
    from its saved description (zstd_decode.cu), so these modes must be in the test set, not just happen to be."""
    datas = [corpus.text(1 << 20, 5), corpus.lz_model(1 << 20, 6), capi.synth_host(16, 65536, seed=0xC0FFEE).tobytes()]
    units, want, treeless, repeats = [], [], 0, 0
    for d in datas:
        for lvl in (1, 3, 9, 19):
            f = S.zstd_compress(d, level=lvl)
            lit, seq = _block_modes(f)
            treeless += lit.count(3)
            repeats += sum(s.count(3) for s in seq)
            units.append(f)
            want.append(d)
    assert treeless >= 10 and repeats >= 10, (treeless, repeats)
    outs, st = ctx().run_host_units(capi.ZSTD, False, units, [len(d) for d in want])
    assert (st == 0).all(), [(i, int(s)) for i, s in enumerate(st) if s]
    assert outs == want
