"""-m gpu tests: frame containers (snappy framing format, LZ4 frame) through the C ABI.
Decode must be byte-exact against the reference's golden fixtures (tests/test_integration.py:32-50),
the oracle and third-party encoders; encode must be decodable by the oracle and third parties."""
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


def bound(codec, b):
    out = C.c_size_t()
    rc = capi.lib().cj_decompress_bound(codec, b, len(b), C.byref(out))
    return rc, out.value


def dec(codec, units, caps=None):
    if caps is None:
        caps = []
        for u in units:
            rc, v = bound(codec, u)
            caps.append(v if rc == 0 else 16)
    return ctx().run_host_units(codec, False, units, caps)


def enc(codec, units):
    return ctx().run_host_units(codec, True, units, [capi.lib().cj_compress_bound(codec, len(u)) for u in units])


def test_golden_fixtures():
    outs, st = dec(capi.SNAPPY_FRAMED, [open(os.path.join(G, "plaintext.txt.snappy"), "rb").read()])
    assert st[0] == 0 and outs[0] == PLAINTEXT
    outs, st = dec(capi.LZ4_FRAME, [open(os.path.join(G, "plaintext.txt.lz4"), "rb").read()])
    assert st[0] == 0 and outs[0] == PLAINTEXT


def test_sknow_is_an_error():
    # reference tests/test_variants.py:93-97
    for codec in (capi.SNAPPY_FRAMED, capi.LZ4_FRAME):
        outs, st = dec(codec, [b"sknow"], [100])
        assert st[0] != 0 and outs[0] is None


def test_snappy_framed_decode_oracle_streams():
    units = [O.snappy_frame_compress(d) for d in CASES]
    outs, st = dec(capi.SNAPPY_FRAMED, units)
    assert (st == 0).all() and outs == CASES
    a, b = O.snappy_frame_compress(b"hello " * 100), O.snappy_frame_compress(corpus.text(200000, 1))
    cat = a + b"\xfe\x03\x00\x00abc" + b
    outs, st = dec(capi.SNAPPY_FRAMED, [cat])
    assert st[0] == 0 and outs[0] == b"hello " * 100 + corpus.text(200000, 1)


def test_snappy_framed_encode():
    outs, st = enc(capi.SNAPPY_FRAMED, CASES)
    assert (st == 0).all()
    for d, c in zip(CASES, outs):
        assert c[:10] == b"\xff\x06\x00\x00sNaPpY"
        assert O.snappy_frame_decompress(c) == d
    assert len(outs[CASES.index(b"some bytes here")]) == 33       # README.md:96-97 doc example
    assert outs[CASES.index(b"")] == b"\xff\x06\x00\x00sNaPpY"
    back, st = dec(capi.SNAPPY_FRAMED, outs)
    assert (st == 0).all() and back == CASES


def test_snappy_framed_checksum_and_capacity_errors():
    d = corpus.text(100000, 2)
    c = bytearray(O.snappy_frame_compress(d))
    bad = bytearray(c); bad[14] ^= 1                               # CRC word of the first chunk
    outs, st = dec(capi.SNAPPY_FRAMED, [bytes(bad), bytes(c), bytes(c)], [len(d), len(d), len(d) - 1])
    assert st[0] == 7 and st[1] == 0 and outs[1] == d and st[2] == 5
    trunc = bytes(c[:-5])
    outs, st = dec(capi.SNAPPY_FRAMED, [trunc], [len(d)])
    assert st[0] != 0


@pytest.mark.skipif(not S.have_lz4, reason="liblz4.so.1 not present")
def test_lz4_frame_decode_liblz4_frames():
    units, want = [], []
    for d in CASES:
        for kw in (dict(), dict(independent=True), dict(level=4), dict(block_checksum=True, independent=True),
                   dict(content_size=True), dict(content_checksum=False, block_size_id=5), dict(level=9, block_size_id=7)):
            units.append(S.lz4f_compress(d, **kw))
            want.append(d)
    outs, st = dec(capi.LZ4_FRAME, units)
    assert (st == 0).all(), st.nonzero()
    assert outs == want
    a, b = S.lz4f_compress(b"one " * 1000), S.lz4f_compress(b"two " * 1000)
    skip = (0x184D2A53).to_bytes(4, "little") + (5).to_bytes(4, "little") + b"skip!"
    outs, st = dec(capi.LZ4_FRAME, [a + skip + b])
    assert st[0] == 0 and outs[0] == b"one " * 1000 + b"two " * 1000


def test_lz4_frame_encode():
    outs, st = enc(capi.LZ4_FRAME, CASES)
    assert (st == 0).all()
    for d, c in zip(CASES, outs):
        assert c[:4] == b"\x04\x22\x4d\x18"
        assert O.lz4f_decompress(c) == d
        if d:
            import pyarrow as pa
            assert pa.decompress(c, decompressed_size=len(d), codec="lz4", asbytes=True) == d   # third-party LZ4F decoder
    back, st = dec(capi.LZ4_FRAME, outs)
    assert (st == 0).all() and back == CASES


@pytest.mark.skipif(not S.have_lz4, reason="liblz4.so.1 not present")
def test_lz4_frame_checksum_errors():
    d = corpus.text(50000, 3)
    f = bytearray(S.lz4f_compress(d))
    bad = bytearray(f); bad[-1] ^= 0x10                            # content checksum
    bad2 = bytearray(f); bad2[6] ^= 1                              # header checksum
    outs, st = dec(capi.LZ4_FRAME, [bytes(bad), bytes(bad2), bytes(f[:-9])], [len(d)] * 3)
    assert st[0] == 7 and st[1] == 7 and st[2] != 0


def test_decompress_bound_values():
    d = corpus.text(100000, 5)
    assert bound(capi.SNAPPY_RAW, O.snappy_raw_compress(d)) == (0, len(d))
    assert bound(capi.SNAPPY_FRAMED, O.snappy_frame_compress(d)) == (0, len(d))
    rc, v = bound(capi.LZ4_FRAME, O.lz4f_compress(d, 1 | 2 | 4))
    assert rc == 0 and v == len(d)
    rc, v = bound(capi.LZ4_FRAME, O.lz4f_compress(d, 1 | 2))       # no content size: an upper bound
    assert rc == 0 and v >= len(d)
    assert bound(capi.SNAPPY_FRAMED, b"sknow")[0] != 0


def test_lz4f_large_frames_block_parallel_path_matches_oracle():
    """Frames of >= 128 KiB with independent blocks take the block-parallel decode (frames.cu lz4f_decompress); linked
    frames, corrupted frames and short capacities must still give the oracle's bytes and status codes."""
    if not S.have_lz4:
        pytest.skip("no liblz4")
    data = capi.synth_host(20, 65536, seed=11, first_index=9).tobytes() + b"xyz" * 777
    rng = np.random.default_rng(17)
    units, caps = [], []
    for kw in (dict(independent=True), dict(independent=True, block_checksum=True), dict(independent=True, content_checksum=False),
               dict(independent=False), dict(independent=True, block_size_id=5), dict(independent=True, level=9)):
        f = S.lz4f_compress(data, **kw)
        units += [f, f, f + f, f[:-2]]
        caps += [len(data), len(data) - 1, 2 * len(data), len(data)]
        for _ in range(10):
            m = bytearray(f)
            m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
            units.append(bytes(m)); caps.append(len(data))
    assert_same_as_oracle(capi.LZ4_FRAME, units, caps, "host")
    # and the engine's own frames (independent blocks, content size + checksum)
    bound = capi.lib().cj_compress_bound(capi.LZ4_FRAME, len(data))
    enc, st = ctx().run_host_units(capi.LZ4_FRAME, True, [data], [bound])
    assert st[0] == 0
    outs, st = ctx().run_host_units(capi.LZ4_FRAME, False, enc, [len(data)])
    assert st[0] == 0 and outs[0] == data


@pytest.mark.parametrize("where", [capi.HOST, capi.PINNED])
def test_lz4_frame_large_host_hashed_content_checksum(where):
    """Frames of >= 1 MiB in host memory have their content checksum computed / verified on the host next to the GPU
    work (frames.cu HOST_HASH_MIN): the frame must still be the format's, and a corrupted checksum must still fail."""
    d = corpus.text(3_500_000, 9) + corpus.random_bytes(70_000, 1)
    outs, st = ctx().run_host_units(capi.LZ4_FRAME, True, [d, d[:1_200_000], b"small"], [capi.lib().cj_compress_bound(capi.LZ4_FRAME, len(d))] * 3, where=where)
    assert (st == 0).all()
    for c, want in zip(outs, (d, d[:1_200_000], b"small")):
        assert O.lz4f_decompress(c) == want                        # the oracle verifies header, block and content checksums
    back, st = ctx().run_host_units(capi.LZ4_FRAME, False, outs, [len(d), 1_200_000, 5], where=where)
    assert (st == 0).all() and back == [d, d[:1_200_000], b"small"]
    bad = bytearray(outs[0]); bad[-2] ^= 0x40                      # content checksum of the large frame
    _, st = ctx().run_host_units(capi.LZ4_FRAME, False, [bytes(bad), outs[1]], [len(d), 1_200_000], where=where)
    assert st[0] == 7 and st[1] == 0
    with pytest.raises(O.OracleError):
        O.lz4f_decompress(bytes(bad))
