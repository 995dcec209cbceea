"""Pins the CPU oracle against every golden vector / known answer the reference's tests hold for
the snappy / lz4 / zstd path and against the checksum libraries."""
import os

import numpy as np
import pytest

import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rd(name):
    with open(os.path.join(G, name), "rb") as f:
        return f.read()


PLAINTEXT = _rd("plaintext.txt")


def test_plaintext_fixture_identity():
    import hashlib
    assert len(PLAINTEXT) == 857
    assert hashlib.md5(PLAINTEXT).hexdigest() == "9d57e6dec8ab65f9b9ff7bae22ae7aa4"


# reference: tests/test_integration.py:32-50 (rows zstd/zst, lz4/lz4, snappy/snappy)
def test_golden_snappy_framed():
    assert O.snappy_frame_decompress(_rd("plaintext.txt.snappy")) == PLAINTEXT


def test_golden_lz4_frame():
    assert O.lz4f_decompress(_rd("plaintext.txt.lz4")) == PLAINTEXT


def test_golden_zstd():
    assert O.zstd_decompress(_rd("plaintext.txt.zst")) == PLAINTEXT
    assert O.zstd_len(_rd("plaintext.txt.zst")) == 857


# reference: tests/test_variants.py:329-334 — the only bit-exact compress pin in the tree
def test_golden_lz4_block_howdy():
    assert O.lz4_block_compress(b"howdy neighbor") == b"\xe0howdy neighbor"
    assert O.lz4_block_decompress(b"\xe0howdy neighbor", 14) == b"howdy neighbor"


# reference: README.md:96-97 / src/lib.rs:37-38 — snappy.compress_into(15 bytes) returns 33
def test_doc_example_snappy_33_bytes():
    c = O.snappy_frame_compress(b"some bytes here")
    assert len(c) == 33
    assert c[:10] == b"\xff\x06\x00\x00sNaPpY" and c[10] == 0x01
    assert O.snappy_frame_decompress(c) == b"some bytes here"


# reference: tests/test_variants.py:93-97
@pytest.mark.parametrize("fn", [O.snappy_frame_decompress, O.lz4f_decompress, O.zstd_decompress])
def test_sknow_raises(fn):
    with pytest.raises(O.OracleError):
        fn(b"sknow", 100)


def test_checksum_known_answers():
    assert O.crc32c(b"123456789") == 0xE3069283
    snap = _rd("plaintext.txt.snappy")
    assert int.from_bytes(snap[14:18], "little") == O.crc32c_masked(PLAINTEXT) == 0x8A74E3AF
    lz = _rd("plaintext.txt.lz4")
    assert (O.xxh32(lz[4:6]) >> 8) & 0xFF == lz[6] == 0xA7
    assert int.from_bytes(lz[-4:], "little") == O.xxh32(PLAINTEXT) == 0x33DA7D79
    zs = _rd("plaintext.txt.zst")
    assert int.from_bytes(zs[-4:], "little") == O.xxh64(PLAINTEXT) & 0xFFFFFFFF == 0xEC1CADDE


def test_xxhash_module_agreement():
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(7)
    for n in list(range(0, 70)) + [100, 1000, 4097, 65536]:
        d = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        for seed in (0, 1, 0xDEADBEEF):
            assert O.xxh32(d, seed) == xxhash.xxh32(d, seed=seed).intdigest()
            assert O.xxh64(d, seed) == xxhash.xxh64(d, seed=seed).intdigest()
