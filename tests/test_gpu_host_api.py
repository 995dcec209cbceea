"""-m gpu: the cramjam-compatible Python surface (C++ host binding over the C ABI).  Mirrors what the
reference's tests pin for snappy / lz4 / zstd: tests/test_variants.py (round trips over dtypes and
container types, *_into matrices incl. empty input, raw / block entry points, the LZ4 block golden
vector, streaming classes, error types) and tests/test_integration.py (third-party fixtures, lz4 block
without prepended size)."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import corpus

pytestmark = pytest.mark.gpu
VARIANTS = ("snappy", "lz4", "zstd")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
settings.register_profile("gpu", deadline=None, max_examples=12, derandomize=True, suppress_health_check=list(HealthCheck))
settings.load_profile("gpu")


@pytest.fixture(scope="module")
def cj():
    import cramjam_b200
    return cramjam_b200.cramjam


# ---- tests/test_integration.py:32-50 ----
@pytest.mark.parametrize("variant,suffix", [("zstd", "zst"), ("lz4", "lz4"), ("snappy", "snappy")])
def test_third_party_fixtures(cj, variant, suffix):
    plaintext = open(os.path.join(G, "plaintext.txt"), "rb").read()
    data = open(os.path.join(G, f"plaintext.txt.{suffix}"), "rb").read()
    assert bytes(getattr(cj, variant).decompress(data)) == plaintext


# ---- tests/test_variants.py:49-97 ----
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.float64, np.int64, np.complex128])
def test_any_dtype_roundtrip(cj, variant, dtype):
    mod = getattr(cj, variant)
    rng = np.random.default_rng(1)
    for n in (0, 1, 17, 1000, 10000):
        arr = (rng.integers(0, 50, size=n)).astype(dtype)
        for a in (arr, arr.reshape(-1, 1), arr[: n - n % 4].reshape(-1, 4) if n >= 4 else arr):
            out = mod.decompress(mod.compress(a))
            assert bytes(out) == a.tobytes()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("is_bytearray", [False, True])
def test_simple(cj, variant, is_bytearray):
    mod = getattr(cj, variant)
    for raw in (b"some bytes here" * 10, corpus.text(100000, 1), corpus.random_bytes(70000, 2), b""):
        data = bytearray(raw) if is_bytearray else raw
        compressed = mod.compress(data)
        assert isinstance(compressed, cj.Buffer)
        assert compressed.read() != raw
        compressed.seek(0)
        d1 = mod.decompress(compressed, output_len=len(raw))
        assert isinstance(d1, cj.Buffer) and d1.read() == raw
        compressed.seek(0)
        assert bytes(mod.decompress(compressed)) == raw


@pytest.mark.parametrize("variant", VARIANTS)
def test_output_len_is_a_hint(cj, variant):
    # generic! builds vec![0; output_len] and writes through a growing Cursor (src/lib.rs:217-233): an undersized hint
    # still succeeds, an oversized one returns max(hint, produced) bytes, zero padded
    mod = getattr(cj, variant)
    raw = corpus.text(5000, 3)
    comp = bytes(mod.compress(raw))
    assert bytes(mod.decompress(comp, output_len=10)) == raw
    big = bytes(mod.decompress(comp, output_len=len(raw) + 100))
    assert big[:len(raw)] == raw and big[len(raw):] == b"\x00" * 100
    padded = bytes(mod.compress(raw, output_len=len(comp) + 50))
    assert padded[:len(comp)] == comp and padded[len(comp):] == b"\x00" * 50
    assert bytes(mod.compress(raw, output_len=3)) == comp


@pytest.mark.parametrize("variant", VARIANTS)
def test_raises(cj, variant):
    with pytest.raises(cj.DecompressionError):
        getattr(cj, variant).decompress(b"sknow")


def _containers(cj, tmp_path, raw, length, tag):
    f = cj.File(str(tmp_path / f"{tag}.bin"), truncate=True)
    return {"bytes": lambda: b"\x00" * length, "bytearray": lambda: bytearray(length), "numpy": lambda: np.zeros(length, np.uint8),
            "Buffer": lambda: cj.Buffer(), "File": lambda: f, "memoryview": lambda: memoryview(b"\x00" * length)}


def _as_input(cj, kind, data, tmp_path, tag):
    if kind == "bytes": return bytes(data)
    if kind == "bytearray": return bytearray(data)
    if kind == "numpy": return np.frombuffer(data, dtype=np.uint8)
    if kind == "memoryview": return memoryview(bytes(data))
    if kind == "Buffer":
        return cj.Buffer(data)
    f = cj.File(str(tmp_path / f"in_{tag}.bin"), truncate=True)
    f.write(data); f.seek(0)
    return f


def _read_out(cj, out, n):
    if isinstance(out, (cj.File, cj.Buffer)):
        out.seek(0)
        return out.read()[:n] if n is not None else out.read()
    return bytes(out)[:n] if n is not None else bytes(out)


KINDS = ("bytes", "bytearray", "numpy", "Buffer", "File", "memoryview")


# ---- tests/test_variants.py:100-244 ----
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("in_kind", KINDS)
@pytest.mark.parametrize("out_kind", KINDS)
def test_into_matrix(cj, tmp_path, variant, in_kind, out_kind):
    mod = getattr(cj, variant)
    for i, raw in enumerate((b"", b"a", corpus.text(5000, 3), corpus.lz_model(80000, 4))):
        if len(raw) == 1 and out_kind in ("bytes", "memoryview"):
            continue  # CPython shares one object per single-byte bytes value; writing through it would corrupt the interpreter
        compressed = bytes(mod.compress(raw))
        # compress_into: output sized exactly to the compressed length
        out = _containers(cj, tmp_path, raw, len(compressed), f"c{i}")[out_kind]()
        n = mod.compress_into(_as_input(cj, in_kind, raw, tmp_path, f"c{i}"), out)
        assert n == len(compressed)
        assert bytes(mod.decompress(_read_out(cj, out, n))) == raw
        # decompress_into: returns exactly len(raw)
        out = _containers(cj, tmp_path, raw, len(raw), f"d{i}")[out_kind]()
        n = mod.decompress_into(_as_input(cj, in_kind, compressed, tmp_path, f"d{i}"), out)
        assert n == len(raw)
        assert _read_out(cj, out, n) == raw


@pytest.mark.parametrize("variant", VARIANTS)
def test_into_too_small_is_an_error_not_a_truncation(cj, variant):
    mod = getattr(cj, variant)
    raw = corpus.text(5000, 5)
    c = bytes(mod.compress(raw))
    with pytest.raises(cj.DecompressionError):
        mod.decompress_into(c, bytearray(len(raw) - 1))
    with pytest.raises(cj.CompressionError):
        mod.compress_into(raw, bytearray(len(c) - 1))


# ---- tests/test_variants.py:247-289 ----
@given(data=st.binary(max_size=70000))
def test_snappy_raw_into(cj, data):
    compressed = cj.snappy.compress_raw(data)
    buf = np.zeros(cj.snappy.compress_raw_max_len(data), dtype=np.uint8)
    n = cj.snappy.compress_raw_into(data, buf)
    assert n == len(compressed)
    assert cj.snappy.decompress_raw_len(buf[:n].tobytes()) == len(data)
    out = np.zeros(len(data), dtype=np.uint8)
    m = cj.snappy.decompress_raw_into(buf[:n].tobytes(), out)
    assert m == len(data) and out[:m].tobytes() == data
    assert bytes(cj.snappy.decompress_raw(compressed)) == data


@given(data=st.binary(max_size=70000))
def test_lz4_block_into(cj, data):
    compressed = cj.lz4.compress_block(data)
    buf = np.zeros(cj.lz4.compress_block_bound(data), dtype=np.uint8)
    n = cj.lz4.compress_block_into(data, buf)
    assert n == len(compressed)
    assert bytes(compressed) == buf[:n].tobytes()                      # determinism across the two entry points (:281)
    out = np.zeros(len(data), dtype=np.uint8)
    m = cj.lz4.decompress_block_into(buf[:n].tobytes(), out)
    assert m == len(data) and out[:m].tobytes() == data


# ---- tests/test_variants.py:314-341 ----
@pytest.mark.parametrize("kw", [dict(mode="default", acceleration=1, compression=1, store_size=True),
                                dict(mode="fast", acceleration=2, compression=2, store_size=False),
                                dict(mode="high_compression", acceleration=3, compression=3, store_size=True),
                                dict(mode="default", acceleration=5, compression=4, store_size=False)])
def test_lz4_block_golden(cj, kw):
    data = b"howdy neighbor"
    assert bytes(cj.lz4.compress_block(data)) == b"\x0e\x00\x00\x00\xe0howdy neighbor"
    assert bytes(cj.lz4.compress_block(data, store_size=False)) == b"\xe0howdy neighbor"
    out = cj.lz4.decompress_block(cj.lz4.compress_block(data, **kw), output_len=len(data) if not kw["store_size"] else None)
    assert bytes(out) == data


# ---- tests/test_integration.py:70-102 ----
@given(data=st.binary(min_size=1, max_size=100000))
@pytest.mark.parametrize("set_output_len", (True, False))
def test_lz4_block_into_without_prepended_size(cj, data, set_output_len):
    compressed = cj.lz4.compress_block(data, store_size=False)
    output_len = len(data) if set_output_len else None
    with pytest.raises(cj.DecompressionError):
        cj.lz4.decompress_block_into(compressed, bytearray(0), output_len=output_len)
    with pytest.raises(cj.DecompressionError, match=f"output_len set to {len(data)}, but output is less"):
        cj.lz4.decompress_block_into(compressed, bytearray(0), output_len=len(data))
    out = bytearray(len(data))
    cj.lz4.decompress_block_into(compressed, out, output_len=output_len)
    assert bytes(out) == data
    out = bytearray(len(compressed) * 2 + 16)
    n = cj.lz4.decompress_block_into(compressed, out, output_len=output_len)
    assert bytes(out[:n]) == data


# ---- tests/test_variants.py:361-414 ----
@pytest.mark.parametrize("variant", VARIANTS)
@given(first=st.binary(max_size=3000), second=st.binary(max_size=3000))
def test_stream_compressor(cj, variant, first, second):
    mod = getattr(cj, variant)
    c = mod.Compressor()
    c.compress(first)
    out = bytes(c.flush())
    c.compress(second)
    out += bytes(c.flush())
    out += bytes(c.finish())
    assert bytes(mod.decompress(out)) == first + second
    assert bytes(c.finish()) == b""
    with pytest.raises(cj.CompressionError):
        c.compress(b"data")


@pytest.mark.parametrize("variant", VARIANTS)
def test_stream_decompressor(cj, variant):
    mod = getattr(cj, variant)
    d = mod.Decompressor()
    compressed = mod.compress(b"bytes")
    for _ in range(2):
        assert d.decompress(bytes(compressed)) == 5
    assert bytes(d.flush()) == b"bytesbytes"
    assert bytes(d.flush()) == b""
    d.decompress(bytes(compressed))
    assert bytes(d.finish()) == b"bytes"
    with pytest.raises(cj.DecompressionError):
        d.finish()


def test_doc_example_and_threads(cj):
    # README.md:96-97 / src/lib.rs:37-38
    out = cj.Buffer()
    assert cj.snappy.compress_into(np.frombuffer(b"some bytes here", dtype=np.uint8), out) == 33
    assert out.tell() == 33
    out.seek(0)
    dst = b"0" * 15
    assert cj.snappy.decompress_into(out, dst) == 15 and dst == b"some bytes here"
    # the reference releases the GIL around codec calls; concurrent callers must be safe
    from concurrent.futures import ThreadPoolExecutor
    blobs = [corpus.text(20000 + 1000 * i, i) for i in range(16)]
    with ThreadPoolExecutor(8) as ex:
        back = list(ex.map(lambda b: bytes(cj.lz4.decompress(cj.lz4.compress(b))), blobs))
    assert back == blobs


def test_zstd_frames_are_valid_for_libzstd(cj):
    import syslibs as S
    if not S.have_zstd:
        pytest.skip("no libzstd")
    for raw in (b"", b"abc", corpus.text(300000, 7)):
        assert S.zstd_decompress(bytes(cj.zstd.compress(raw)), len(raw)) == raw


# ---- the B200 staging path of Buffer and the batch entry points (not in the reference's surface) ----
def test_pinned_buffer_roundtrip(cj):
    raw = corpus.text(3_000_000, 11)
    src = cj.Buffer(raw, pinned=True)
    assert src.pinned and len(src) == len(raw)
    for variant in VARIANTS:
        mod = getattr(cj, variant)
        comp = cj.Buffer(pinned=True)
        n = mod.compress_into(src, comp)
        assert n == len(comp) and n < len(raw)
        back = cj.Buffer(pinned=True)
        back.reserve(len(raw))
        m = mod.decompress_into(comp, back)
        assert m == len(raw) and bytes(back) == raw
        # the same Buffers again (storage stays registered), and appending at the cursor
        m2 = mod.decompress_into(comp, back)
        assert m2 == len(raw) and len(back) == 2 * len(raw) and bytes(back)[len(raw):] == raw
        # a pinned output and a pageable input still work (staged path)
        out2 = cj.Buffer(pinned=True)
        assert mod.decompress_into(bytes(comp), out2) == len(raw) and bytes(out2) == raw


def test_device_arrays_cuda_array_interface(cj):
    import torch
    raw = corpus.lz_model(60000, 5)
    t_raw = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    for name, comp_into, decomp_into, bound in (
            ("snappy raw", cj.snappy.compress_raw_into, cj.snappy.decompress_raw_into, cj.snappy.compress_raw_max_len(raw)),
            ("lz4 block", lambda a, b: cj.lz4.compress_block_into(a, b, store_size=False), cj.lz4.decompress_block_into, cj.lz4.compress_block_bound(raw))):
        t_comp = torch.zeros(bound, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        n = comp_into(t_raw, t_comp)
        assert 0 < n < len(raw), name
        t_back = torch.zeros(len(raw), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        m = decomp_into(t_comp[:n], t_back)
        assert m == len(raw) and t_back.cpu().numpy().tobytes() == raw, name
    # zstd frames decode device to device as well
    zc = torch.frombuffer(bytearray(bytes(cj.zstd.compress(raw))), dtype=torch.uint8).cuda()
    t_back = torch.zeros(len(raw), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert cj.zstd.decompress_into(zc, t_back) == len(raw) and t_back.cpu().numpy().tobytes() == raw
    with pytest.raises(ValueError):
        cj.snappy.decompress_raw_into(t_raw, bytearray(10))


def test_batch_entry_points(cj):
    datas = [corpus.text(n, n) for n in (10, 1000, 70000, 300000)] + [corpus.random_bytes(5000, 1), b"", corpus.lz_model(65536, 2)]
    got = cj.snappy.decompress_raw_batch(cj.snappy.compress_raw_batch(datas))
    assert [bytes(b) for b in got] == datas and all(isinstance(b, cj.Buffer) for b in got)
    # units the engine made must equal what the single-buffer call makes (same kernels, same bytes)
    assert [bytes(b) for b in cj.snappy.compress_raw_batch(datas)] == [bytes(cj.snappy.compress_raw(d)) for d in datas]
    assert [bytes(b) for b in cj.lz4.decompress_block_batch(cj.lz4.compress_block_batch(datas))] == datas
    nonempty = [d for d in datas if d]
    blocks = cj.lz4.compress_block_batch(nonempty, store_size=False)
    assert [bytes(b) for b in cj.lz4.decompress_block_batch(blocks, output_lens=[len(d) for d in nonempty])] == nonempty
    for mod in (cj.snappy, cj.lz4, cj.zstd):
        assert [bytes(b) for b in mod.decompress_batch(mod.compress_batch(datas))] == datas
        assert [bytes(b) for b in mod.decompress_batch([mod.compress(d) for d in datas])] == datas
    with pytest.raises(cj.DecompressionError):
        cj.snappy.decompress_raw_batch([bytes(cj.snappy.compress_raw(b"abc")), b"sknow"])
