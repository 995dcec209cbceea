"""Host binding, no GPU needed: Buffer / File object semantics of the reference
(src/io.rs; pinned upstream by tests/test_rust_io.py, tests/test_buffer_view.py,
tests/test_variants.py::test_dunders / test_buffer_cmp)."""
import gc

import numpy as np
import pytest


@pytest.fixture(scope="module")
def cj():
    from cramjam_b200 import build
    build.build_host()
    import cramjam_b200
    return cramjam_b200.cramjam


def test_module_surface(cj):
    assert isinstance(cj.__version__, str)
    for sub, names in (("snappy", ["compress", "decompress", "compress_into", "decompress_into", "compress_raw", "decompress_raw",
                                   "compress_raw_into", "decompress_raw_into", "compress_raw_max_len", "decompress_raw_len", "Compressor", "Decompressor"]),
                       ("lz4", ["compress", "decompress", "compress_into", "decompress_into", "compress_block", "decompress_block",
                                "compress_block_into", "decompress_block_into", "compress_block_bound", "Compressor", "Decompressor"]),
                       ("zstd", ["compress", "decompress", "compress_into", "decompress_into", "Compressor", "Decompressor"])):
        mod = getattr(cj, sub)
        assert all(hasattr(mod, n) for n in names), sub
    assert issubclass(cj.CompressionError, Exception) and issubclass(cj.DecompressionError, Exception)
    assert cj.snappy.compress_raw_max_len(b"x" * 65536) == 76490
    assert cj.lz4.compress_block_bound(b"x" * 65536) == 65809 + 4


@pytest.mark.parametrize("kind", ["file", "buffer"])
def test_file_and_buffer_cursor_api(cj, tmp_path, kind):
    obj = cj.File(str(tmp_path / "f.bin")) if kind == "file" else cj.Buffer()
    assert obj.write(b"bytes") == 5 and obj.tell() == 5
    assert obj.seek(0) == 0 and obj.read() == b"bytes"
    assert obj.seek(-1, 2) == 4 and obj.read() == b"s"
    assert obj.seek(-2, whence=1) == 3 and obj.read() == b"es"
    with pytest.raises(ValueError):
        obj.seek(1, 3)
    for out in (b"12345", bytearray(b"12345"), cj.File(str(tmp_path / "o.bin")), cj.Buffer(), np.zeros(5, np.uint8)):
        obj.seek(0)
        assert obj.readinto(out) == 5
        if isinstance(out, (cj.File, cj.Buffer)):
            out.seek(0)
            assert out.read() == b"bytes"
        else:
            assert bytes(out) == b"bytes"
    obj.set_len(2); obj.seek(0)
    assert obj.read() == b"by"
    obj.set_len(10); obj.seek(0)
    assert obj.read() == b"by" + b"\x00" * 8
    obj.truncate(); obj.seek(0)
    assert obj.read() == b"" and len(obj) == 0 and not obj
    assert obj.seekable()


def test_dunders_and_cmp(cj, tmp_path):
    for data in (b"", b"x", b"some bytes" * 10):
        b = cj.Buffer()
        f = cj.File(str(tmp_path / f"d{len(data)}.bin"))
        for o in (b, f):
            assert len(o) == 0 and bool(o) is False
            o.write(data)
            assert len(o) == len(data) and bool(o) is bool(len(data))
            assert f"len={len(data)}" in repr(o)
        assert f"path={tmp_path}" in repr(f)
    assert cj.Buffer() == cj.Buffer()
    assert cj.Buffer(b"some bytes") == cj.Buffer(b"some bytes")
    assert cj.Buffer(b"some bytes") != cj.Buffer(b"other bytes")
    assert b"me by" in cj.Buffer(b"some bytes") and b"zz" not in cj.Buffer(b"some bytes")
    assert bytes(cj.Buffer(b"abc")) == b"abc" and memoryview(cj.Buffer(b"abc")).readonly


@pytest.mark.parametrize("copy", (None, True, False))
def test_buffer_copy_or_view(cj, copy):
    data = bytearray(b"bytes")
    buf = cj.Buffer(data) if copy is None else cj.Buffer(data, copy=copy)
    buf.write(b"0")
    assert data == (b"0ytes" if copy is False else b"bytes")
    assert (buf.get_view_reference() is data) == (copy is False)


def test_view_bounds(cj):
    data = bytearray(b"bytes")
    buf = cj.Buffer(data, copy=False)
    with pytest.raises(OSError, match="Too much to write on view"):
        buf.write(b"0" * 6)
    assert data == b"bytes"
    for _ in range(5):
        buf.write(b"0")
    with pytest.raises(OSError, match="Too much to write on view"):
        buf.write(b"0")
    assert data == b"00000"
    ro = cj.Buffer(b"bytes", copy=False)
    for n in range(7):
        with pytest.raises(OSError, match="Cannot set length on unowned buffer"):
            ro.set_len(n)
    with pytest.raises(OSError, match="Cannot truncate unowned buffer"):
        ro.truncate()
    assert ro.read(10) == b"bytes"
    ro.seek(0)
    assert b"".join(ro.read(i) for i in range(10)) == b"bytes"
    for whence in (0, 1, 2):
        v = cj.Buffer(bytearray(b"bytes"), copy=False)
        v.seek(2, whence=0); v.seek(2, whence=1); v.seek(-2, whence=2); v.seek(0)
        with pytest.raises(OSError, match="Bad seek: cannot seek outside bounds of unowned buffer"):
            v.seek(10, whence=whence)
        v.write(b"0")


def test_view_keeps_its_target_alive_and_tracks_resizes(cj):
    def make():
        d = bytearray(b"bytes")
        b = cj.Buffer(d, copy=False)
        return b, b.get_view_reference_count()
    buf, n0 = make()
    gc.collect()
    assert 0 < buf.get_view_reference_count() < n0
    assert buf.read() == b"bytes"
    assert cj.Buffer(b"x").get_view_reference_count() is None

    data = cj.Buffer()
    data.write(b"bytes")
    view = cj.Buffer(data, copy=False)
    view.write(b"12345")
    with pytest.raises(IOError, match="Too much to write on view"):
        view.write(b"6")
    assert len(view) == 5
    data.write(b"s")
    assert len(view) == 6 and view.tell() == 5
    view.write(b"6")
    assert view.tell() == 6
    data.set_len(2)
    assert view.tell() == 2
    with pytest.raises(IOError, match="Too much to write on view"):
        view.write(b"6")
    view.seek(1); view.write(b"1")
    assert view.tell() == 2 and len(view) == 2
