"""Deterministic test inputs (numpy only; no reference files needed at run time)."""
import numpy as np

WORDS = (b"the quick brown fox jumps over the lazy dog lorem ipsum dolor sit amet consectetur adipiscing elit "
         b"sed do eiusmod tempor incididunt ut labore et dolore magna aliqua snappy lz4 zstd cramjam buffer "
         b"<row id=\"12345\" name=\"value\"/> {\"key\": [1, 2, 3], \"other\": null} 0123456789 ABCDEF\n").split(b" ")


def text(n, seed=0):
    rng = np.random.default_rng(seed)
    out = bytearray()
    idx = rng.zipf(1.3, size=n // 3 + 16) % len(WORDS)
    for i in idx:
        out += WORDS[i] + b" "
        if len(out) >= n:
            break
    return bytes(out[:n])


def random_bytes(n, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8).tobytes()


def lz_model(n, seed=0, lit_mean=6.0, match_mean=9.0, alphabet=64):
    """LZ77-style source: skewed literals alternating with back-references (incl. overlapping)."""
    rng = np.random.default_rng(seed)
    out = np.zeros(n, dtype=np.uint8)
    pos = 0
    while pos < n:
        ll = int(min(rng.geometric(1.0 / lit_mean), n - pos))
        out[pos:pos + ll] = (rng.integers(0, alphabet, size=ll) * rng.integers(0, alphabet, size=ll)) // alphabet
        pos += ll
        if pos >= n or pos < 4:
            continue
        ml = int(min(3 + rng.geometric(1.0 / match_mean), n - pos))
        off = int(1 + rng.integers(0, 1 << int(rng.integers(0, 16))) % pos)
        if off >= ml:
            out[pos:pos + ml] = out[pos - off:pos - off + ml]
        else:
            for i in range(ml):
                out[pos + i] = out[pos + i - off]
        pos += ml
    return out.tobytes()


def edge_cases():
    """Small and awkward inputs, as the reference's hypothesis tests generate (st.binary())."""
    cases = [b"", b"a", b"ab", b"abcd", b"some bytes here", b"howdy neighbor", b"a" * 13, b"a" * 17, b"ab" * 40,
             b"\x00" * 1000, b"abc" * 5000, bytes(range(256)) * 10, b"a" * 65536, b"a" * 65537, b"xyz" * 50000]
    cases += [random_bytes(n, n) for n in (1, 5, 12, 13, 16, 17, 64, 100, 1000, 4096, 65535, 65536, 65537, 70000, 200000)]
    cases += [text(n, n) for n in (20, 100, 1000, 10000, 65536, 100000, 300000)]
    cases += [lz_model(n, n) for n in (50, 500, 5000, 65536, 150000)]
    return cases
