"""Tiny end-to-end exercise of every kernel (used under compute-sanitizer racecheck / memcheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle as O
from cramjam_b200 import _capi as capi
c = capi.Context(0)
n, U = 24, 65536
data = capi.synth_host(n, U, seed=11)
blocks = [data[i * U:(i + 1) * U].tobytes() for i in range(n)]
for codec in (capi.SNAPPY_RAW, capi.LZ4_BLOCK, capi.SNAPPY_FRAMED, capi.LZ4_FRAME, capi.ZSTD):
    enc, st = c.run_host_units(codec, True, blocks, [capi.lib().cj_compress_bound(codec, U)] * n)
    assert (st == 0).all()
    dec, st = c.run_host_units(codec, False, enc, [U] * n)
    assert (st == 0).all() and dec == blocks, codec
    print("codec", codec, "ok, ratio", round(n * U / sum(map(len, enc)), 3))
c.close()
