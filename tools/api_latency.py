"""Latency / throughput of the cramjam-compatible Python API (host module) for single-buffer calls
(BASELINE.json configs[0]: snappy round trip on 1 MiB bytes).  usage: python tools/api_latency.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from cramjam_b200 import cramjam, _capi as capi

def timeit(f, reps):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    return (time.perf_counter() - t0) / reps, r

for mib in (1, 16, 256):
    n = mib << 20
    data = capi.synth_host(n // 65536, 65536).tobytes()
    reps = 20 if mib <= 16 else 3
    for name, mod in (("snappy", cramjam.snappy), ("lz4", cramjam.lz4), ("zstd", cramjam.zstd)):
        tc, c = timeit(lambda: mod.compress(data), reps)
        cb = bytes(c)
        td, d = timeit(lambda: mod.decompress(cb), reps)
        assert bytes(d) == data
        print(f"{name:6s} {mib:4d} MiB: compress {tc*1e3:8.2f} ms ({n/tc/1e9:6.2f} GB/s) ratio {n/len(cb):.2f} | decompress {td*1e3:8.2f} ms ({n/td/1e9:6.2f} GB/s)", flush=True)
    if mib == 1:
        tc, c = timeit(lambda: cramjam.snappy.compress_raw(data), reps)
        cb = bytes(c)
        td, d = timeit(lambda: cramjam.snappy.decompress_raw(cb), reps)
        print(f"snappy raw 1 MiB (one block): compress {tc*1e3:.2f} ms | decompress {td*1e3:.2f} ms")
