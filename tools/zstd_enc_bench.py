"""zstd encoder timing and ratio, device resident (N x 256 KiB frames of the synthetic corpus), per level.
usage: python tools/zstd_enc_bench.py [frames]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from cramjam_b200 import _capi as capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
U = 262144
dev = torch.device("cuda:0")
c = capi.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); c.set_stream(stream.cuda_stream)
raw = torch.empty(n * U, dtype=torch.uint8, device=dev)
c.synth_device(raw, n * 4, 65536)
slot = (capi.lib().cj_compress_bound(capi.ZSTD, U) + 15) // 16 * 16
i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
so, sl = i64(np.arange(n, dtype=np.uint64) * U), i64(np.full(n, U, np.uint64))
do, dc = i64(np.arange(n, dtype=np.uint64) * slot), i64(np.full(n, slot, np.uint64))
out = torch.empty(n * slot, dtype=torch.uint8, device=dev)
dl = torch.zeros(n, dtype=torch.int64, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
back = torch.empty(n * U, dtype=torch.uint8, device=dev)
for level in (1, 3):
    for _ in range(2):
        c.compress_batch(capi.ZSTD, capi.DEVICE, n, raw, so, sl, out, do, dc, dl, st, level=level)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(3):
        c.compress_batch(capi.ZSTD, capi.DEVICE, n, raw, so, sl, out, do, dc, dl, st, level=level)
    b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    assert int((st != 0).sum()) == 0
    ratio = n * U / float(dl.sum().item())
    c.decompress_batch(capi.ZSTD, capi.DEVICE, n, out, do, dl, back, so, sl, torch.zeros_like(dl), st)
    c.synchronize()
    assert int((st != 0).sum()) == 0 and bool(torch.equal(back, raw))
    print(f"zstd compress level {level}: {n * U / ms / 1e6:.1f} GB/s  ratio {ratio:.3f}  (round trip through the GPU decoder exact)", flush=True)
