#!/bin/bash
# generation 2 vs 7 at smaller batches (where does the thread-per-block path start to pay?)
mkdir -p gpurun_out
for n in 4096 8192 16384 24576 32768; do
python - $n <<'PY' 2>&1 | tee -a gpurun_out/r3_scaling.log
import os, sys
import numpy as np
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "tests")]
import torch
from cramjam_b200 import _capi as capi
n = int(sys.argv[1]); U = 65536
dev = torch.device("cuda:0"); c = capi.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); c.set_stream(stream.cuda_stream)
data = torch.from_numpy(capi.synth_host(n, U)).to(dev)
i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
for name, codec in (("snappy", capi.SNAPPY_RAW), ("lz4", capi.LZ4_BLOCK)):
    slot = (capi.lib().cj_compress_bound(codec, U) + 15) // 16 * 16
    t_cmp = torch.zeros(n * slot + 64, dtype=torch.uint8, device=dev)
    t_uo, t_ul = i64(np.arange(n, dtype=np.uint64) * U), i64(np.full(n, U, np.uint64))
    t_co, t_cc = i64(np.arange(n, dtype=np.uint64) * slot), i64(np.full(n, slot, np.uint64))
    t_cl = torch.zeros(n, dtype=torch.int64, device=dev); t_st = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    c.compress_batch(codec, capi.DEVICE, n, data, t_uo, t_ul, t_cmp, t_co, t_cc, t_cl, t_st); c.synchronize()
    t_dst = torch.zeros(n * U + 64, dtype=torch.uint8, device=dev); t_dl = torch.zeros(n, dtype=torch.int64, device=dev)
    out = []
    for gen in (2, 7):
        c.set_decode_path(gen, 1)
        for _ in range(2): c.decompress_batch(codec, capi.DEVICE, n, t_cmp, t_co, t_cl, t_dst, t_uo, t_ul, t_dl, t_st)
        c.synchronize()
        assert int((t_st != 0).sum()) == 0 and torch.equal(t_dst[:n * U], data)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(stream)
        for _ in range(5): c.decompress_batch(codec, capi.DEVICE, n, t_cmp, t_co, t_cl, t_dst, t_uo, t_ul, t_dl, t_st)
        ev[1].record(stream); torch.cuda.synchronize()
        out.append(ev[0].elapsed_time(ev[1]) / 5)
    print(f"{name} n={n}: gen2 {out[0]:.3f} ms ({n*U/out[0]/1e6:.0f} GB/s)  gen7 {out[1]:.3f} ms ({n*U/out[1]/1e6:.0f} GB/s)", flush=True)
PY
done
