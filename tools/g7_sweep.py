"""Device-resident decode timing of the thread-per-block paths (generations 4 and 7) over knob settings, one process
(development tool; bench.py is the contract).  usage: python tools/g7_sweep.py [n_blocks] [snappy|lz4 ...]
Streams are made by this engine's GPU encoder; every variant's output is compared with the generator's bytes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from cramjam_b200 import _capi as capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
codecs = sys.argv[2:] or ["snappy", "lz4"]
U = 65536
dev = torch.device("cuda:0")
c = capi.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
c.set_stream(stream.cuda_stream)
data = torch.from_numpy(capi.synth_host(n, U)).to(dev)
i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
for name in codecs:
    codec = capi.LZ4_BLOCK if name == "lz4" else capi.SNAPPY_RAW
    slot = (capi.lib().cj_compress_bound(codec, U) + 15) // 16 * 16
    t_cmp = torch.zeros(n * slot + 64, dtype=torch.uint8, device=dev)
    t_uo, t_ul = i64(np.arange(n, dtype=np.uint64) * U), i64(np.full(n, U, np.uint64))
    t_co, t_cc = i64(np.arange(n, dtype=np.uint64) * slot), i64(np.full(n, slot, np.uint64))
    t_cl = torch.zeros(n, dtype=torch.int64, device=dev)
    t_st = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    c.compress_batch(codec, capi.DEVICE, n, data, t_uo, t_ul, t_cmp, t_co, t_cc, t_cl, t_st)
    c.synchronize()
    assert int((t_st != 0).sum()) == 0
    ratio = n * U / float(t_cl.sum().item())
    t_dst = torch.zeros(n * U + 64, dtype=torch.uint8, device=dev)
    t_dl = torch.zeros(n, dtype=torch.int64, device=dev)
    variants = [(4, {}), (7, {"CJ_G7_D": "3"}), (7, {"CJ_G7_D": "2"})]
    if os.environ.get("SWEEP_ONLY"):   # e.g. SWEEP_ONLY=7:3 -> generation 7 with CJ_G7_D=3 only
        g_, d_ = os.environ["SWEEP_ONLY"].split(":")
        variants = [(int(g_), {"CJ_G7_D": d_})]
    for w in os.environ.get("SWEEP_WARPS", "").split(","):
        if w:
            variants.append((7, {"CJ_G7_D": "3", "CJ_G7_WARPS": w}))
    for gen, env in variants:
        for k in ("CJ_G7_D", "CJ_G7_WARPS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        c.set_decode_path(gen, 1)
        t_dst.zero_()
        torch.cuda.synchronize()
        for _ in range(2):
            c.decompress_batch(codec, capi.DEVICE, n, t_cmp, t_co, t_cl, t_dst, t_uo, t_ul, t_dl, t_st)
        c.synchronize()
        ok = int((t_st != 0).sum()) == 0 and torch.equal(t_dst[:n * U], data)
        redo = c.last_redo_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        K = 5
        ev[0].record(stream)
        for _ in range(K):
            c.decompress_batch(codec, capi.DEVICE, n, t_cmp, t_co, t_cl, t_dst, t_uo, t_ul, t_dl, t_st)
        ev[1].record(stream)
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / K
        gbs = n * U / ms / 1e6
        print(f"[{name}] gen {gen} {env}: {ms:.3f} ms  {gbs:.1f} GB/s uncompressed  frac {gbs * (1 + 1 / ratio) / 6537:.4f}  "
              f"bit-exact {ok}  redo {redo}  ratio {ratio:.3f}", flush=True)
