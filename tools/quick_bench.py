"""Quick device-resident decode timing (development tool; bench.py is the contract).
usage: python tools/quick_bench.py [n_blocks] [codec: lz4|snappy]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import oracle as O
from cramjam_b200 import _capi as capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
codecs = sys.argv[2:] or ["lz4", "snappy"]
U = 65536
nt = os.cpu_count()
data = capi.synth_host(n, U)
for name in codecs:
    codec = capi.LZ4_BLOCK if name == "lz4" else capi.SNAPPY_RAW
    ocodec = O.LZ4_BLOCK if name == "lz4" else O.SNAPPY_RAW
    bound = capi.lib().cj_compress_bound(codec, U)
    slot = (bound + 15) // 16 * 16
    comp = np.zeros(n * slot, dtype=np.uint8)
    so = np.arange(n, dtype=np.uint64) * U
    do = np.arange(n, dtype=np.uint64) * slot
    clen, sec = O.batch(ocodec, 1, data, so, np.full(n, U, np.uint64), comp, do, np.full(n, slot, np.uint64), nthreads=nt)
    assert (clen > 0).all()
    # pack tightly at 16 B alignment
    lens = clen.astype(np.uint64)
    po = np.zeros(n, dtype=np.uint64); po[1:] = np.cumsum((lens[:-1] + 15) & ~np.uint64(15))
    packed = np.zeros(int(po[-1] + lens[-1]) + 64, dtype=np.uint8)
    for i in range(n):
        packed[int(po[i]):int(po[i] + lens[i])] = comp[int(do[i]):int(do[i] + lens[i])]
    ratio = n * U / lens.sum()
    print(f"[{name}] cpu compress {n*U/sec/1e9:.2f} GB/s on {nt} threads, ratio {ratio:.3f}", flush=True)
    # CPU decode baseline (oracle, all threads)
    back = np.zeros(n * U, dtype=np.uint8)
    dl, sec = O.batch(ocodec, 0, packed, po, lens, back, so, np.full(n, U, np.uint64), nthreads=nt)
    assert (dl == U).all() and np.array_equal(back, data)
    print(f"[{name}] cpu decode (oracle) {n*U/sec/1e9:.2f} GB/s on {nt} threads", flush=True)
    dev = torch.device("cuda:0")
    i64 = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
    t_src = torch.from_numpy(packed).to(dev); t_dst = torch.zeros(n * U, dtype=torch.uint8, device=dev)
    t_so, t_sl, t_do, t_dc = i64(po), i64(lens), i64(so), i64(np.full(n, U, np.uint64))
    t_dl = torch.zeros(n, dtype=torch.int64, device=dev); t_st = torch.zeros(n, dtype=torch.int32, device=dev)
    c = capi.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    c.set_stream(stream.cuda_stream)
    for it in range(3):
        c.decompress_batch(codec, capi.DEVICE, n, t_src, t_so, t_sl, t_dst, t_do, t_dc, t_dl, t_st)
    torch.cuda.synchronize()
    assert (t_st == 0).all() and torch.equal(t_dst.cpu(), torch.from_numpy(data))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    K = 5
    ev[0].record()
    for it in range(K):
        c.decompress_batch(codec, capi.DEVICE, n, t_src, t_so, t_sl, t_dst, t_do, t_dc, t_dl, t_st)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / K
    gbs = n * U / ms / 1e6
    print(f"[{name}] GPU decode {ms:.3f} ms/batch  {gbs:.1f} GB/s uncompressed  achieved(in+out) {gbs*(1+1/ratio):.1f} GB/s  "
          f"= {gbs*(1+1/ratio)/6530.3*100:.2f}% of measured HBM peak", flush=True)
    # GPU compress of the same blocks (device resident)
    t_raw = torch.from_numpy(data).to(dev)
    t_cmp = torch.zeros(n * slot, dtype=torch.uint8, device=dev)
    t_co, t_cc = i64(do), i64(np.full(n, slot, np.uint64))
    t_cl = torch.zeros(n, dtype=torch.int64, device=dev)
    for it in range(2):
        c.compress_batch(codec, capi.DEVICE, n, t_raw, t_do, t_dc, t_cmp, t_co, t_cc, t_cl, t_st)
    torch.cuda.synchronize()
    assert (t_st == 0).all()
    ev[0].record()
    for it in range(K):
        c.compress_batch(codec, capi.DEVICE, n, t_raw, t_do, t_dc, t_cmp, t_co, t_cc, t_cl, t_st)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / K
    gratio = n * U / float(t_cl.sum().item())
    print(f"[{name}] GPU compress {ms:.3f} ms/batch  {n*U/ms/1e6:.1f} GB/s uncompressed  ratio {gratio:.3f} (cpu {ratio:.3f})", flush=True)
    # decode what the GPU encoder produced
    c.decompress_batch(codec, capi.DEVICE, n, t_cmp, t_co, t_cl, t_dst.zero_(), t_do, t_dc, t_dl, t_st)
    torch.cuda.synchronize()
    assert (t_st == 0).all() and torch.equal(t_dst, t_raw)
    c.close()
