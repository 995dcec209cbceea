"""zstd level-3 frame decompress timing (BASELINE configs[3] shape: N x 256 KiB frames made by libzstd)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import syslibs as S
from cramjam_b200 import _capi as capi
from concurrent.futures import ThreadPoolExecutor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
U = 262144
data = capi.synth_host(n * 4, 65536)
blocks = [data[i * U:(i + 1) * U].tobytes() for i in range(n)]
t0 = time.time()
with ThreadPoolExecutor(os.cpu_count()) as ex:
    frames = list(ex.map(lambda b: S.zstd_compress(b, 3), blocks))
print(f"libzstd level-3 compress {n*U/(time.time()-t0)/1e9:.2f} GB/s ({os.cpu_count()} threads), ratio {n*U/sum(map(len,frames)):.3f}", flush=True)
t0 = time.time()
with ThreadPoolExecutor(os.cpu_count()) as ex:
    back = list(ex.map(lambda f: S.zstd_decompress(f, U), frames))
print(f"libzstd decompress {n*U/(time.time()-t0)/1e9:.2f} GB/s ({os.cpu_count()} threads, ctypes)", flush=True)
lens = np.array([len(f) for f in frames], dtype=np.uint64)
off = np.zeros(n, dtype=np.uint64); off[1:] = np.cumsum((lens[:-1] + 15) & ~np.uint64(15))
src = np.zeros(int(off[-1] + lens[-1]) + 64, dtype=np.uint8)
for i, f in enumerate(frames):
    src[int(off[i]):int(off[i]) + len(f)] = np.frombuffer(f, dtype=np.uint8)
dev = torch.device("cuda:0")
i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
t_src = torch.from_numpy(src).to(dev); t_dst = torch.zeros(n * U, dtype=torch.uint8, device=dev)
t_so, t_sl = i64(off), i64(lens)
t_do, t_dc = i64(np.arange(n, dtype=np.uint64) * U), i64(np.full(n, U, np.uint64))
t_dl = torch.zeros(n, dtype=torch.int64, device=dev); t_st = torch.zeros(n, dtype=torch.int32, device=dev)
c = capi.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); c.set_stream(stream.cuda_stream)
for _ in range(2):
    c.decompress_batch(capi.ZSTD, capi.DEVICE, n, t_src, t_so, t_sl, t_dst, t_do, t_dc, t_dl, t_st)
torch.cuda.synchronize()
assert (t_st == 0).all() and torch.equal(t_dst.cpu(), torch.from_numpy(data[: n * U]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 3
e0.record()
for _ in range(K):
    c.decompress_batch(capi.ZSTD, capi.DEVICE, n, t_src, t_so, t_sl, t_dst, t_do, t_dc, t_dl, t_st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"GPU zstd decode {ms:.2f} ms/batch  {n*U/ms/1e6:.1f} GB/s uncompressed ({n} x 256 KiB frames)")
# GPU zstd compress of the same 256 KiB units
bound = capi.lib().cj_compress_bound(capi.ZSTD, U)
slot = (bound + 15) // 16 * 16
t_raw = torch.from_numpy(data[: n * U]).to(dev)
t_cmp = torch.zeros(n * slot, dtype=torch.uint8, device=dev)
t_co, t_cc = i64(np.arange(n, dtype=np.uint64) * slot), i64(np.full(n, slot, np.uint64))
t_cl = torch.zeros(n, dtype=torch.int64, device=dev)
for _ in range(2):
    c.compress_batch(capi.ZSTD, capi.DEVICE, n, t_raw, t_do, t_dc, t_cmp, t_co, t_cc, t_cl, t_st)
torch.cuda.synchronize()
assert (t_st == 0).all()
e0.record()
for _ in range(K):
    c.compress_batch(capi.ZSTD, capi.DEVICE, n, t_raw, t_do, t_dc, t_cmp, t_co, t_cc, t_cl, t_st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"GPU zstd compress {ms:.2f} ms/batch  {n*U/ms/1e6:.1f} GB/s uncompressed  ratio {n*U/float(t_cl.sum().item()):.3f}")
c.decompress_batch(capi.ZSTD, capi.DEVICE, n, t_cmp, t_co, t_cl, t_dst.zero_(), t_do, t_dc, t_dl, t_st)
torch.cuda.synchronize()
assert (t_st == 0).all() and torch.equal(t_dst, t_raw)
e0.record()
for _ in range(K):
    c.decompress_batch(capi.ZSTD, capi.DEVICE, n, t_cmp, t_co, t_cl, t_dst, t_do, t_dc, t_dl, t_st)
e1.record(); torch.cuda.synchronize()
print(f"GPU zstd decode of GPU-made frames {n*U/(e0.elapsed_time(e1)/K)/1e6:.1f} GB/s")
