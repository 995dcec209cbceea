"""End-to-end decode timing through cj_decompress_batch(CJ_PINNED) (development tool; bench.py is the contract): the bench's
e2e leg alone, on GPU-encoded Snappy streams packed into a pinned arena, so that pipeline knobs (CJ_PIPE_CHUNKS, CJ_PIPE_RAMP)
can be A/B'd in separate processes.  usage: python tools/e2e_ab.py [n_blocks]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from cramjam_b200 import _capi as capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
U = 65536
dev = torch.device("cuda:0")
c = capi.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
c.set_stream(stream.cuda_stream)
i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
codec = capi.SNAPPY_RAW
slot = (capi.lib().cj_compress_bound(codec, U) + 15) // 16 * 16
raw = torch.empty(n * U, dtype=torch.uint8, device=dev)
c.synth_device(raw, n, U)
t_cmp = torch.zeros(n * slot + 64, dtype=torch.uint8, device=dev)
t_uo, t_ul = i64(np.arange(n, dtype=np.uint64) * U), i64(np.full(n, U, np.uint64))
t_co, t_cc = i64(np.arange(n, dtype=np.uint64) * slot), i64(np.full(n, slot, np.uint64))
t_cl = torch.zeros(n, dtype=torch.int64, device=dev)
t_st = torch.zeros(n, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
c.compress_batch(codec, capi.DEVICE, n, raw, t_uo, t_ul, t_cmp, t_co, t_cc, t_cl, t_st)
c.synchronize()
assert int((t_st != 0).sum()) == 0
clen = t_cl.cpu().numpy().astype(np.uint64)
coff = np.zeros(n, dtype=np.uint64)
coff[1:] = np.cumsum((clen[:-1] + np.uint64(15)) & ~np.uint64(15))
span = int(coff[-1] + clen[-1])
packed = torch.zeros(span + 64, dtype=torch.uint8, device=dev)
c.copy_units(n, t_cmp, t_co, t_cl, packed, i64(coff))
c.synchronize()
h_comp = torch.empty(span + 64, dtype=torch.uint8).pin_memory()
h_comp.copy_(packed)
del t_cmp, packed
h_out = torch.empty(n * U, dtype=torch.uint8).pin_memory()
h_dl, h_st, h_cap = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.int32), np.full(n, U, np.uint64)
raw_off = np.arange(n, dtype=np.uint64) * U


def step():
    c.decompress_batch(codec, capi.PINNED, n, h_comp, coff, clen, h_out, raw_off, h_cap, h_dl, h_st)


step()
assert int((h_st != 0).sum()) == 0 and torch.equal(h_out[: 64 * U], raw[: 64 * U].cpu()) and torch.equal(h_out[-64 * U:], raw[-64 * U:].cpu())
times = []
for _ in range(6):
    t0 = time.perf_counter()
    step()
    times.append(1e3 * (time.perf_counter() - t0))
if os.environ.get("E2E_LINK"):   # what plain copies of the same bytes do on this box: both directions at once, and each alone
    d_c = torch.empty(span + 64, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def timed(f):
        torch.cuda.synchronize(); f(); torch.cuda.synchronize()
        t0 = time.perf_counter(); f(); torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0)
    def both():
        with torch.cuda.stream(s1): d_c.copy_(h_comp, non_blocking=True)
        with torch.cuda.stream(s2): h_out.copy_(raw, non_blocking=True)
    def h2d():
        with torch.cuda.stream(s1): d_c.copy_(h_comp, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2): h_out.copy_(raw, non_blocking=True)
    def both_chunked(k=32):
        cs, co = (span + k - 1) // k, (n * U + k - 1) // k
        for i in range(k):
            with torch.cuda.stream(s1): d_c[i * cs:(i + 1) * cs].copy_(h_comp[i * cs:(i + 1) * cs], non_blocking=True)
            with torch.cuda.stream(s2): h_out[i * co:(i + 1) * co].copy_(raw[i * co:(i + 1) * co], non_blocking=True)
    def both_paced(k=32, lead=2):
        # input pieces issued just in time: piece i + lead goes out when output piece i is done (the input link idles in between)
        cs, co = (span + k - 1) // k, (n * U + k - 1) // k
        evs = []
        for i in range(min(lead, k)):
            with torch.cuda.stream(s1): d_c[i * cs:(i + 1) * cs].copy_(h_comp[i * cs:(i + 1) * cs], non_blocking=True)
        for i in range(k):
            with torch.cuda.stream(s2):
                h_out[i * co:(i + 1) * co].copy_(raw[i * co:(i + 1) * co], non_blocking=True)
                e = torch.cuda.Event(); e.record(s2); evs.append(e)
            j = i + lead
            if j < k:
                if i >= 1: evs[i - 1].synchronize()
                with torch.cuda.stream(s1): d_c[j * cs:(j + 1) * cs].copy_(h_comp[j * cs:(j + 1) * cs], non_blocking=True)
    def pieces(kh, kd):
        def f():
            cs, co = (span + kh - 1) // kh if kh else 0, (n * U + kd - 1) // kd if kd else 0
            for i in range(max(kh, kd)):
                if i < kh:
                    with torch.cuda.stream(s1): d_c[i * cs:(i + 1) * cs].copy_(h_comp[i * cs:(i + 1) * cs], non_blocking=True)
                if i < kd:
                    with torch.cuda.stream(s2): h_out[i * co:(i + 1) * co].copy_(raw[i * co:(i + 1) * co], non_blocking=True)
        return f
    print("link pieces (input x output -> ms): " + "  ".join(f"{a}x{b} {timed(pieces(a, b)):.2f}" for a, b in ((0, 1), (0, 32), (1, 0), (32, 0), (1, 32), (32, 1), (8, 8), (16, 16), (32, 32))), flush=True)
    print(f"link paced: lead 2 {timed(both_paced):.2f} ms  lead 4 {timed(lambda: both_paced(32, 4)):.2f} ms  lead 8 {timed(lambda: both_paced(32, 8)):.2f} ms", flush=True)
    print(f"link: both {timed(both):.2f} ms  both in 32 pieces {timed(both_chunked):.2f} ms  h2d alone {timed(h2d):.2f} ms ({span / 1e9:.2f} GB)  d2h alone {timed(d2h):.2f} ms ({n * U / 1e9:.2f} GB)", flush=True)
print(f"e2e CJ_PIPE_CHUNKS={os.environ.get('CJ_PIPE_CHUNKS', 'default')} CJ_PIPE_RAMP={os.environ.get('CJ_PIPE_RAMP', '1')}: "
      f"median {np.median(times):.2f} ms  best {min(times):.2f} ms  -> {n * U / np.median(times) / 1e6:.2f} GB/s uncompressed  "
      f"(h2d {span / 1e9:.2f} GB, d2h {n * U / 1e9:.2f} GB)", flush=True)
