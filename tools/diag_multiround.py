"""Diagnostic for the multi-round thread-per-block decode (tests/test_gpu_lz_decode4.py::test_multi_round_batch_configs4_size):
which blocks differ, where, and whether the set is stable across repetitions / kernel variants.
usage: python tools/diag_multiround.py [codec: snappy|lz4] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import oracle as O
from cramjam_b200 import _capi as capi

name = sys.argv[1] if len(sys.argv) > 1 else "snappy"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
codec = capi.SNAPPY_RAW if name == "snappy" else capi.LZ4_BLOCK
S_, U, REP = 4096, 65536, int(os.environ.get("REP", "32"))
n = S_ * REP
data = O.synth(S_, U, first_index=7000)
bound = (32 + U + U // 6) if codec == capi.SNAPPY_RAW else (U + U // 255 + 16)
slot = (bound + 15) // 16 * 16
src = np.zeros(S_ * slot + 64, dtype=np.uint8)
so1 = np.arange(S_, dtype=np.uint64) * np.uint64(U)
do1 = np.arange(S_, dtype=np.uint64) * np.uint64(slot)
lens = O.batch(O.SNAPPY_RAW if codec == capi.SNAPPY_RAW else O.LZ4_BLOCK, 1, data, so1, np.full(S_, U, dtype=np.uint64), src, do1,
               np.full(S_, slot, dtype=np.uint64), nthreads=16)[0]
dev = torch.device("cuda:0")
i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
t_src = torch.from_numpy(src).to(dev)
t_so, t_sl = i64(np.tile(do1, REP)), i64(np.tile(lens.astype(np.uint64), REP))
t_do, t_dc = i64(np.arange(n, dtype=np.uint64) * np.uint64(U)), i64(np.full(n, U, dtype=np.uint64))
t_dl = torch.zeros(n, dtype=torch.int64, device=dev); t_st = torch.full((n,), -99, dtype=torch.int32, device=dev)
t_dst = torch.zeros(n * U + 64, dtype=torch.uint8, device=dev)
want = torch.from_numpy(data).to(dev).view(S_, U)
c = capi.Context(0)
for gen in (4, 2):
    c.set_decode_path(gen, 1)
    for rep in range(reps):
        t_dst.fill_(0xEE); t_st.fill_(-99); torch.cuda.synchronize()
        c.decompress_batch(codec, capi.DEVICE, n, t_src, t_so, t_sl, t_dst, t_do, t_dc, t_dl, t_st)
        c.synchronize()
        got = t_dst[:n * U].view(REP, S_, U)
        bad = (got != want.unsqueeze(0)).any(dim=2)          # [REP, S_]
        idx = bad.nonzero().cpu().numpy()
        print(f"[{name} gen{gen} rep{rep}] status!=0: {int((t_st != 0).sum())}  redo: {c.last_redo_count()}  bad blocks: {len(idx)}", flush=True)
        for r, s in idx[:12]:
            d = (got[r, s] != want[s]).nonzero().flatten().cpu().numpy()
            g = int(r) * S_ + int(s)
            print(f"   global block {g} (replica {r}, sample {s}, lane {g % 32}, warp {g // 32}): {len(d)} bytes differ, first at {d[0]}, last at {d[-1]};"
                  f" got {got[r, s, d[0]:d[0]+8].cpu().numpy().tolist()} want {want[s, d[0]:d[0]+8].cpu().numpy().tolist()}", flush=True)
