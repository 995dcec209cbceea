"""Offline model of the execute phase: how many dependency rounds and copy calls a batch needs under
different readiness rules (development aid for lz_decode.cuh)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle as O
from cramjam_b200 import _capi as capi

def parse_snappy(cs):
    ip = 0
    while cs[ip] & 0x80: ip += 1
    ip += 1
    els = []  # (is_lit, length, off)
    n = len(cs)
    while ip < n:
        t = cs[ip]; ip += 1
        ty = t & 3
        if ty == 0:
            l = (t >> 2) + 1
            if l > 60:
                nb = l - 60; l = int.from_bytes(cs[ip:ip + nb], "little") + 1; ip += nb
            ip += l; els.append((True, l, 0))
        elif ty == 1:
            els.append((False, 4 + ((t >> 2) & 7), ((t >> 5) << 8) | cs[ip])); ip += 1
        elif ty == 2:
            els.append((False, 1 + (t >> 2), cs[ip] | (cs[ip + 1] << 8))); ip += 2
        else:
            els.append((False, 1 + (t >> 2), int.from_bytes(cs[ip:ip + 4], "little"))); ip += 4
    return els

if __name__ != "__main__": sys.argv = sys.argv[:1]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
data = capi.synth_host(64, 65536)
stats = {"prefix": [0, 0, 0], "precise": [0, 0, 0]}  # rounds, calls(32-job groups), jobs
nb = 0
for blk in range(8, 40):
    cs = O.snappy_raw_compress(data[blk * 65536:(blk + 1) * 65536].tobytes())
    els = [e for e in parse_snappy(cs) if e[1] <= 64]
    if len(els) < 1000: continue
    pos = 0; o = []
    for e in els:
        o.append(pos); pos += e[1]
    for b0 in range(0, len(els) - B, B):
        batch = list(range(b0, b0 + B)); Ob = o[b0]
        nb += 1
        for mode in ("prefix", "precise"):
            done = [els[j][0] or (o[j] - els[j][2] + min(els[j][1], els[j][2]) <= Ob) for j in batch]   # pass 1: literals + early copies
            rounds = 1; calls = -(-sum(done) // 32)
            while not all(done):
                rounds += 1
                newly = []
                f = done.index(False)
                for k, j in enumerate(batch):
                    if done[k]: continue
                    s = o[j] - els[j][2]; se = s + min(els[j][1], els[j][2])
                    if mode == "prefix":
                        ok = se <= o[batch[f]]
                    else:
                        ok = all(done[kk] for kk, jj in enumerate(batch[:k]) if o[jj] + els[jj][1] > s and o[jj] < se)
                    if ok: newly.append(k)
                for k in newly: done[k] = True
                calls += -(-len(newly) // 32)
            stats[mode][0] += rounds; stats[mode][1] += calls; stats[mode][2] += B
for mode, (r, c, j) in stats.items():
    print(f"B={B} {mode:8s}: rounds/batch {r/nb:.2f}  copy-calls/batch {c/nb:.2f}  calls per 32 elements {c/(j/32):.2f}")
