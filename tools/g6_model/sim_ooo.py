import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np
import oracle as O
from cramjam_b200 import _capi as capi
from sim_ranges import elements, pieces_of_range

def sim(els, U, T, R, P, lookahead=64):
    starts = np.array([e[0] for e in els])
    nr = (U + R - 1) // R
    ready = np.zeros(U + 64, dtype=bool)
    nxt = 0; lanes = [None] * T; iters = 0; done = 0; work = 0; idle = 0
    while done < nr:
        iters += 1
        newly = []
        for l in range(T):
            st = lanes[l]
            if st is None:
                if nxt >= nr: continue
                k = nxt; nxt += 1
                st = lanes[l] = [k, pieces_of_range(els, starts, k * R, min((k + 1) * R, U), P)]
            ps = st[1]
            hit = -1
            for i, (a, n, se) in enumerate(ps[:lookahead]):
                if se < 0 or ready[se - n:se].all(): hit = i; break
                # own earlier pending pieces may be the source -> not ready, fine
            if hit >= 0:
                a, n, se = ps.pop(hit); newly.append((a, n)); work += 1
                if not ps: lanes[l] = None; done += 1
            else: idle += 1
        for a, n in newly: ready[a:a + n] = True
    return iters, work, idle

data = capi.synth_host(64, 65536)
blocks = []
for b in range(64):
    c = O.snappy_raw_compress(data[b * 65536:(b + 1) * 65536].tobytes())
    els, U = elements(c)
    if len(els) > 1000: blocks.append((els, U))
    if len(blocks) >= 4: break
for T, R, P, LA in ((256, 128, 8, 64), (512, 128, 8, 64), (512,128,8,4), (512, 128, 64, 64), (256, 256, 8, 64), (512,64,8,64)):
    ti = tw = tb = 0
    for els, U in blocks:
        it, w, bl = sim(els, U, T, R, P, LA); ti += it; tw += w; tb += bl
    n = len(blocks)
    print(f"T={T} R={R} P={P} LA={LA}: iterations/block {ti/n:.0f} pieces {tw/n:.0f} efficiency {tw/(ti*T):.2f} idle {tb/n:.0f} warp-iters {ti/n*T/32:.0f}")
