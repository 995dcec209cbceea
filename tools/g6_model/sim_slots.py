import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/tests", "/root/repo/tools"]
import numpy as np
import oracle as O
from sim_ranges import elements, pieces_of_range

def elems_of_range(els, starts, lo, hi):
    out = []
    i = np.searchsorted(starts, lo, side="right") - 1
    while i < len(els) and els[i][0] < hi:
        o, ln, off = els[i]
        a = max(o, lo); b = min(o + ln, hi)
        out.append([a, b - a, off])
        i += 1
    return out

def sim(els, U, T, R, slots, polls_per_iter=1):
    starts = np.array([e[0] for e in els])
    nr = (U + R - 1) // R
    ready = np.zeros(U + 64, dtype=bool)
    lanes = [None] * T; nxt = 0; done = 0; iters = 0; work = 0; stall = 0
    while done < nr:
        iters += 1
        newly = []
        for l in range(T):
            st = lanes[l]
            if st is None:
                if nxt >= nr: continue
                k = nxt; nxt += 1
                st = lanes[l] = {"q": elems_of_range(els, starts, k * R, min((k + 1) * R, U)), "pend": [], "rr": 0}
            did = False
            # poll pending slots (round robin, limited number per iteration)
            for _ in range(min(polls_per_iter, len(st["pend"]))):
                i = st["rr"] % len(st["pend"]); st["rr"] += 1
                a, n, off = st["pend"][i]
                m = min(n, 8, off)
                if ready[a - off:a - off + m].all():
                    newly.append((a, m)); st["pend"][i][0] += m; st["pend"][i][1] -= m
                    if st["pend"][i][1] == 0: st["pend"].pop(i)
                    did = True; break
            if not did and st["q"]:
                a, n, off = st["q"][0]
                m = min(n, 8) if off == 0 else min(n, 8, off)
                if off == 0 or ready[a - off:a - off + m].all():
                    newly.append((a, m)); st["q"][0][0] += m; st["q"][0][1] -= m
                    if st["q"][0][1] == 0: st["q"].pop(0)
                    did = True
                elif len(st["pend"]) < slots:
                    st["pend"].append(st["q"].pop(0)); did = True   # parking costs the iteration
            if did: work += 1
            else: stall += 1
            if not st["q"] and not st["pend"]:
                lanes[l] = None; done += 1
        for a, m in newly: ready[a:a + m] = True
    return iters, work, stall

data = O.synth(64, 65536)
blocks = []
for b in range(64):
    c = O.snappy_raw_compress(data[b * 65536:(b + 1) * 65536].tobytes())
    els, U = elements(c)
    if len(els) > 1000: blocks.append((els, U))
    if len(blocks) >= 4: break
for T, R, slots, ppi in ((512, 128, 2, 1), (512, 128, 3, 1), (512, 128, 4, 1), (512, 128, 8, 1), (512,128,4,4), (256, 128, 4, 1), (512, 64, 3, 1)):
    ti = tw = ts = 0
    for els, U in blocks:
        it, w, s = sim(els, U, T, R, slots, ppi); ti += it; tw += w; ts += s
    n = len(blocks)
    print(f"T={T} R={R} slots={slots} polls/iter={ppi}: iterations/block {ti/n:.0f} work {tw/n:.0f} stall {ts/n:.0f} warp-iters {ti/n*T/32:.0f}")
