"""Offline model of a dynamic lane state machine: 32 lanes claim elements in stream order, each lane
moves <= B bytes of its element per iteration; a back-reference chunk is ready when its source lies
below the frontier F (lowest start of any in-flight element) or inside the lane's own finished part.
Reports warp iterations per element (development aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle as O
from cramjam_b200 import _capi as capi
_ARGV = sys.argv[:]
from sim_rounds import parse_snappy  # noqa
sys.argv = _ARGV


def simulate(els, T, B, precise):
    n = len(els)
    o = [0] * (n + 1)
    for i, e in enumerate(els):
        o[i + 1] = o[i] + e[1]
    INF = 1 << 60
    wt = np.full(o[n], INF, dtype=np.int64)
    nxt = 0
    cur = [-1] * T
    done = [0] * T
    it = 0
    stall = lane_it = 0
    remaining = n
    while remaining:
        it += 1
        for k in range(T):
            if cur[k] < 0 and nxt < n:
                cur[k] = nxt; done[k] = 0; nxt += 1
        F = min((o[cur[k]] for k in range(T) if cur[k] >= 0), default=o[n])
        writes = []
        for k in range(T):
            j = cur[k]
            if j < 0: continue
            lane_it += 1
            is_lit, ln, off = els[j]
            d = done[k]; c = min(B, ln - d); p = o[j] + d
            if not is_lit:
                c = min(c, off); s = p - off
                if precise:
                    ok = wt[s:s + c].max() < it
                else:
                    ok = (s + c <= F) or (s + c <= p and (s >= o[j] or F == o[j]))
                if not ok:
                    stall += 1; continue
            writes.append((p, c))
            d += c
            if d == ln:
                cur[k] = -1; remaining -= 1
            else:
                done[k] = d
        for p, c in writes:
            wt[p:p + c] = it
    return it, stall, lane_it


if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    precise = len(sys.argv) > 3 and sys.argv[3] == "precise"
    data = capi.synth_host(64, 65536)
    tot = [0, 0, 0, 0, 0]
    for blk in range(8, 20):
        cs = O.snappy_raw_compress(data[blk * 65536:(blk + 1) * 65536].tobytes())
        els = [e for e in parse_snappy(cs) if e[1] <= 64]
        if len(els) < 1000: continue
        it, stall, lane_it = simulate(els, T, B, precise)
        nb = sum(e[1] for e in els)
        for i, v in enumerate((len(els), it, stall, lane_it, nb)):
            tot[i] += v
    print(f"T={T} B={B} {'precise' if precise else 'frontier'}: iterations/element {tot[1]/tot[0]:.4f}; bytes/iteration {tot[4]/tot[1]:.1f}; "
          f"stalled {100*tot[2]/tot[3]:.1f}% of lane-iterations; lane occupancy {100*tot[3]/(tot[1]*T):.0f}%")
