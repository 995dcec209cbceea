"""Offline model of the segment-parallel decoder (generation 3): a block's element stream is cut at
checkpoints every K elements, T threads each walk one segment serially moving <= B bytes per
iteration, a back-reference chunk waits until its source bytes were written in an EARLIER iteration.
Reports iterations per wave versus the dependency-free ideal (development aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle as O
from cramjam_b200 import _capi as capi
_ARGV = sys.argv[:]
from sim_rounds import parse_snappy  # noqa
sys.argv = _ARGV


def simulate(els, K, T, B):
    """els: (is_lit, len, off).  Returns (iterations, ideal_iterations, stall_lane_iterations, lane_iterations)."""
    n = len(els)
    o = np.zeros(n + 1, dtype=np.int64)
    for i, e in enumerate(els):
        o[i + 1] = o[i] + e[1]
    total = int(o[n])
    INF = 1 << 60
    wt = np.full(total, INF, dtype=np.int64)   # iteration at which each output byte was written
    segs = [(s, min(s + K, n)) for s in range(0, n, K)]
    it = 0
    ideal = 0
    stall = 0
    lane_it = 0
    for w0 in range(0, len(segs), T):
        wave = segs[w0:w0 + T]
        cur = [s for s, _ in wave]
        done = [0] * len(wave)
        chunks = [sum(-(-els[j][1] // B) for j in range(s, e)) for s, e in wave]
        ideal += max(chunks)
        active = len(wave)
        while active:
            it += 1
            writes = []
            for k, (s, e) in enumerate(wave):
                j = cur[k]
                if j >= e:
                    continue
                lane_it += 1
                is_lit, ln, off = els[j]
                d = done[k]
                c = min(B, ln - d)
                p = int(o[j]) + d
                if not is_lit:
                    c = min(c, off)
                    src = p - off
                    if wt[src:src + c].max() >= it:
                        stall += 1
                        continue
                writes.append((p, c))
                d += c
                if d == ln:
                    cur[k] = j + 1
                    done[k] = 0
                    if j + 1 >= e:
                        active -= 1
                else:
                    done[k] = d
            for p, c in writes:
                wt[p:p + c] = it
    return it, ideal, stall, lane_it


if __name__ == "__main__":
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    data = capi.synth_host(64, 65536)
    tot = [0, 0, 0, 0, 0]
    for blk in range(8, 20):
        cs = O.snappy_raw_compress(data[blk * 65536:(blk + 1) * 65536].tobytes())
        els = parse_snappy(cs)
        it, ideal, stall, lane_it = simulate(els, K, T, B)
        print(f"block {blk}: {len(els)} elements, iterations {it}, ideal {ideal}, stalled lane-iterations {stall}/{lane_it}")
        for i, v in enumerate((len(els), it, ideal, stall, lane_it)):
            tot[i] += v
    print(f"K={K} T={T} B={B}: iterations/element {tot[1]/tot[0]:.3f}  (ideal {tot[2]/tot[0]:.3f}); "
          f"x{tot[1]/tot[2]:.2f} over ideal; stalled {100*tot[3]/tot[4]:.1f}% of lane-iterations")
