import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np
import oracle as O
from cramjam_b200 import _capi as capi
from sim_ranges import elements

def sim_words(els, U, T, W, inorder_grab=True):
    # per-byte source position (-1 literal)
    src = np.full(U, -1, dtype=np.int64)
    for o, ln, off in els:
        if off:
            for i in range(ln): src[o + i] = o + i - off
    nw = (U + W - 1) // W
    # dependencies of word w: set of words containing src bytes (excluding itself -> self dependency means overlapping copy: treat via period, ignore)
    deps = []
    for w in range(nw):
        s = src[w * W:(w + 1) * W]; s = s[s >= 0]
        d = set((s // W).tolist()); d.discard(w)
        deps.append(d)
    done = np.zeros(nw, dtype=bool)
    lanes = [None] * T
    nxt = 0; iters = 0; ndone = 0; blocked = 0
    while ndone < nw:
        iters += 1
        fin = []
        for l in range(T):
            if lanes[l] is None:
                if nxt >= nw: continue
                lanes[l] = nxt; nxt += 1
            w = lanes[l]
            if all(done[d] for d in deps[w]):
                fin.append(w); lanes[l] = None
            else: blocked += 1
        for w in fin: done[w] = True
        ndone += len(fin)
    return iters, nw, blocked

def depth_stats(els, U):
    depth = np.zeros(U, dtype=np.int32)
    for o, ln, off in els:
        if off:
            for i in range(ln): depth[o + i] = depth[o + i - off] + (1 if i < off else 0)
    return depth.max(), depth.mean()

data = capi.synth_host(64, 65536)
blocks = []
for b in range(64):
    c = O.snappy_raw_compress(data[b * 65536:(b + 1) * 65536].tobytes())
    els, U = elements(c)
    if len(els) > 1000: blocks.append((els, U))
    if len(blocks) >= 4: break
for els, U in blocks: print("byte dataflow depth max/mean", depth_stats(els, U))
for T, W in ((128, 16), (256, 16), (512, 16), (1024,16), (512, 8), (256, 32)):
    ti = tw = tb = 0
    for els, U in blocks:
        it, w, bl = sim_words(els, U, T, W); ti += it; tw += w; tb += bl
    n = len(blocks)
    print(f"T={T} W={W}: iterations/block {ti/n:.0f} words {tw/n:.0f} efficiency {tw/(ti*T):.2f} blocked {tb/n:.0f}")
