"""Offline model of the generation-6 decoder: a CTA decodes one block with its 64 KiB output window in
shared memory; every lane owns one R-byte output range at a time (grabbed in order), walks the elements that cover it and
moves one piece of at most P bytes per iteration when the piece's source bytes have been published by their owners.
Reports iterations per block, lane efficiency and how often a lane is blocked, for a few (T, R, P).
usage: python tools/sim_ranges.py [blocks]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import oracle as O
from cramjam_b200 import _capi as capi


def elements(c):
    ip = 0
    while c[ip] & 0x80: ip += 1
    ip += 1
    els = []  # (out_pos, len, off) off == 0 -> literal
    o = 0
    while ip < len(c):
        t = c[ip]; ty = t & 3
        if ty == 0:
            ln = (t >> 2) + 1
            if ln > 60:
                nb = ln - 60
                ln = int.from_bytes(c[ip + 1:ip + 1 + nb], "little") + 1
                ip += 1 + nb + ln
            else:
                ip += 1 + ln
            off = 0
        elif ty == 1:
            ln = 4 + ((t >> 2) & 7); off = ((t >> 5) << 8) | c[ip + 1]; ip += 2
        else:
            ln = (t >> 2) + 1; off = c[ip + 1] | (c[ip + 2] << 8); ip += 3
        els.append((o, ln, off)); o += ln
    return els, o


def pieces_of_range(els, starts, lo, hi, P):
    """pieces (out_pos, n, src_end or -1) covering [lo, hi)"""
    out = []
    i = np.searchsorted(starts, lo, side="right") - 1
    while i < len(els) and els[i][0] < hi:
        o, ln, off = els[i]
        a = max(o, lo); b = min(o + ln, hi)
        while a < b:
            n = min(P, b - a)
            out.append((a, n, -1 if off == 0 else min(a - off + n, a)))
            a += n
        i += 1
    return out


def simulate(els, U, T, R, P):
    starts = np.array([e[0] for e in els])
    nr = (U + R - 1) // R
    prog = np.zeros(nr, dtype=np.int64)  # published position per range
    for k in range(nr): prog[k] = k * R
    nxt = 0
    lanes = [None] * T  # [range, pieces, idx]
    iters = 0; work = 0; blocked = 0
    done = 0
    while done < nr:
        iters += 1
        newprog = []
        for l in range(T):
            st = lanes[l]
            if st is None:
                if nxt >= nr: continue
                k = nxt; nxt += 1
                st = lanes[l] = [k, pieces_of_range(els, starts, k * R, min((k + 1) * R, U), P), 0]
            k, ps, i = st
            a, n, se = ps[i]
            ok = True
            if se >= 0:
                ke = (se - 1) // R
                if ke != k:
                    ok = prog[ke] >= se
                    ks = (se - n) // R if se - n >= 0 else 0
                    if ok and ks != ke: ok = prog[ks] >= (ks + 1) * R
                # own range: always ready (flush rule)
            if ok:
                work += 1
                newprog.append((k, a + n))
                st[2] += 1
                if st[2] == len(ps):
                    lanes[l] = None; done += 1
            else:
                blocked += 1
        for k, p in newprog: prog[k] = p
    return iters, work, blocked


if __name__ == "__main__":
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    data = capi.synth_host(64, 65536)
    blocks = []
    for b in range(64):
        c = O.snappy_raw_compress(data[b * 65536:(b + 1) * 65536].tobytes())
        els, U = elements(c)
        if len(els) > 1000: blocks.append((els, U))
        if len(blocks) >= nb: break
    for T, R, P in ((256, 128, 8), (512, 128, 8), (256, 256, 8), (512, 64, 8), (512, 128, 16), (1024, 64, 8)):
        ti = tw = tb = 0
        for els, U in blocks:
            it, w, bl = simulate(els, U, T, R, P)
            ti += it; tw += w; tb += bl
        n = len(blocks)
        print(f"T={T:4d} R={R:3d} P={P:2d}: iterations/block {ti/n:7.1f}  pieces/block {tw/n:7.0f}  lane efficiency {tw/(ti*T):.2f}"
              f"  blocked lane-iterations/block {tb/n:7.0f}  warp-iterations/block {ti/n*T/32:6.0f}")
