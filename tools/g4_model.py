"""Offline numbers for planning the next step of the thread-per-block decoder: sub-iterations per block for
8- and 16-byte chunks, and the share of back-references a per-lane mirror of a given size would serve on chip.
usage: python tools/g4_model.py [blocks]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import oracle as O
from cramjam_b200 import _capi as capi

n, U = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 65536
data = capi.synth_host(n, U)
c8 = []; c16 = []; offs = []; clen = []
for i in range(n):
    c = O.snappy_raw_compress(data[i * U:(i + 1) * U].tobytes())
    ip = 0
    while c[ip] & 0x80: ip += 1
    ip += 1
    a8 = a16 = 0
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
        elif ty == 1:
            ln = 4 + ((t >> 2) & 7); offs.append(((t >> 5) << 8) | c[ip + 1]); clen.append(ln); ip += 2
        else:
            ln = (t >> 2) + 1; offs.append(c[ip + 1] | (c[ip + 2] << 8)); clen.append(ln); ip += 3
        a8 += (ln + 7) // 8; a16 += (ln + 15) // 16
    c8.append(a8); c16.append(a16)
c8 = np.array(c8); c16 = np.array(c16); offs = np.array(offs); clen = np.array(clen)
w8 = c8[: n // 32 * 32].reshape(-1, 32).max(axis=1); w16 = c16[: n // 32 * 32].reshape(-1, 32).max(axis=1)
print(f"sub-iterations per block, 8-byte chunks: mean {c8.mean():.0f}, per-warp max {w8.mean():.0f}; 16-byte chunks: mean {c16.mean():.0f}, per-warp max {w16.mean():.0f}")
print(f"copies per block {len(offs) / n:.0f}, mean length {clen.mean():.1f} B")
for m in (64, 128, 192, 256, 448, 512, 1024, 2048, 4096):
    print(f"  offset <= {m:5d}: {100 * (offs <= m).mean():5.1f} % of copies")
