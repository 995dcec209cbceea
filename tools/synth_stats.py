"""Prints what the CPU encoders make of the synthetic corpus (ratio, elements per block) so the
generator in cramjam_b200/csrc/synth.cuh can be tuned to the Silesia aggregates of SURVEY.md §6/App. B."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle as O
import syslibs as S
from cramjam_b200 import _capi, build

build.build()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
data = _capi.synth_host(N, 65536)
tot_s = tot_l = 0
seqs = els = 0
per = []
for i in range(N):
    blk = data[i * 65536:(i + 1) * 65536].tobytes()
    cs = O.snappy_raw_compress(blk); cl = O.lz4_block_compress(blk)
    tot_s += len(cs); tot_l += len(cl)
    per.append(65536 / len(cl))
    # count lz4 sequences
    ip = 0; n = len(cl); k = 0
    while ip < n:
        t = cl[ip]; ip += 1; ll = t >> 4
        if ll == 15:
            while True:
                b = cl[ip]; ip += 1; ll += b
                if b != 255: break
        ip += ll
        k += 1
        if ip >= n: break
        ip += 2
        if (t & 15) == 15:
            while True:
                b = cl[ip]; ip += 1
                if b != 255: break
    seqs += k
    # count snappy elements
    ip = 0; n = len(cs)
    while cs[ip] & 0x80: ip += 1
    ip += 1; e = 0
    while ip < n:
        t = cs[ip]; ip += 1; e += 1
        ty = t & 3
        if ty == 0:
            l = (t >> 2) + 1
            if l > 60:
                nb = l - 60; l = int.from_bytes(cs[ip:ip + nb], "little") + 1; ip += nb
            ip += l
        else:
            ip += (1, 2, 4)[ty - 1]
    els += e
per = np.array(per)
print(f"blocks={N} snappy ratio={N*65536/tot_s:.3f} lz4 ratio={N*65536/tot_l:.3f}  lz4 seq/block={seqs/N:.0f} snappy elem/block={els/N:.0f}")
print("per-block lz4 ratio pct [5,25,50,75,95]:", np.percentile(per, [5, 25, 50, 75, 95]).round(2))
if S.have_zstd:
    M = N // 4
    tz = sum(len(S.zstd_compress(data[i * 262144:(i + 1) * 262144].tobytes(), 3)) for i in range(M))
    print(f"zstd-3 @256KiB ratio={M*262144/tz:.3f}")
