#!/bin/bash
# usage: tools/g5_sweep.sh "shares" "depths" [gen] [blocks] -- bench value of the co-scheduled split (CJ_DECODE_GEN=5) for several gen-4 shares / chunk pipeline depths
GEN=${3:-5}; N=${4:-65536}
for d in ${2:-3}; do for s in ${1:-"25 40 50 60"}; do
  CJ_DECODE_GEN=$GEN CJ_G3_MIN_UNITS=1024 CJ_G4_D=$d CJ_G4_SHARE=$s timeout 200 python bench.py --no-extras --steps 3 --warmup 3 --blocks $N 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('gen $GEN blocks $N depth $d share $s: value', round(d['value'],1), 'GB/s  ms/step', round(d['ms_per_step'],3))
"
done; done
