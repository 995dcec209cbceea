#!/bin/bash
# usage: tools/gen_scaling.sh  -- kernel time of generations 2 and 4 alone at several batch sizes
for g in 2 4; do for n in 16384 32768 49152; do
  CJ_DECODE_GEN=$g timeout 200 python bench.py --no-extras --steps 3 --warmup 3 --blocks $n 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('gen $g blocks $n: value', round(d['value'],1), 'GB/s  ms/step', round(d['ms_per_step'],3))
"
done; done
