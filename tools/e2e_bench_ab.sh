#!/bin/bash
# the bench's own e2e leg under different pipeline chunkings (same box, back to back)
mkdir -p gpurun_out
for cfg in "16 0" "16 1" "32 1" "16 0" "32 1"; do set -- $cfg
  CJ_PIPE_CHUNKS=$1 CJ_PIPE_RAMP=$2 timeout 600 python bench.py --no-extras --steps 5 --warmup 3 --cpu-seconds 1 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read())
e=d['e2e']; print('chunks $1 ramp $2: e2e %.2f ms %.2f GB/s  link-only %.2f ms  value %.1f GB/s' % (e['ms_per_step'], e['value'], e['link_only_ms_per_step'], d['value']))"
done 2>&1 | tee gpurun_out/e2e_bench_ab.log
