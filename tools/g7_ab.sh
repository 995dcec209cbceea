#!/bin/bash
# A/B of generation-7 builds in variants/ (CJ_LIB_PATH) on host-encoded streams (like the bench headline's), then ncu counters
# of each (or of $NCU_LIST).  usage: bash tools/g7_ab.sh lib_a.so lib_b.so ...
mkdir -p gpurun_out
export SWEEP_CACHE=/tmp/g7cache SWEEP_ONLY=7:3
for v in "$@"; do
  echo "== $v"
  CJ_LIB_PATH=$PWD/variants/$v timeout 600 python tools/g7_sweep.py 65536 snappy lz4 --oracle 2>&1 | grep -E "gen 7|Error|error" 
done 2>&1 | tee gpurun_out/g7_ab.log
for v in ${NCU_LIST:-"$@"}; do
  CJ_LIB_PATH=$PWD/variants/$v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum,lts__t_requests_srcunit_tex.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct \
    --clock-control none -k regex:g7_kernel -s 2 -c 1 --csv --log-file gpurun_out/g7_ab_$v.csv python tools/g7_sweep.py 65536 snappy --oracle > /dev/null 2>&1
  python - "$v" <<'PY'
import csv, sys
v = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/g7_ab_{v}.csv")))
hi = [i for i, r in enumerate(rows) if "Metric Name" in r][0]
h = rows[hi]
print(v, {r[h.index("Metric Name")]: r[h.index("Metric Value")] for r in rows[hi + 1:] if len(r) > h.index("Metric Value")})
PY
done 2>&1 | tee -a gpurun_out/g7_ab.log
