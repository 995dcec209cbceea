#!/bin/bash
# usage: tools/g4_times.sh [blocks]   -- per-kernel durations of the generation-4 path at bench size (ncu launch list) + a bench line
N=${1:-65536}
export CJ_DECODE_GEN=4
timeout 300 python bench.py --no-extras --steps 3 --warmup 3 --blocks $N > gpurun_out/g4_bench.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/g4_bench.log"):
    if l.startswith("{"):
        d=json.loads(l); print("bench value", round(d["value"],1), "GB/s  ms/step", round(d["ms_per_step"],3), " frac", round(d["roofline"]["frac"],4))
PY
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"g4_|lz_decode" -s 5 -c 6 --csv --log-file gpurun_out/g4_launches.csv python bench.py --no-extras --steps 2 --warmup 1 --blocks $N > gpurun_out/g4_q.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/g4_launches.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[1:]: print(r[ki][:50], float(r[vi])/1e6, "ms")
PY
