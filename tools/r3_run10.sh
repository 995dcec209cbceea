#!/bin/bash
mkdir -p gpurun_out
for g in 32 128; do
EARLY_L2_FETCH=$g SWEEP_ONLY=7:3 timeout 300 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum --clock-control none -k regex:g7_kernel -s 2 -c 1 --csv python tools/g7_sweep.py 65536 snappy > gpurun_out/l2f_$g.log 2>&1
grep -E "cudaDevice" gpurun_out/l2f_$g.log; grep "g7_kernel" gpurun_out/l2f_$g.log | sed 's/.*G7)",//' | cut -c1-200
done
