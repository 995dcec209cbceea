#!/bin/bash
# usage: tools/g7_ncu.sh name [codec]  -- one ncu --set full capture of the generation-7 kernel at bench size -> gpurun_out/<name>.ncu-rep
mkdir -p gpurun_out
SWEEP_ONLY=${G7D:-7:3} timeout 600 ncu --set full --clock-control none --import-source on -k regex:g7_kernel -s 2 -c 1 -f -o gpurun_out/$1 python tools/g7_sweep.py 65536 ${2:-snappy} > gpurun_out/$1.log 2>&1
ls -la gpurun_out/$1.ncu-rep; tail -3 gpurun_out/$1.log
