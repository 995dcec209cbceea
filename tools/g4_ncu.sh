#!/bin/bash
# usage: tools/g4_ncu.sh name   -- one ncu --set full capture of the generation-4 kernel at bench size -> gpurun_out/<name>.ncu-rep
CJ_DECODE_GEN=4 timeout 400 ncu --set full --clock-control none --import-source on -k regex:g4_kernel -s 2 -c 1 -f -o gpurun_out/$1 python bench.py --no-extras --steps 1 --warmup 1 > gpurun_out/$1.log 2>&1
ls -la gpurun_out/$1.ncu-rep
