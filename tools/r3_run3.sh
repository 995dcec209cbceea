#!/bin/bash
mkdir -p gpurun_out
SWEEP_WARPS=7,10,12 timeout 600 python tools/g7_sweep.py 65536 snappy 2>&1 | tail -8 | tee gpurun_out/r3_sweep_warps.log
