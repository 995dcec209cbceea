#!/bin/bash
mkdir -p gpurun_out
for cap in 65536 4096 1024 256; do
SWEEP_ONLY=7:2 timeout 600 python tools/g7_sweep.py 65536 snappy --near=$cap 2>&1 | tail -1 | tee -a gpurun_out/r3_sweep_near.log
done
