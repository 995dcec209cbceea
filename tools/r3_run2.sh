#!/bin/bash
# real-corpus sweep of the thread-per-block decode paths
mkdir -p gpurun_out
timeout 600 python tools/g7_sweep.py 65536 snappy lz4 --real 2>&1 | tail -10 | tee gpurun_out/r3_sweep_real.log
