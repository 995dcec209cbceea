#!/bin/bash
# round 2, session 2, run 1: generation-7 parity tests, then the sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lz_decode4.py -m gpu -x -q > gpurun_out/r3_tests1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3_tests1.log
timeout 300 python tools/g7_sweep.py 65536 snappy lz4 2>&1 | tail -12 | tee gpurun_out/r3_sweep1.log
CJ_L2_FETCH=32 timeout 300 python tools/g7_sweep.py 65536 snappy 2>&1 | tail -4 | tee gpurun_out/r3_sweep1_l2_32.log
