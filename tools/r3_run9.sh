#!/bin/bash
mkdir -p gpurun_out
echo "OUT_G=8 (NEAR 96)"; SWEEP_ONLY=7:2 timeout 300 python tools/g7_sweep.py 65536 snappy lz4 2>&1 | tail -2
echo "OUT_G=16 (NEAR 224)"; CJ_LIB_PATH=$PWD/variants/lib_out16.so SWEEP_ONLY=7:2 timeout 300 python tools/g7_sweep.py 65536 snappy lz4 2>&1 | tail -2
