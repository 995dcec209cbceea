#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_host_api.py -m gpu -x -q > gpurun_out/r2_tests6.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_tests6.log
python tools/zstd_enc_bench.py 4096 2>&1 | tail -4
