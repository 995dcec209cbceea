#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_frames.py tests/test_gpu_host_api.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/zstd_bench.py 4096 2>&1 | grep GPU
timeout 600 python tools/zstd_bench.py 16384 2>&1 | grep "GPU zstd decode"
