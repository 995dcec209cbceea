#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/quick_bench.py 32768 snappy lz4 2>&1 | grep -E "GPU compress|GPU decode" | tee gpurun_out/r3_enc14.log
timeout 600 python tools/zstd_bench.py 4096 2>&1 | grep GPU | tee -a gpurun_out/r3_enc14.log
timeout 900 python -m pytest tests/test_gpu_lz_encode.py tests/test_gpu_zstd.py -m gpu -x -q 2>&1 | tail -2
