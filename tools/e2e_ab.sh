#!/bin/bash
# A/B of the pinned-arena pipeline's chunking (tools/e2e_ab.py), plus the tests of that path
mkdir -p gpurun_out
{
for cfg in "16 0" "16 1" "32 1" "24 1" "16 0" "16 1"; do set -- $cfg; CJ_PIPE_CHUNKS=$1 CJ_PIPE_RAMP=$2 timeout 300 python tools/e2e_ab.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_lz_decode.py tests/test_gpu_lz_encode.py -m gpu -q -k "pinned or host" 2>&1 | tail -2
} 2>&1 | tee gpurun_out/e2e_ab.log
