#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_tests5.log
CJ_DECODE_GEN=2 timeout 300 python bench.py --no-extras --steps 5 --warmup 3 > gpurun_out/r2_gen2_tma.log 2>&1
grep '^{' gpurun_out/r2_gen2_tma.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gen2 (TMA rings):', round(d['value'],1), 'GB/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],1))"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_lz_decode.py -m gpu -x -q -k "edge or hostile or capacity" > gpurun_out/r2_memcheck_gen2.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r2_memcheck_gen2.log
