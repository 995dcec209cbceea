#!/bin/bash
# compute-sanitizer over the generation-7 parity tests (hostile streams, capacities, unaligned units, long literals / matches)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_lz_decode4.py -x -q -m gpu -k "gen7 and not multi_round" > gpurun_out/r3_memcheck_g7.log 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r3_memcheck_g7.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/small_roundtrip.py > gpurun_out/r3_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard" gpurun_out/r3_racecheck.log | tail -3
# the zstd decoder after the shared Huffman / FSE table region and the one-window sequence reader: libzstd frames at every level,
# multi-block frames with treeless literals and Repeat_Mode tables, 750 mutated frames
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_zstd.py -x -q -m gpu -k "not synthetic_256k" > gpurun_out/r3_memcheck_zstd.log 2>&1; echo "memcheck zstd rc=$?"
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r3_memcheck_zstd.log | tail -3
