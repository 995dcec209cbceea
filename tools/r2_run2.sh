#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_multiround.py snappy 3 > gpurun_out/r2_diag_snappy.log 2>&1; tail -45 gpurun_out/r2_diag_snappy.log
REP=16 timeout 300 python tools/diag_multiround.py snappy 2 > gpurun_out/r2_diag_snappy16.log 2>&1; grep "^\[" gpurun_out/r2_diag_snappy16.log
timeout 300 python tools/diag_multiround.py lz4 2 > gpurun_out/r2_diag_lz4.log 2>&1; grep "^\[" gpurun_out/r2_diag_lz4.log
timeout 900 python bench.py > gpurun_out/r2_bench2.log 2>&1; echo "bench rc=$?"; tail -c 6000 gpurun_out/r2_bench2.log
