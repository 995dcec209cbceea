#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz_encode_kernel -s 1 -c 1 -f -o gpurun_out/r3_enc python tools/quick_bench.py 16384 snappy > gpurun_out/r3_enc.log 2>&1
tail -2 gpurun_out/r3_enc.log; ls -la gpurun_out/r3_enc.ncu-rep
