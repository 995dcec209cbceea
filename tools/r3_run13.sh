#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zstd_decode_kernel -s 1 -c 1 -f -o gpurun_out/r3_zstd_dec python tools/zstd_bench.py 4096 > gpurun_out/r3_zstd_dec.log 2>&1
tail -3 gpurun_out/r3_zstd_dec.log; ls -la gpurun_out/r3_zstd_dec.ncu-rep
