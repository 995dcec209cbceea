#!/bin/bash
# A/B of builds in variants/ on the real corpus (tiled) and on oracle-encoded synthetic blocks.  usage: bash tools/g7_ab_real.sh lib_a.so lib_b.so
mkdir -p gpurun_out
export SWEEP_CACHE=/tmp/g7cache SWEEP_ONLY=7:3
for v in "$@"; do
  echo "== $v"
  CJ_LIB_PATH=$PWD/variants/$v timeout 600 python tools/g7_sweep.py 65536 snappy lz4 --real 2>&1 | grep -E "gen 7|rror"
  CJ_LIB_PATH=$PWD/variants/$v timeout 600 python tools/g7_sweep.py 65536 snappy lz4 --oracle 2>&1 | grep -E "gen 7|rror"
  CJ_LIB_PATH=$PWD/variants/$v timeout 600 python tools/g7_sweep.py 131072 snappy 2>&1 | grep -E "gen 7|rror"
done 2>&1 | tee gpurun_out/g7_ab_real.log
