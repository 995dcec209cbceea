#!/bin/bash
# encoder throughput vs resident CTAs per SM (4 warps each): fewer warps keep more of the in-flight input blocks in L2
for c in 2 3 4 5 6 7; do
  CJ_ENC_CTAS=$c python tools/quick_bench.py 16384 snappy 2>&1 | grep "GPU compress" | sed "s/^/ctas=$c /"
done
