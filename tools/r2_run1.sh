#!/bin/bash
# round-2 first GPU call: parity suite, then the thread-per-block decoder with 64- and 32-byte L2 fetch granularity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests.log
tail -5 gpurun_out/r2_tests.log
for g in 64 32; do
  CJ_L2_FETCH=$g timeout 300 python bench.py --no-extras --steps 5 --warmup 3 > gpurun_out/r2_l2fetch_$g.log 2>&1
  grep '^{' gpurun_out/r2_l2fetch_$g.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('L2 fetch $g:', round(d['value'],1), 'GB/s', round(d['ms_per_step'],3), 'ms')"
done
CJ_L2_FETCH=32 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:g4_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_l2fetch32_dram.csv python bench.py --no-extras --steps 1 --warmup 3 > gpurun_out/r2_ncu32.log 2>&1
tail -3 gpurun_out/r2_l2fetch32_dram.csv
