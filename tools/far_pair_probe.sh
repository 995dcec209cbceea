#!/bin/bash
# runs tools/far_pair_probe.cu (built into variants/) plain for the times and under ncu for the DRAM traffic per launch
mkdir -p gpurun_out
for spin in 20 60; do timeout 120 ./variants/far_pair_probe $spin; done > gpurun_out/far_pair_times.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,l1tex__t_sector_hit_rate.pct,lts__t_requests_srcunit_tex_op_read.sum \
  --clock-control none --csv --log-file gpurun_out/far_pair_ncu.csv ./variants/far_pair_probe 60 > gpurun_out/far_pair_ncu.log 2>&1
tail -25 gpurun_out/far_pair_times.log
