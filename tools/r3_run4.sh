#!/bin/bash
mkdir -p gpurun_out
for g in 32 128; do
CJ_L2_FETCH=$g SWEEP_ONLY=7:2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:g7_kernel -s 2 -c 1 --csv python tools/g7_sweep.py 65536 snappy 2>&1 | grep -E "g7_kernel" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | sed "s/^/L2_FETCH=$g /"
done
