#!/bin/bash
# last evidence run of the round: tools/r3_final.sh, then the headline kernel's counters and one full capture on the bench's own streams
bash tools/r3_final.sh
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:"g7_kernel" -c 4 --csv --log-file gpurun_out/r02c_g7_metrics.csv python bench.py --no-extras --steps 1 --warmup 3 --cpu-seconds 0.5 > gpurun_out/r02c_metrics_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"g7_kernel" -s 3 -c 1 -f -o gpurun_out/r02c_g7_full python bench.py --no-extras --steps 1 --warmup 3 --cpu-seconds 0.5 > gpurun_out/r02c_full_run.log 2>&1
ls -la gpurun_out/r02c*
