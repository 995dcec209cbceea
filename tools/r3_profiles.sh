#!/bin/bash
# round-2 (second session) evidence run: bench line, reference arm, ncu launch list, per-kernel DRAM traffic, one full capture of the headline kernel (generation 7)
mkdir -p gpurun_out
echo "(GPU suite: see gpurun_out/r3_tests7.log)"
timeout 900 python bench.py > gpurun_out/r02b_bench.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/r02b_bench.log > gpurun_out/r02b_bench_line.json
timeout 900 python bench.py --impl reference > gpurun_out/r02b_bench_reference.log 2>&1; grep '^{' gpurun_out/r02b_bench_reference.log > gpurun_out/r02b_bench_reference_line.json
# every launch of a short headline run with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_bench_launches.csv python bench.py --no-extras --steps 2 --warmup 3 > gpurun_out/r02b_launches_run.log 2>&1
# per-kernel traffic and issue numbers at the bench sizes (the extras launch every codec kernel)
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 1500 ncu --metrics $M --clock-control none -k regex:"g7_kernel|g4_kernel|lz_encode_kernel|zstd_decode_kernel|zstd_encode_kernel|lz_decode_kernel" -c 120 --csv --log-file gpurun_out/r02b_kernels_metrics.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 0.5 > gpurun_out/r02b_metrics_run.log 2>&1
# the headline kernel, full set with source
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"g7_kernel" -s 3 -c 1 -f -o gpurun_out/r02b_g7_full python bench.py --no-extras --steps 1 --warmup 3 > gpurun_out/r02b_full_run.log 2>&1
ls -la gpurun_out | tail -12
