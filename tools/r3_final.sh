#!/bin/bash
# final evidence of the session: full GPU suite, bench line, reference arm, launch list (the per-kernel metric pass and the full capture are tools/r3_profiles.sh)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02b_gputests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02b_gputests.log
timeout 900 python bench.py > gpurun_out/r02b_bench.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/r02b_bench.log > gpurun_out/r02b_bench_line.json
timeout 900 python bench.py --impl reference > gpurun_out/r02b_bench_reference.log 2>&1; grep '^{' gpurun_out/r02b_bench_reference.log > gpurun_out/r02b_bench_reference_line.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_bench_launches.csv python bench.py --no-extras --steps 2 --warmup 3 > gpurun_out/r02b_launches_run.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
