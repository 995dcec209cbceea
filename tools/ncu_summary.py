"""Summarise an .ncu-rep (first kernel) into the handful of metrics the roofline discussion needs.
usage: python tools/ncu_summary.py report.ncu-rep [out.csv] [title]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio'] + [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h] 
out = [f"# {sys.argv[3] if len(sys.argv) > 3 else rep}", "metric,unit,value"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        out.append(f"{h},{u},{v}")
txt = "\n".join(out) + "\n"
if len(sys.argv) > 2 and sys.argv[2] != "-":
    open(sys.argv[2], "w").write(txt)
print(txt)
