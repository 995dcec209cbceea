"""Per-launch table from an `ncu --metrics ... --csv` log: one line per launch longer than a threshold.
usage: python tools/ncu_kernel_table.py metrics.csv [min_ms]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
min_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, mi, vi, ii, gi, bi = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size", "Block Size"))
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        v = float("nan")
    L.setdefault((r[ii], r[ki], r[gi], r[bi]), {})[r[mi]] = v
print("id,kernel,grid,block,ms,dram_read_GB,dram_write_GB,l2_hit_pct,l1_hit_pct,warp_inst_G,warps_active_pct,issue_active_pct,alu_pipe_pct,dram_throughput_pct")
for (i, k, g, b), m in L.items():
    t = m.get("gpu__time_duration.sum", 0) / 1e6
    if t < min_ms:
        continue
    name = k.replace("void ", "").replace("cj::", "")
    print(",".join(str(x) for x in (i, '"' + name + '"', '"' + g + '"', '"' + b + '"', round(t, 3), round(m["dram__bytes_read.sum"] / 1e9, 3), round(m["dram__bytes_write.sum"] / 1e9, 3),
                                    round(m["lts__t_sector_hit_rate.pct"], 1), round(m["l1tex__t_sector_hit_rate.pct"], 1), round(m["smsp__inst_executed.sum"] / 1e9, 3),
                                    round(m["sm__warps_active.avg.pct_of_peak_sustained_active"], 1), round(m["smsp__issue_active.avg.pct"], 1),
                                    round(m["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"], 1), round(m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"], 1))))
