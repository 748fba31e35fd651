"""writes the judged subset of an `ncu --page raw --csv` export as metric,unit,value rows:
python scripts/ncu_select.py gpurun_out/x_raw.csv profiles/r02_ncu_x_selected.csv"""
import csv
import sys

KEEP = ("gpu__time_duration", "dram__bytes", "dram__throughput", "gpu__dram_throughput", "sm__throughput", "l1tex__throughput", "lts__throughput",
        "lts__t_sectors", "lts__t_sector_hit_rate", "sm__warps_active", "smsp__issue_active", "smsp__inst_executed.sum", "launch__",
        "l1tex__data_pipe_lsu_wavefronts", "l1tex__data_bank_conflicts", "l1tex__t_sectors_pipe_lsu_mem_global", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "sm__inst_executed_pipe", "smsp__cycles_active.avg")
rows = list(csv.reader(open(sys.argv[1])))
h, units, vals = rows[0], rows[1], rows[2]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", vals[h.index("Kernel Name")][:60]])
    for i, m in enumerate(h):
        if m == "Kernel Name" or any(m.startswith(k) or ("." + k) in m for k in KEEP):
            w.writerow([m, units[i], vals[i]])
print(sys.argv[2])
