"""Selected columns of an `ncu --set full` report as a small CSV for profiles/:
   python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv"""
import csv
import subprocess
import sys

COLS = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
idx = [hdr.index(c) for c in COLS if c in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
print("wrote", sys.argv[2], len(rows) - 2, "kernels")
