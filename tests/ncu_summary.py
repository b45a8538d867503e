"""Summarise an `ncu --set full` capture for profiles/: one row per launch, the metrics the roofline argument uses.
    python tests/ncu_summary.py gpurun_out/capture.ncu-rep > profiles/rN_<what>_full_summary.csv      (reads the report; needs no GPU)"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__cluster_dim_x",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
head, units, data = rows[0], rows[1], rows[2:]
cols = [head.index(c) for c in ("ID", "Kernel Name", "Grid Size", "Block Size")] + [head.index(m) for m in METRICS if m in head]
w = csv.writer(sys.stdout)
w.writerow([head[i] for i in cols])
w.writerow([units[i] for i in cols])
for r in data:
    w.writerow([r[i][:70] if head[i] == "Kernel Name" else r[i] for i in cols])
