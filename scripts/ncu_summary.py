#!/usr/bin/env python
"""ncu .ncu-rep -> compact per-launch CSV of the metrics the roofline uses.
  python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/r1/x.csv"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard_ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(f"{out}: {len(rows) - 2} launches x {len(idx)} metrics")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
