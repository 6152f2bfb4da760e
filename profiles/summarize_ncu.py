#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics the roofline discussion uses.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (expected 0)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}  (ncu --set full --clock-control none; per-launch, cold-ish caches, serialised)")
    for r in rows[2:]:
        print("\n== " + r[idx["Kernel Name"]].split("(")[0])
        for key, label in WANT:
            if key in idx:
                print(f"  {label:34s} {r[idx[key]]} {units[idx[key]]}")


if __name__ == "__main__":
    main()
