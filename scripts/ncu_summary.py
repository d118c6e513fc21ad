#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one CSV row per distinct kernel (first launch) with the counters
DESIGN.md argues from, and profiles/ncu_traffic.json = DRAM bytes per launch for bench.py's roofline.traffic.

    python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/X_ncu_full_summary.csv [size]
"""
import csv
import json
import os
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    seen, traffic = set(), {}
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            w.writerow([r[i] for i in idx])
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            traffic[name] = float(r[ir]) * SCALE.get(units[ir], 1.0) + float(r[iw]) * SCALE.get(units[iw], 1.0)
    tj = os.path.join(os.path.dirname(out), "ncu_traffic.json")
    json.dump({"size": size, "source": os.path.basename(rep), "kernels": traffic}, open(tj, "w"), indent=1)
    print("wrote", out, tj)


if __name__ == "__main__":
    main()
