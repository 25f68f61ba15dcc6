#!/usr/bin/env python
"""Turn an `ncu --set full` capture into the markdown summary kept under profiles/.

    ncu -i gpurun_out/prof_X.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_ncu.py /tmp/raw.csv "title" > profiles/X_ncu_full_summary.md

One table per distinct kernel (first captured launch of each), restricted to the
counters the roofline discussion in DESIGN.md uses.
"""
import csv
import sys

WANT = [
    "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "ncu --set full summary")
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    seen = {}
    for r in body:
        name = r[idx["Kernel Name"]]
        seen.setdefault(name, []).append(r)
    for name, rs in seen.items():
        r = rs[0]
        times = [float(x[idx["gpu__time_duration.sum"]]) for x in rs]
        print(f"## {name}\n")
        print(f"captured launches: {len(rs)}; gpu__time_duration per launch [{units[idx['gpu__time_duration.sum']]}]: "
              + ", ".join(f"{t:.1f}" for t in times) + "\n")
        print("| metric | unit | value |\n|---|---|---|")
        for w in WANT:
            if w in idx:
                print(f"| {w} | {units[idx[w]]} | {r[idx[w]]} |")
        # every other warp-stall reason that matters (> 0.3 warps per issue)
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and h not in WANT:
                try:
                    if float(r[idx[h]]) > 0.3:
                        print(f"| {h} | {units[idx[h]]} | {r[idx[h]]} |")
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    main()
