"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel -> markdown (profiles/*_launch_summary.md).

    python scripts/summarize_launches.py gpurun_out/r1i_launches.csv "r1i" > profiles/r1i_launch_summary.md
"""
import csv
import sys
from collections import OrderedDict


def main(path, tag):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], us))
    agg = OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"# {tag} launch list (ncu --metrics gpu__time_duration.sum --clock-control none, `bench.py --steps 2 --warmup 3`), aggregated per kernel\n")
    print(f"{len(rows)} launches captured. Times are cold-cache, serialised per-launch durations.\n")
    print("| kernel | launches | mean us | total ms | share |\n|---|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[:90]} | {n} | {us / n:.1f} | {us / 1e3:.2f} | {100 * us / total:.1f} % |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
