#!/usr/bin/env python
"""Summarise ncu outputs for profiles/: (1) a launch list CSV (gpu__time_duration.sum per launch) ->
per-kernel totals and shares; (2) .ncu-rep files (--set full) -> the metrics the roofline uses.

    python scripts/ncu_summary.py --launches gpurun_out/launches.csv --reps gpurun_out/a.ncu-rep ... > profiles/x.md
"""
import argparse
import collections
import csv
import io
import subprocess

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]


def launches(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v * 1e6 if unit == "s" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"### launch list `{path}` (cold-cache, serialised; compare shares, not absolutes)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:70]}` | {n} | {t:.1f} | {t / n:.2f} | {t / tot:.1%} |")
    print()


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"### `{path}` (ncu --set full --clock-control none)\n")
    for row in rows[2:]:
        print(f"**{row[hdr.index('Kernel Name')][:90]}**\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"- `{w}` = {row[i]} {units[i]}")
        print()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches", nargs="*", default=[])
    ap.add_argument("--reps", nargs="*", default=[])
    a = ap.parse_args()
    for p in a.launches:
        launches(p)
    for p in a.reps:
        rep(p)
