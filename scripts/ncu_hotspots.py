#!/usr/bin/env python
"""Per-kernel stall summary of `ncu --set full` captures: the warp-state ratios of the raw page and the SASS
instructions that collect the most stall samples (source page), as markdown.

  python scripts/ncu_hotspots.py gpurun_out/r02_k_*.ncu-rep > profiles/r02_hotspots.md
"""
import csv
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def main(reps):
    print("# Round 2 — where the warps wait (from the committed `ncu --set full` captures; `scripts/ncu_hotspots.py`)\n")
    print("Ratios are `smsp__average_warps_issue_stalled_*_per_issue_active`: warps in that state per issued instruction "
          "(1.0 `selected` = the issuing warp itself).  Samples are the PC-sampling counts of the source page.\n")
    for rep in reps:
        raw = page(rep, "raw")
        d = dict(zip(raw[0], raw[2]))
        kname = d["Kernel Name"]
        print(f"## `{kname}` ({rep.split('/')[-1]})\n")
        keys = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]
        print("* " + " · ".join(f"`{k}` = {d[k]}" for k in keys if k in d))
        st = [(k.split("issue_stalled_")[1].split("_per_issue")[0], num(v)) for k, v in d.items()
              if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
              and "not_issued" not in k]
        st.sort(key=lambda x: -x[1])
        print("* warp states per issued instruction: " + ", ".join(f"{k} {v:.2f}" for k, v in st[:8]) + "\n")
        src = page(rep, "source")
        hdr = src[1]
        col = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        rows = [r for r in src[2:] if len(r) == len(hdr)]
        total = sum(num(r[col["# Samples"]]) for r in rows) or 1.0
        rows.sort(key=lambda r: -num(r[col["# Samples"]]))
        print("| # samples | share | SASS | dominant stall |\n|---:|---:|---|---|")
        for r in rows[:14]:
            n = num(r[col["# Samples"]])
            reasons = sorted(((num(r[col[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
            rs = ", ".join(f"{c} {int(v)}" for v, c in reasons if v > 0)
            print(f"| {int(n)} | {100 * n / total:.1f}% | `{r[col['Source']].strip()}` | {rs} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
