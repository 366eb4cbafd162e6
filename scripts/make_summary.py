#!/usr/bin/env python
"""Builds profiles/r02_summary.md (and the measured section of DESIGN.md) from the bench JSON lines the round-2 GPU
runs left under profiles/ (copied there from gpurun_out/ by hand: gpurun_out/ is scratch)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    p = os.path.join(P, name)
    if not os.path.exists(p):
        return None
    return json.loads(open(p).read().strip().splitlines()[-1])


def fmt_regions(d):
    r = d["regions_ms_per_step"]
    return (f"assign {r['assign']:.3f} · update_R {r['update_R']:.3f} (kernel {r['k_update_steps']:.3f}) · plan {r['plan']:.3f} · "
            f"stats {r['ridge_stats']:.3f} · solve {r['ridge_solve']:.3f} · apply {r['ridge_apply']:.3f}")


def main():
    out = []
    c3 = load("r02_bench_c3.json")
    out.append("| workload (per GPU) | GPUs | ms / Harmony iteration | cells/s/iteration | whole-step HBM roofline (SURVEY §8d bytes ÷ measured 6486.5 GB/s) | per phase (ms, timing pass) |")
    out.append("|---|---:|---:|---:|---:|---|")
    for name, label in (("r02_bench_c3.json", "config 3: 1M cells × 50 PCs, 20 batches, K=100"),
                        ("r02_bench_c4.json", "config 4 shard: 1.25M × 50, dataset+donor (J=40), K=100"),
                        ("r02_bench_c5.json", "config 5 shard: 6.25M × 100, 3 covariates (J=40), K=200"),
                        ("r02_bench_2gpu_c3.json", "config 3, weak scaling"),
                        ("r02_bench_4gpu_c3.json", "config 3, weak scaling"),
                        ("r02_bench_8gpu_c3.json", "config 3, weak scaling"),
                        ("r02_bench_8gpu_c4.json", "config 4: 10M cells × 50, 2 covariates, K=100 on 8 GPUs")):
        d = load(name)
        if d is None:
            continue
        out.append(f"| {label} | {d['n_gpus']} | {d['ms_per_step']:.3f} | {d['value']:.3e} | {d['roofline_step']['frac']:.3f} | {fmt_regions(d)} |")
    out.append("")
    if c3:
        out.append("Per kernel at config 3 (algorithmic bytes of DESIGN.md §3 ÷ region time ÷ 6486.5 GB/s):")
        out.append("")
        out.append("| kernel | ms | algorithmic GB/s | fraction of HBM peak | share of the step |")
        out.append("|---|---:|---:|---:|---:|")
        for k, v in c3["roofline_kernels"].items():
            out.append(f"| {k} | {v['avg_launch_us'] / 1e3:.3f} | {v['achieved']:.0f} | {v['frac']:.3f} | {v['share_of_step']:.3f} |")
        out.append("")
        e = c3.get("e2e")
        if e:
            out.append(f"End to end through the public API with host buffers (setup H2D + native k-means initialisation + 10 iterations + "
                       f"getZcorr D2H): {e['seconds'] * 1e3:.1f} ms for 10 iterations of 1M cells = {e['value']:.3e} cells/s/iteration "
                       f"(round 1: 153–247 ms).")
        c = c3.get("cpu_baseline")
        if c:
            out.append(f"CPU restatement of the reference on the box's host cores ({c['cores']} thread, {c['sample']}): {c['value']:.3e} cells/s/iteration.")
        pz = c3.get("parity")
        if pz and "oracle64" in pz:
            o32, o64 = pz["oracle32"], pz["oracle64"]
            out.append(f"Parity on a bounded sample inside the bench run ({pz['cells']} cells, {pz['iterations']} iterations, same centroids and "
                       f"update orders): rel-L2(Z) {o64['rel_l2_Z']:.2e} vs the fp64 oracle and {o32['rel_l2_Z']:.2e} vs the reference-order fp32 "
                       f"oracle (which itself is {pz['oracle32_vs_oracle64_rel_l2_Z']:.2e} from fp64); hard cluster index differs in "
                       f"{o64['argmax_mismatch']} cells vs fp64 (largest oracle top-2 gap among them {o64['largest_oracle_top2_gap_among_them']:.1e}) "
                       f"and {o32['argmax_mismatch']} vs fp32.")
        out.append(f"Clocks during the timed region: {c3['clocks']}.")
    txt = "\n".join(out) + "\n"
    sys.stdout.write(txt)


if __name__ == "__main__":
    main()
