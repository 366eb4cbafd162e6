#!/bin/bash
# single-GPU check used throughout round 2 (run under gpurun): smoke, GPU test suite, bench lines of configs 3 / 4 / 5,
# stamp traces of one CTA.  Outputs under gpurun_out/.
#   usage: scripts/check_gpu.sh [quick]      (quick: parity tests without the 200k / 1M-cell cases, no e2e / CPU leg)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1
if [ "$1" = quick ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu -k "not large" 2>&1 | tail -12
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c3.json
else
  timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -12
  HB_TRACE_HOST=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/host_trace.txt
  grep "hb_" gpurun_out/host_trace.txt | tail -8
fi
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
HB_TRACE_STEPS=0 HB_TRACE_ASSIGN=1 HB_TRACE_APPLY=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null
python - <<'PY'
import json
for n in ("c3", "c4", "c5"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"], 4), f'{d["value"]:.4g}', round(d["roofline_step"]["frac"], 4), d.get("regions_ms_per_step"),
              (d.get("e2e") or {}).get("seconds"), d.get("parity"))
    except Exception as e:
        print(n, "failed:", e)
PY
# the driver loop from plain C (examples/harmonize.c) and the tcgen05 descriptor address-map probes
gcc -std=c99 -O2 -Iinclude examples/harmonize.c -Lharmony_b200 -lharmony_b200 -Wl,-rpath,$PWD/harmony_b200 -lm -o gpurun_out/harmonize_demo \
  && timeout 120 gpurun_out/harmonize_demo 200000 > gpurun_out/harmonize_demo.txt 2>&1; tail -4 gpurun_out/harmonize_demo.txt
timeout 200 bash scripts/mb/run_layout_probe.sh gpurun_out/umma_layout_probe.txt; head -12 gpurun_out/umma_layout_probe.txt
