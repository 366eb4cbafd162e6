#!/bin/bash
# full GPU suite (incl. large) + c3 bench with e2e + ncu evidence
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25
echo "suite exit: $?"
HB_TRACE_HOST=1 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3_full.json 2> gpurun_out/host_trace.txt
grep "hb_" gpurun_out/host_trace.txt | tail -8
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c3_full.json").read().strip().splitlines()[-1])
print("c3", d["ms_per_step"], f'{d["value"]:.4g}', d["roofline_step"]["frac"], d["regions_ms_per_step"], d["e2e"], d["cpu_baseline"])
PY
bash scripts/profile_r2.sh
