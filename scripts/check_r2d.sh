#!/bin/bash
# full GPU suite + c3/c4/c5 bench lines
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1
timeout 1200 python -m pytest tests -x -q -m gpu -k "not large" 2>&1 | tail -15
echo "suite exit: $?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c3.json
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
tail -3 gpurun_out/bench_c4.err gpurun_out/bench_c5.err
HB_TRACE_APPLY=1 HB_TRACE_ASSIGN=1 HB_TRACE_STEPS=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null
python - <<'PY'
import json
for n in ("c3", "c4", "c5"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], f'{d["value"]:.4g}', d["roofline_step"]["frac"], d.get("regions_ms_per_step"))
    except Exception as e:
        print(n, "failed:", e)
PY
HB_TRACE_HOST=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/host_trace.txt
tail -2 gpurun_out/bench_e2e.json | cut -c1-600
