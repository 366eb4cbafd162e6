#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runharmony.py -x -q -m gpu -k "not large" 2>&1 | tail -12
echo "suite exit: $?"
HB_TRACE_HOST=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_e2e.json 2> gpurun_out/host_trace.txt
grep "hb_" gpurun_out/host_trace.txt | tail -7
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python - <<'PY'
import json
for n in ("c3_e2e", "c4", "c5"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], f'{d["value"]:.4g}', d["roofline_step"]["frac"], d.get("regions_ms_per_step"), (d.get("e2e") or {}).get("seconds"))
    except Exception as e:
        print(n, "failed:", e)
PY
