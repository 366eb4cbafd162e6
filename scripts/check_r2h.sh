#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config5 or K128 or K100 or config4 or fallback" -s 2>&1 | grep -E "relL2|passed|failed|Error|error" | cut -c1-220 | tail -20
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c3_h.json
python - <<'PY'
import json
for n in ("c3_h", "c4", "c5"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], f'{d["value"]:.4g}', d["roofline_step"]["frac"], d.get("regions_ms_per_step"))
    except Exception as e:
        print(n, "failed:", e)
PY
