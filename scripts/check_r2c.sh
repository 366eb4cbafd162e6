#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runharmony.py -x -q -m gpu -k "not large" 2>&1 | tail -15
echo "parity exit: $?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c.json
HB_U5_FLAGS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c_nored.json
HB_U5_FLAGS=3 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c_nored_nostore.json
HB_TRACE_ASSIGN=1 HB_TRACE_APPLY=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null
python - <<'PY'
import json
for n in ("c", "c_nored", "c_nored_nostore"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d.get("regions_ms_per_step"))
    except Exception as e:
        print(n, "failed:", e)
PY
