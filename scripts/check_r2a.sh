#!/bin/bash
# round 2 GPU check: smoke, parity suite, A/B bench against the first persistent update generation (HB_UPDATE_V2=1)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if ! timeout 120 python -c "import __graft_entry__ as g; g.smoke()"; then
  echo "SMOKE FAILED with the tensor-core assignment kernel; continuing with HB_ASSIGN_FFMA=1"
  export HB_ASSIGN_FFMA=1
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1
fi
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runharmony.py -x -q -m gpu -k "not large" 2>&1 | tail -40
echo "parity exit: $?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v4.json
HB_ASSIGN_FFMA=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v4_ffma_assign.json
HB_UPDATE_V2=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v2.json
HB_TRACE_STEPS=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null
python - <<'PY'
import json
for n in ("v4", "v4_ffma_assign", "v2"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d.get("regions_ms_per_step"))
    except Exception as e:
        print(n, "failed:", e)
PY
