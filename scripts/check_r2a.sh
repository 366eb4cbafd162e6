#!/bin/bash
# round 2, call A: the single-pass update kernel (default) against the parity suite, then A/B bench with the
# first persistent generation (HB_UPDATE_V2=1)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runharmony.py -x -q -m gpu -k "not large and not guards" 2>&1 | tail -40
echo "parity exit: $?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v4.json
HB_UPDATE_V2=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v2.json
HB_TRACE_STEPS=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null
python - <<'PY'
import json
for n in ("v4", "v2"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d.get("regions_ms_per_step"))
    except Exception as e:
        print(n, "failed:", e)
PY
