#!/bin/bash
# Validation of the experimental update kernel (HB_UPDATE_V3=1, harmony_b200/csrc/update_kernel3.cuh) on a GPU
# box: the whole parity suite with the switch on, then the bench with and without it.
#   gpurun --timeout 1500 -- 'bash scripts/check_v3.sh > gpurun_out/check_v3.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# does the TMA unit sustain the row gather of k_update_steps3?  (prebuilt binary travels with the snapshot)
[ -x scripts/mb/tma_rows ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/mb/tma_rows scripts/mb/tma_rows.cu
timeout 60 scripts/mb/tma_rows
# a hang costs its whole timeout in GPU minutes: smoke every switch first with a short leash
for sw in HB_UPDATE_V3 HB_APPLY_V2 HB_STATS_V2 HB_ASSIGN_V2 HB_DOWNLOAD_MT; do
  env $sw=1 timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
  echo "$sw smoke exit: $?"
done
HB_UPDATE_V3=1 timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runharmony.py -x -q -m gpu
echo "parity exit: $?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_default.json
HB_UPDATE_V3=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v3.json
for sw in HB_APPLY_V2 HB_STATS_V2 HB_ASSIGN_V2; do   # the other experimental kernels, one at a time
  env $sw=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_run or stepwise"
  echo "$sw parity exit: $?"
  env $sw=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$sw.json
done
HB_UPDATE_V3=1 HB_TRACE_STEPS=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null
# the golden tables of the reference's vignette through the library (init: default path; after cluster_cpp: legacy step)
timeout 300 python -m pytest tests/test_gpu_runharmony.py -q -m gpu -k vignette -rxX
# end-to-end: where do the one-off milliseconds go, and does the threaded download help?
HB_TRACE_HOST=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e_trace.json 2> gpurun_out/host_trace.txt
HB_DOWNLOAD_MT=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e_mt.json
python - <<'PY'
import json
for n in ("default", "v3", "HB_APPLY_V2", "HB_STATS_V2", "HB_ASSIGN_V2", "e2e_trace", "e2e_mt"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], (d.get("e2e") or {}).get("seconds"), d.get("regions_ms_per_step"))
    except Exception as e:
        print(n, "failed:", e)
PY
