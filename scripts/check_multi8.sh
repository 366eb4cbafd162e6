#!/bin/bash
# 8-GPU check (gpurun --gpus 8): 2-GPU parity test, weak-scaling bench of config 3 at N = 8 (and 4), BASELINE.json
# config 4 (10M cells, 2 covariates) sharded over the 8 GPUs
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | tail -8 | tee gpurun_out/multi_parity_8gpu_box.log
run() {  # N config tag extra
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $1 --config $2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $4 > gpurun_out/bench_$3.json 2> gpurun_out/bench_$3.err
  tail -2 gpurun_out/bench_$3.err
}
run 8 c3 8gpu_c3 ""
run 8 c3 8gpu_c3_noplanoverlap "--kernel-set 16"
run 8 c4 8gpu_c4 ""
run 4 c3 4gpu_c3 ""
HB_TRACE_STEPS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
cp gpurun_out/step_trace.txt gpurun_out/step_trace_8gpu.txt
python - <<'PY'
import json
for n in ("8gpu_c3", "8gpu_c3_noplanoverlap", "8gpu_c4", "4gpu_c3"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "value", f'{d["value"]:.4g}', d.get("regions_ms_per_step"))
    except Exception as e:
        print(n, "failed:", e)
PY
