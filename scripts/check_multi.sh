#!/bin/bash
# multi-GPU check (run with gpurun --gpus N, N >= 2): 2-GPU == 1-GPU parity test, then the N-GPU bench with the
# peer-memory step exchange (default) and with one all-reduce per block step (HB_NO_PEER_EXCHANGE=1)
set -x
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | tail -15 | tee gpurun_out/multi_parity_${N}gpu.log
for mode in xch nccl; do
  if [ $mode = nccl ]; then KS="--kernel-set 8"; else KS=""; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $KS > gpurun_out/bench_${N}gpu_${mode}.json 2> gpurun_out/bench_${N}gpu_${mode}.err
  tail -3 gpurun_out/bench_${N}gpu_${mode}.err
done
for v in "" "--kernel-set 16"; do
  tag=c4$(echo $v | tr -d ' -')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus $N --config c4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $v > gpurun_out/bench_${N}gpu_${tag}.json 2> gpurun_out/bench_${N}gpu_${tag}.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_1gpu_ref.json
python - <<PY
import json
for n in ("${N}gpu_xch", "${N}gpu_nccl", "${N}gpu_c4", "${N}gpu_c4kernelset16", "1gpu_ref"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "value", f'{d["value"]:.4g}', d.get("regions_ms_per_step"))
    except Exception as e:
        print(n, "failed:", e)
PY
HB_TRACE_STEPS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
cp gpurun_out/step_trace.txt gpurun_out/step_trace_${N}gpu.txt
