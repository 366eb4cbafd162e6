#!/bin/bash
# 2-GPU run-to-run variance of the free-running timed loop vs the period of the NVML clock sampler
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cp in 0.003 0.003 0.05 1.0; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --clock-period $cp 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['regions_ms_per_step']
print('period $cp', 'ms/step', round(d['ms_per_step'],3), 'regions sum', round(sum(r[k] for k in ('update_R','assign','plan','ridge_stats','ridge_solve','ridge_apply')),3), 'samples', d['clocks']['samples'])"
done
