#!/bin/bash
# 2-GPU validation of the peer-memory step exchange (HB_UPDATE_V3=1 HB_PEER_EXCHANGE=1):
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/check_v3_multi.sh > gpurun_out/check_v3_multi.log 2>&1'
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export HB_UPDATE_V3=1
HB_PEER_EXCHANGE=1 timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu
echo "multi-GPU parity exit: $?"
run() {  # $1 = output name; environment selects the mode
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > "gpurun_out/bench2_$1.json"
}
HB_PEER_EXCHANGE=1 run xch
run v3_nccl
unset HB_UPDATE_V3
run default
python - <<'PY'
import json
for n in ("default", "v3_nccl", "xch"):
    try:
        d = json.loads(open(f"gpurun_out/bench2_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d.get("regions_ms_per_step"))
    except Exception as e:
        print(n, "failed:", e)
PY
