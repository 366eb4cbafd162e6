#!/bin/bash
# address-map probes of tcgen05 MN-major tf32 operand descriptors (one process per setting; see umma_layout_probe.cu)
out=$(realpath -m "${1:-/dev/stdout}")
cd "$(dirname "$0")"
[ -x umma_layout_probe ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_layout_probe umma_layout_probe.cu
while read -r args; do
  timeout 15 ./umma_layout_probe $args || echo "probe $args: exit $?"
done > "$out" 2>&1 <<'CFG'
0 1 0 256 1024
0 1 2 4096 1024
1 1 0 256 1024
1 1 2 4096 1024
0 1 0 2048 128
0 1 0 128 2048
0 1 2 1024 4096
0 0 0 1024 128
0 1 4 2048 512
0 1 6 1024 256
1 1 2 1024 4096
CFG
