#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/dev/nvl_probe.py 2>&1 | grep -v "^W09\|\*\*\*\*\|OMP_NUM" | tee gpurun_out/nvl_probe.log | tail -60
