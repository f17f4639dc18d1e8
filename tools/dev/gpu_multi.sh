#!/bin/bash
# multi-GPU visit: real-NCCL CP parity tests, then the CP bench at N ranks
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_dist_attn_gpu.py -m gpu -q -x 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 2 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -5 gpurun_out/bench_n$N.err
