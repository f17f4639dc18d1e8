#!/bin/bash
# short 2-GPU visit: shared backward row statistics — unit test, CP parity on the NVLink engine, one CP bench line
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 100 python -m pytest tests/test_ffa_gpu.py -m gpu -q -x -k "backward_accumulates or sink_then" 2>&1 | tail -4 | tee gpurun_out/rowstats_unit.log
timeout 150 python -m pytest tests/test_dist_attn_gpu.py -m gpu -q -x -k "nvl" 2>&1 | tail -6 | tee gpurun_out/rowstats_cp.log
MAGI_B200_KV_EXCHANGE=nvl timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_n2_rowstats.json 2> gpurun_out/bench_n2_rowstats.err
tail -c 1500 gpurun_out/bench_n2_rowstats.json; tail -3 gpurun_out/bench_n2_rowstats.err
