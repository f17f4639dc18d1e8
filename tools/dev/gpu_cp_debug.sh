#!/bin/bash
# multi-GPU debugging visit: full pytest output of the real-NCCL CP tests into gpurun_out/
N=${1:-4}
mkdir -p gpurun_out
python -m pytest tests/test_dist_attn_gpu.py -m gpu -q -x -k real_cp 2>&1 | tail -120 > gpurun_out/test_cp_$N.log
grep -E "Error|error|assert|passed|failed|mismatch|inf-norm" gpurun_out/test_cp_$N.log | head -40
