#!/bin/bash
# head_dim 64 bring-up: the new parity test first (bounded), then the whole GPU suite, then d=64 kernel throughput
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ffa_gpu.py -m gpu -q -x -k "head_dim_64 or error_behaviour" 2>&1 | tail -15
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
FFA_CHECK_D=64 timeout 120 python tools/dev/fwd_check.py --stage C 2>&1 | tail -5
FFA_CHECK_D=64 timeout 120 python tools/dev/bwd_check.py --stage C 2>&1 | tail -5
timeout 120 python tools/dev/fwd_check.py --stage C 2>&1 | tail -5
timeout 120 python tools/dev/bwd_check.py --stage C 2>&1 | tail -5
