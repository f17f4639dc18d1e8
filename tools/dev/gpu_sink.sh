#!/bin/bash
# attention sink / max_logits / fuzz bring-up: new parity tests first (bounded), the whole GPU suite, fwd throughput check
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "sink or max_logits or fuzz or error_behaviour" 2>&1 | tail -40
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 120 python tools/dev/fwd_check.py --stage C 2>&1 | tail -5
