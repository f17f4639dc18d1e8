#!/bin/bash
# bring-up of the experimental CTA-pair forward (ffa_fwd2.cu, B200_FFA_FWD_VARIANT=2): correctness stages of the forward harness
# against the oracle, then throughput next to the default kernel.  Every step is bounded; barrier waits trap after 4 s.
set -x
mkdir -p gpurun_out
export B200_FFA_FWD_VARIANT=2
timeout 120 python tools/dev/fwd_check.py --stage A 2>&1 | tail -15
timeout 200 python tools/dev/fwd_check.py --stage B 2>&1 | tail -20
timeout 120 python tools/dev/fwd_check.py --stage C 2>&1 | tail -5
unset B200_FFA_FWD_VARIANT
timeout 120 python tools/dev/fwd_check.py --stage C 2>&1 | tail -5
