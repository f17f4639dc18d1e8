#!/bin/bash
# first GPU call of the next round: validate the CTA-pair UMMA forms before any kernel is built on them.
# (bounded: a wrong barrier protocol must end in a trap / timeout, never in a hung box)
set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/probe/probe2cta tools/probe/probe2cta.cu -lcuda 2>&1 | grep -v warning | tail -5
timeout 30 tools/probe/probe2cta 2>&1 | tee gpurun_out/probe2cta.txt
nvidia-smi --query-gpu=name,clocks.sm,power.draw --format=csv
