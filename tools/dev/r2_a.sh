#!/bin/bash
# round-2 visit A (1 GPU): reference FFA_FA4 smoke + C4, own bench C4, single-GPU parity suite
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi -L
( time timeout 900 python bench.py --impl reference --config C1 --steps 3 --warmup 2 > gpurun_out/ref_c1.json 2> gpurun_out/ref_c1.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/ref_c1.json; echo; tail -5 gpurun_out/ref_c1.err
if grep -q gpu_ffa_fa4 gpurun_out/ref_c1.json; then
  ( time timeout 900 python bench.py --impl reference --config C4 --steps 3 --warmup 3 > gpurun_out/ref_c4.json 2> gpurun_out/ref_c4.err ) 2>&1 | tail -3
  tail -c 700 gpurun_out/ref_c4.json; echo; tail -3 gpurun_out/ref_c4.err
fi
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/gputests_a.log
