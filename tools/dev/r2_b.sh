#!/bin/bash
# round-2 visit B (N GPUs): CP parity tests on both exchange engines, then the CP bench with each engine
N=${1:-2}; STEPS=${2:-5}
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi -L; nvidia-smi topo -m 2>/dev/null | head -12
timeout 1200 python -m pytest tests/test_dist_attn_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/cp_tests_n$N.log
for eng in ${ENGINES:-nvl nccl}; do
  MAGI_B200_KV_EXCHANGE=$eng timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/bench_n${N}_$eng.json 2> gpurun_out/bench_n${N}_$eng.err
  python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_$eng.json") if l.startswith("{")][-1])
    print("$eng", "value", round(d["value"],1), "per_gpu", round(d["per_gpu_tflops"],1), "ms", round(d["ms_per_step"],2), "exposed", d["exposed_comm"], "parity", d["parity"] and (d["parity"]["ok"], d["parity"]["rel_err"]), "e2e", d["e2e"] and round(d["e2e"]["value"],1))
except Exception as e:
    print("$eng", "FAILED", e)
P
  tail -4 gpurun_out/bench_n${N}_$eng.err
done
