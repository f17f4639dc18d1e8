#!/bin/bash
# round-2 N-GPU visit: world-4 CP tests (both engines), the CP bench (nvl) and the reference FFA_FA4 CP run at N ranks
N=${1:-8}; STEPS=${2:-5}
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
nvidia-smi -L | head -2
timeout 900 python -m pytest tests/test_dist_attn_gpu.py -m gpu -q -x -k "real_cp" 2>&1 | tail -4 | tee gpurun_out/cp_tests_n$N.log
run() { # name, extra env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps $STEPS --warmup 3 $EXTRA > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_$name.json") if l.startswith("{")][-1])
    print("$name", "value", round(d["value"],1), "per_gpu", round(d["per_gpu_tflops"],1), "ms", round(d["ms_per_step"],2), "exposed", d.get("exposed_comm"), "parity", d.get("parity") and (d["parity"]["ok"], d["parity"]["rel_err"]), "e2e", d.get("e2e") and round(d["e2e"]["value"],1), d.get("reference_class"), d.get("reason"))
except Exception as e:
    print("$name", "FAILED", e)
P
  grep -v "^W09\|Warning\|warn" gpurun_out/bench_n${N}_$name.err | tail -3
}
run nvl_d4 B200_OVERLAP_DEGREE=4
run nvl_d2 B200_OVERLAP_DEGREE=2
EXTRA="--impl reference" run ref X=1
