#!/bin/bash
# last single-GPU visit of round 2: full GPU parity suite, smoke(), C4 / C2 / C3 bench lines, C5 one-rank share with plan-kernel share
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/gputests_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in C4 C2 C3; do
  timeout 400 python bench.py --config $c --steps 5 --warmup 3 $( [ $c = C4 ] || echo --no-cpu ) > gpurun_out/bench_final_$c.json 2> gpurun_out/bench_final_$c.err
  python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_final_$c.json") if l.startswith("{")][-1])
    print("$c", "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), "fwd", round(d["roofline"].get("fwd_kernel",{}).get("achieved",0),1), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$c FAILED", e)
P
done
timeout 300 python tools/dev/c5_share.py 2>/dev/null | tail -1 | tee gpurun_out/c5_share.json
