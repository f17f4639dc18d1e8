#!/bin/bash
# last single-GPU visit of round 2: full GPU parity suite, smoke(), C4 / C1 / C2 / C3 bench lines, C5 one-rank share with the
# plan-kernel share, ncu launch list (64K), ncu --set full of the two attention kernels (32K), DRAM traffic of both at 256K
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/gputests_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in C4 C1 C2 C3; do
  timeout 400 python bench.py --config $c --steps 5 --warmup 3 $( [ $c = C4 ] || echo --no-cpu ) > gpurun_out/bench_final_$c.json 2> gpurun_out/bench_final_$c.err
  python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_final_$c.json") if l.startswith("{")][-1])
    print("$c", "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), "fwd", round(d["roofline"].get("fwd_kernel",{}).get("achieved",0),1), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$c FAILED", e)
P
  tail -2 gpurun_out/bench_final_$c.err | grep -v Warn
done
timeout 300 python tools/dev/c5_share.py 2>/dev/null | tail -1 | tee gpurun_out/c5_share.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --seqlen 65536 > gpurun_out/bench_ncu_list.json 2> gpurun_out/ncu_list.err
tail -1 gpurun_out/ncu_list.err
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:ffa_(fwd|bwd)_kernel' -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --seqlen 32768 > gpurun_out/bench_ncu_full.json 2> gpurun_out/ncu_full.err
tail -1 gpurun_out/ncu_full.err
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:ffa_(fwd|bwd)_kernel' -c 2 \
    --csv --log-file gpurun_out/traffic_256k.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_ncu_traffic.json 2> gpurun_out/ncu_traffic.err
tail -1 gpurun_out/ncu_traffic.err
ls -la gpurun_out | tail -12
