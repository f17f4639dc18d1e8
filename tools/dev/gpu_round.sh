#!/bin/bash
# one GPU-box visit: tests, smoke, bench, ncu launch list, ncu full capture of the two attention kernels,
# DRAM traffic of the dominant kernel at the full bench size
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --seqlen 65536 > gpurun_out/bench_ncu_list.json 2> gpurun_out/ncu_list.err
tail -3 gpurun_out/ncu_list.err
ncu --set full --clock-control none --import-source on -k 'regex:ffa_(fwd|bwd)_kernel' -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --seqlen 32768 > gpurun_out/bench_ncu_full.json 2> gpurun_out/ncu_full.err
tail -3 gpurun_out/ncu_full.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:ffa_(fwd|bwd)_kernel' -c 2 \
    --csv --log-file gpurun_out/traffic_256k.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_ncu_traffic.json 2> gpurun_out/ncu_traffic.err
tail -3 gpurun_out/ncu_traffic.err
timeout 300 python tools/dev/c2_bench.py 2>&1 | tail -3 | tee gpurun_out/c2_like.txt
timeout 300 python tools/dev/c2_bench.py --c5 2>&1 | tail -3 | tee gpurun_out/c5_like.txt
ls -la gpurun_out
