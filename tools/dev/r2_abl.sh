#!/bin/bash
# kernel ablation study (timing only): which part of each kernel's loop bounds it
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
V=magiattention_b200/lib/variants
{
echo "== base"; python tools/dev/bwd_check.py --stage C 2>&1 | grep PERF; python tools/dev/fwd_check.py --stage C 2>&1 | grep PERF
for v in abl_b1 abl_b3 abl_b7 abl_b8 abl_b16 abl_b32; do echo "== $v"; B200_FFA_LIB=$V/$v.so python tools/dev/bwd_check.py --stage C 2>&1 | grep PERF; done
for v in abl_f1 abl_f2 abl_f4 abl_f7; do echo "== $v"; B200_FFA_LIB=$V/$v.so python tools/dev/fwd_check.py --stage C 2>&1 | grep PERF; done
} | tee gpurun_out/ablation.txt
BENCH_REF_SPLIT=1 timeout 600 python bench.py --impl reference --config C4 --steps 3 --warmup 3 > gpurun_out/ref_c4_split.json 2> gpurun_out/ref_c4_split.err
grep "REF_SPLIT" gpurun_out/ref_c4_split.err
