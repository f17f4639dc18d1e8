#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
V=magiattention_b200/lib/variants
{
./tools/probe/probe_ld | tail -24
echo "== drain1 correctness"; B200_FFA_LIB=$V/drain1.so timeout 300 python tools/dev/bwd_check.py --stage A 2>&1 | grep -E "PASS|FAIL|Error" ; B200_FFA_LIB=$V/drain1.so timeout 300 python tools/dev/bwd_check.py --stage B 2>&1 | grep -E "PASS|FAIL|Error"
echo "== base"; python tools/dev/bwd_check.py --stage C 2>&1 | grep PERF; python tools/dev/fwd_check.py --stage C 2>&1 | grep PERF
for v in drain1 abl_b64 abl_b128; do echo "== $v"; B200_FFA_LIB=$V/$v.so python tools/dev/bwd_check.py --stage C 2>&1 | grep PERF; done
for v in abl_f8 abl_f16 abl_f15 abl_f31 abl_f47 abl_f79; do echo "== $v"; B200_FFA_LIB=$V/$v.so python tools/dev/fwd_check.py --stage C 2>&1 | grep PERF; done
} 2>&1 | tee gpurun_out/ablation2.txt
