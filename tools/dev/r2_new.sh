#!/bin/bash
# correctness ladders + timing of the new kernel variants, then the timeline of the traced build
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
V=magiattention_b200/lib/variants
{
for v in new_e2 new; do
  echo "== $v correctness"
  for st in A B; do
    B200_FFA_LIB=$V/$v.so timeout 300 python tools/dev/fwd_check.py --stage $st 2>&1 | grep -E "PASS|FAIL|Error|error" | awk '{print $1}' | sort | uniq -c | tr '\n' ' '
    B200_FFA_LIB=$V/$v.so timeout 300 python tools/dev/bwd_check.py --stage $st 2>&1 | grep -E "PASS|FAIL|Error|error" | awk '{print $1}' | sort | uniq -c | tr '\n' ' '
  done; echo
  B200_FFA_LIB=$V/$v.so timeout 300 python tools/dev/fwd_check.py --stage A 2>&1 | grep -E "FAIL|rror" | head -5
done
for v in base new new_e1 new_e2; do
  echo "== $v"
  if [ "$v" = base ]; then unset B200_FFA_LIB; else export B200_FFA_LIB=$V/$v.so; fi
  python tools/dev/bwd_check.py --stage C 2>&1 | grep PERF; python tools/dev/fwd_check.py --stage C 2>&1 | grep PERF
done
B200_FFA_LIB=$V/new_tr.so python tools/dev/trace.py 2>&1
} 2>&1 | tee gpurun_out/new_kernels.txt
