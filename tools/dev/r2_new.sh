#!/bin/bash
# correctness ladders + timing of kernel variants, then the timeline of the traced build:  r2_new.sh <check-variant> <traced> <perf variants...>
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
V=magiattention_b200/lib/variants
CHK=$1; TR=$2; shift 2
{
echo "== $CHK correctness"
for st in A B; do
  B200_FFA_LIB=$V/$CHK.so timeout 300 python tools/dev/fwd_check.py --stage $st 2>&1 | grep -E "PASS|FAIL|Error|error" | awk '{print $1}' | sort | uniq -c | tr '\n' ' '
  B200_FFA_LIB=$V/$CHK.so timeout 300 python tools/dev/bwd_check.py --stage $st 2>&1 | grep -E "PASS|FAIL|Error|error" | awk '{print $1}' | sort | uniq -c | tr '\n' ' '
done; echo
for v in "$@"; do
  echo "== $v"
  if [ "$v" = base ]; then unset B200_FFA_LIB; else export B200_FFA_LIB=$V/$v.so; fi
  python tools/dev/bwd_check.py --stage C 2>&1 | grep PERF; python tools/dev/fwd_check.py --stage C 2>&1 | grep PERF
done
B200_FFA_LIB=$V/$TR.so python tools/dev/trace.py 2>&1
} 2>&1 | tee gpurun_out/new_kernels.txt
