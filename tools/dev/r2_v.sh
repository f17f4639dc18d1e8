#!/bin/bash
# time named library variants:  r2_v.sh <variant> [<variant> ...]   ("base" = the in-tree library)
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
V=magiattention_b200/lib/variants
{
for v in "$@"; do
  echo "== $v"
  if [ "$v" = base ]; then unset B200_FFA_LIB; else export B200_FFA_LIB=$V/$v.so; fi
  python tools/dev/bwd_check.py --stage C 2>&1 | grep PERF; python tools/dev/fwd_check.py --stage C 2>&1 | grep PERF
done
} 2>&1 | tee gpurun_out/variants.txt
