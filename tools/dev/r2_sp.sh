#!/bin/bash
# split-row forward: correctness ladders + timing next to the default layout
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out
export B200_FFA_LIB=magiattention_b200/lib/variants/${1:-sp}.so
{
for sp in 1 0; do
  echo "== split=$sp correctness"
  for st in A B; do B200_FFA_FWD_SPLIT=$sp timeout 300 python tools/dev/fwd_check.py --stage $st 2>&1 | grep -E "PASS|FAIL|Error|error" | awk '{print $1}' | sort | uniq -c | tr '\n' ' '; done; echo
  B200_FFA_FWD_SPLIT=$sp timeout 300 python tools/dev/fwd_check.py --stage A 2>&1 | grep -E "FAIL|rror" | head -4
  B200_FFA_FWD_SPLIT=$sp timeout 300 python tools/dev/fwd_check.py --stage B 2>&1 | grep -E "FAIL|rror" | head -4
  echo "== split=$sp perf"; B200_FFA_FWD_SPLIT=$sp python tools/dev/fwd_check.py --stage C 2>&1 | grep PERF
done
} 2>&1 | tee gpurun_out/split.txt
