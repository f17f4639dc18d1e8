#!/bin/bash
# build a tuning variant of the library:  tools/dev/build_variant.sh <name> <extra nvcc flags...>
set -e
name=$1; shift
cd "$(dirname "$0")/../.."
out=magiattention_b200/lib/variants; mkdir -p $out/$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -diag-suppress 177"
for f in ffa_api ffa_plan ffa_fwd ffa_bwd ffa_aux nvl_exchange; do
  nvcc $FLAGS "$@" -c magiattention_b200/csrc/$f.cu -o $out/$name/$f.o &
done
wait
nvcc -shared -o $out/$name.so $out/$name/*.o -cudart static -Wno-deprecated-gpu-targets
echo $out/$name.so
