#!/bin/bash
# Build the UNMODIFIED reference (MagiAttention FFA_FA4 backend) into baseline/_ref so that
# `bench.py --impl reference` can time the incumbent on the same B200 box.
#
# Follows /root/reference/scripts/install_magiattn_on_blackwell.sh:21-52 with the steps that need an
# index (pip install -r requirements.txt, versioningit) replaced by direct builds of the same sources:
#   * magi_to_hstu_cuda / create_block_mask_cuda  (flash-attention fork, csrc/utils/*/setup.py)
#   * flash_attn_cute                             (pure-python CuTe-DSL package = flash_attn/cute)
#   * magi_attn_ext                               (csrc/extensions, cmake; MAGI_ATTENTION_PREBUILD_FFA=0)
#   * the magi_attention python package, copied as is (no source edits)
# magi_attn_comm (NVSHMEM native grpcoll) is skipped exactly as MAGI_ATTENTION_SKIP_MAGI_ATTN_COMM_BUILD=1
# does; the published CP tuning uses the a2av path which does not need it.
# The build writes into its source tree, so it runs from a copy under /tmp (/root/reference is read-only).
# Output: baseline/_ref/ (git-ignored, NOT gpurun-ignored → travels to the GPU box).
set -uo pipefail
SRC=${REF_SRC:-/root/reference}
REPO=$(cd "$(dirname "$0")/.." && pwd)
OUT=$REPO/baseline/_ref
W=${REF_WORK:-/tmp/refbuild}
LOG=$OUT/build.log
[ -d "$SRC" ] || { echo "no reference at $SRC"; exit 1; }
mkdir -p "$OUT"; : > "$LOG"
rm -rf "$W"; mkdir -p "$W"
# copy sources only (skip assets/docs: hundreds of MB of PNGs)
(cd "$SRC" && tar cf - --exclude=assets --exclude=docs --exclude=.git .) | (cd "$W" && tar xf -)
export TORCH_CUDA_ARCH_LIST="10.0"
export MAGI_ATTENTION_BUILD_COMPUTE_CAPABILITY=100
export MAX_JOBS=${MAX_JOBS:-8}
FA=$W/magi_attention/functional/flash-attention
ok=1
for ext in magi_to_hstu create_block_mask; do
  echo "== $ext" | tee -a "$LOG"
  (cd "$FA/csrc/utils/$ext" && python setup.py build_ext --inplace >> "$LOG" 2>&1) || { echo "FAILED $ext" | tee -a "$LOG"; ok=0; }
  cp -f "$FA/csrc/utils/$ext"/*.so "$OUT"/ 2>/dev/null
done
echo "== flash_attn_cute" | tee -a "$LOG"
rm -rf "$OUT/flash_attn_cute"; cp -r "$FA/flash_attn/cute" "$OUT/flash_attn_cute"
echo "== magi_attn_ext (cmake)" | tee -a "$LOG"
B=$W/magi_attention/csrc/extensions/build; mkdir -p "$B"
(cd "$B" && cmake .. -DCMAKE_PREFIX_PATH="$(python -c 'import torch;print(torch.utils.cmake_prefix_path)')" \
      -DMAGI_CUDA_ARCHITECTURES="100" >> "$LOG" 2>&1 && make -j"$MAX_JOBS" >> "$LOG" 2>&1) \
  || { echo "FAILED magi_attn_ext" | tee -a "$LOG"; ok=0; }
echo "== magi_attention package" | tee -a "$LOG"
rm -rf "$OUT/magi_attention"; mkdir -p "$OUT/magi_attention"
(cd "$W/magi_attention" && tar cf - --exclude=csrc --exclude='functional/flash-attention' --exclude=__pycache__ .) \
  | (cd "$OUT/magi_attention" && tar xf -)
# bench drivers/helpers of the reference that the reference arm mirrors (data only, not imported by the product)
mkdir -p "$OUT/exps_data"
cp -f "$SRC/exps/dist_attn/benchmark/datasets/default/doc_length_distribution.csv" "$OUT/exps_data/" 2>/dev/null
ls -la "$OUT" | tee -a "$LOG"
[ $ok = 1 ] && echo "REFERENCE BUILD OK" | tee -a "$LOG" || echo "REFERENCE BUILD INCOMPLETE" | tee -a "$LOG"
