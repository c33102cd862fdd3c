#!/bin/bash
# tools/build_variants.sh — build libdtsim.so variants into gym-duckietown_b200/variants/ (for tools/ab_all.sh on the GPU box)
#   each line of VARIANTS: <tag> <extra nvcc flags>
set -e
cd "$(dirname "$0")/../gym-duckietown_b200"
mkdir -p variants
while read -r tag flags; do
  [ -z "$tag" ] && continue
  DTS_NVCC_EXTRA="$flags" python build.py --force > /dev/null 2>&1
  cp libdtsim.so "variants/libdtsim_$tag.so"
  grep -A2 "k_rasterILb0ELb0" build.log | grep -o "Used [0-9]* registers" | head -1 | sed "s/^/$tag: /"
done <<'VARIANTS'
base
nosolo -DDTS_SOLO=0
VARIANTS
cp variants/libdtsim_base.so libdtsim.so   # the tree's library = the base variant, built from the current sources
