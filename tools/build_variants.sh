#!/bin/bash
# tools/build_variants.sh — build libdtsim.so variants into gym-duckietown_b200/variants/ (for tools/ab_all.sh on the GPU box)
#   each line of VARIANTS: <tag> <extra nvcc flags>
set -e
cd "$(dirname "$0")/../gym-duckietown_b200"
mkdir -p variants
cp libdtsim.so /tmp/libdtsim_base_keep.so
while read -r tag flags; do
  [ -z "$tag" ] && continue
  DTS_NVCC_EXTRA="$flags" python build.py --force > /dev/null 2>&1
  cp libdtsim.so "variants/libdtsim_$tag.so"
  grep -A2 "k_rasterILb0ELb0" build.log | grep -o "Used [0-9]* registers" | head -1 | sed "s/^/$tag: /"
done <<'VARIANTS'
base
gi4 -DDTS_GEO_INLINE=4
gx1 -DDTS_GEO_X=1
gx2 -DDTS_GEO_X=2
gx3 -DDTS_GEO_X=3
VARIANTS
cp /tmp/libdtsim_base_keep.so libdtsim.so
