#!/bin/bash
# A/B of libdtsim.so variants inside one GPU call: tools/ab.sh <variant.so> <label> [bench args...]
set -e
cd "$(dirname "$0")/../gym-duckietown_b200"
v=$1; label=$2; shift 2
cp libdtsim.so /tmp/libdtsim_keep.so
cp "$v" /tmp/libdtsim_variant.so
cp /tmp/libdtsim_variant.so libdtsim.so
cd ..
python bench.py --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', round(d['value']), 'env-steps/s  ms/step', round(d['ms_per_step'],4), 'render_ms', round(d['roofline']['kernel_ms'],4))"
cp /tmp/libdtsim_keep.so gym-duckietown_b200/libdtsim.so
