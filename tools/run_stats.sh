#!/bin/bash
# tools/run_stats.sh — on the GPU box: tools/raster_stats.py for the c2 and c3 maps with the -DDTS_STATS=1 variant of the library
cd "$(dirname "$0")/.."
export DTS_NO_REBUILD=1
cp gym-duckietown_b200/libdtsim.so /tmp/libdtsim_keep.so
cp gym-duckietown_b200/variants/libdtsim_stats.so gym-duckietown_b200/libdtsim.so
for m in small_loop loop_obstacles; do echo "== $m"; python tools/raster_stats.py $m 2>&1 | grep -v Warn; done
cp /tmp/libdtsim_keep.so gym-duckietown_b200/libdtsim.so
