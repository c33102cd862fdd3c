#!/bin/bash
# per-kernel A/B: tools/ab_kernel.sh <variant.so> <label> <kernel regex> [bench args]  -> mean launch duration under ncu
cd "$(dirname "$0")/../gym-duckietown_b200"
v=$1; label=$2; k=$3; shift 3
cp "$v" /tmp/libdtsim_variant.so; cp /tmp/libdtsim_variant.so libdtsim.so
cd ..
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$k" -s 4 -c 8 --csv --log-file /tmp/ab_$label.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('/tmp/ab_$label.csv')) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=='ID'][0]; H=rows[h]
v=[float(r[H.index('Metric Value')].replace(',','')) for r in rows[h+1:]]
print('$label', '$k', 'mean %.1f us  min %.1f  n=%d' % (sum(v)/len(v)/1e3, min(v)/1e3, len(v)))
PY
