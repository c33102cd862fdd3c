"""Per-source-line sample and instruction shares from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, x in enumerate(r) if x and x[0] == "Line No"][0]
h = r[hi]
ix = {}
for i, n in enumerate(h):
    ix.setdefault(n, i)
def f(x):
    try: return float(x)
    except ValueError: return 0.0
rows = [x for x in r[hi + 1:] if len(x) == len(h) and x[2] == "-"]   # the CUDA-line aggregate rows
ts = sum(f(x[ix["# Samples"]]) for x in rows); ti = sum(f(x[ix["Instructions Executed"]]) for x in rows)
print("samples", ts, "warp-instructions", ti)
for x in sorted(rows, key=lambda x: -f(x[ix["# Samples"]]))[:top]:
    print("%5s %5.1f%% smp %5.1f%% inst  %s" % (x[0], 100 * f(x[ix["# Samples"]]) / ts, 100 * f(x[ix["Instructions Executed"]]) / ti, x[1][:120]))
