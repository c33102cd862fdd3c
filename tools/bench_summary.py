"""One line per config of a bench.py JSON line: value, ms/step, per-kernel ms."""
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f)
    def row(k, v):
        km = v.get('kernel_ms', {})
        print('  %-3s %10.0f  %7.3f ms  %s' % (k, v['value'], v['ms_per_step'], ' '.join('%s %.3f' % (a.replace('k_', ''), b) for a, b in km.items())))
    row('c2', d)
    for k, v in d.get('configs', {}).items(): row(k, v)
    e = d.get('e2e', {}); er = d.get('e2e_resized', {})
    print('  e2e %.0f  e2e_resized %.0f  frac %.4f  clocks %s' % (e.get('value', 0), er.get('value', 0), d['roofline']['frac'], d.get('clocks')))
