"""Aggregates an ncu source-page CSV by line ranges of one file: ncu_regions.py dump.csv file.cu name:lo-hi ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); fname = sys.argv[2]
regions = []
for spec in sys.argv[3:]:
    n, rng = spec.split(':'); lo, hi = rng.split('-'); regions.append((n, int(lo), int(hi)))
hdr = None; cur = ''; agg = {}; tot = 0
keys = ('stall_wait', 'stall_long_sb', 'stall_barrier', 'stall_math', 'stall_short_sb', 'stall_selected',
        'stall_not_selected', 'stall_mio', 'stall_dispatch', 'stall_no_inst', 'stall_lg')
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    col = {h: i for i, h in enumerate(hdr)}
    def g(k):
        try: return float(r[col[k]])
        except Exception: return 0.0
    s = g('# Samples'); tot += s
    name = 'other:' + cur
    if cur == fname:
        for n, lo, hi in regions:
            if lo <= ln <= hi: name = n
    a = agg.setdefault(name, [0.0, 0.0, {}]); a[0] += s; a[1] += g('Instructions Executed')
    for k in keys: a[2][k] = a[2].get(k, 0) + g(k)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    st = ' '.join('%s=%.0f' % (a.replace('stall_', ''), b) for a, b in sorted(v[2].items(), key=lambda kv: -kv[1])[:5] if b > 0)
    print('%-26s %5.1f%%  inst %.2e  [%s]' % (k, 100 * v[0] / tot, v[1], st))
