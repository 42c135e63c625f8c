"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
cur_file = ''
out = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or r[0] in ('', 'Function Name'): continue
    try: ln = int(r[0])
    except ValueError: continue
    col = {h: i for i, h in enumerate(hdr)}
    def g(name):
        try: return float(r[col[name]])
        except Exception: return 0.0
    out.append((g('# Samples'), cur_file, ln, r[1].strip()[:90], g('Instructions Executed'),
                {k: g(k) for k in ('stall_wait', 'stall_long_sb', 'stall_barrier', 'stall_math', 'stall_short_sb', 'stall_selected', 'stall_not_selected', 'stall_mio', 'stall_dispatch', 'stall_branch_resolving', 'stall_no_inst')}))
tot = sum(o[0] for o in out)
print('total samples', tot)
for s, f, ln, src, inst, st in sorted(out, reverse=True)[:top]:
    stalls = ' '.join('%s=%d' % (k.replace('stall_', ''), v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4] if v > 0)
    print('%5.1f%% %s:%d  inst=%.2e  [%s]\n        %s' % (100 * s / tot, f, ln, inst, stalls, src))

# region summary for stokes_grid.cu (line ranges given as name:lo-hi on the command line)
if len(sys.argv) > 3:
    regs = []
    for spec in sys.argv[3:]:
        name, rng = spec.split(':'); lo, hi = rng.split('-'); regs.append((name, int(lo), int(hi)))
    agg = {n: [0.0, 0.0, {}] for n, _, _ in regs}; other = 0.0
    for s, f, ln, src, inst, st in out:
        if f != 'stokes_grid.cu': continue
        for n, lo, hi in regs:
            if lo <= ln <= hi:
                agg[n][0] += s; agg[n][1] += inst
                for k, v in st.items(): agg[n][2][k] = agg[n][2].get(k, 0) + v
                break
    tot2 = sum(s for s, f, *_ in out if f == 'stokes_grid.cu')
    for n, (s, inst, st) in agg.items():
        stalls = ' '.join('%s=%.0f%%' % (k.replace('stall_', ''), 100 * v / max(s, 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:5])
        print('%-12s %5.1f%% samples  inst=%.3e  %s' % (n, 100 * s / tot2, inst, stalls))
