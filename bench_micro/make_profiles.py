"""Turns the scratch outputs of bench_micro/r2_profiles.sh (gpurun_out/r2f_*) into the tracked files under profiles/."""
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')

NAMES = {
    'c3b_ka': 'atm_moment_kernel (K_A), batch of 8 distinct C3 cubes',
    'c3_ka': 'atm_moment_kernel (K_A), one C3 field',
    'c3b_kb': 'legendre_moment_kernel (K_B), batch of 8 C3 fields',
    'c4_ka': 'pressure_moment_kernel (K_A pressure), C4 721x1440 LMAX 360',
    'c4_kb': 'legendre_moment_kernel (K_B), C4',
    'hyp': 'hypsometric_kernel<4>, 37x721x1440 float32',
    'c2_fused': 'points_fused_kernel<false>, C2 1e5 points LMAX 60',
}


def summary(rep):
    return subprocess.run(['python', os.path.join(ROOT, 'bench_micro', 'ncu_summary.py'), rep], capture_output=True,
                          text=True).stdout.rstrip()


def raw(rep):
    o = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(o.splitlines()))
    return {k: (rows[1][i], rows[2][i]) for i, k in enumerate(rows[0])}


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
    h = rows[0]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    return [(r[ki], float(r[vi])) for r in rows[1:] if len(r) > vi]


def main():
    out = ['# Round 2 ncu summaries (ncu --set full --clock-control none, one capture per kernel; bench_micro/r2_profiles.sh)',
           '# per-launch times under ncu are cold-cache and serialised; the timed numbers are in r02_bench_n1.json', '']
    for k, v in NAMES.items():
        out += ['== %s' % v, summary(os.path.join(OUT, 'r2f_%s.ncu-rep' % k)), '']
    open(os.path.join(PROF, 'r02_kernels_ncu.txt'), 'w').write('\n'.join(out))

    txt = ['# Round 2 launch lists (ncu --metrics gpu__time_duration.sum --clock-control none); times in ns, cold-cache and serialised', '']
    for c, desc in (('c3b', 'C3 atmosphere, batch of 8 distinct cubes (one call = memset + 5 launches)'),
                    ('c4', 'C4 pressure 721x1440 LMAX 360'), ('c1', 'C1 pressure 12 x 181x360 LMAX 60'),
                    ('c2', 'C2 1e5 points LMAX 60')):
        L = launches(os.path.join(OUT, 'r2f_launches_%s.csv' % c))
        txt.append('== %s' % desc)
        for k, v in (L[-12:] if c != 'c2' else L[-8:]):
            txt.append('%-90s %10.0f' % (k[:90], v))
        txt.append('')
    agg = {}
    for k, v in launches(os.path.join(OUT, 'r2f_bench_launches.csv')):
        a = agg.setdefault(k.split('(')[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    txt.append('== bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --c5-fields 16 (first 400 launches): share of device time per kernel')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        txt.append('%-70s n=%4d  %10.3f ms  %5.1f %%' % (k, n, t / 1e6, 100 * t / tot))
    open(os.path.join(PROF, 'r02_bench_launches.txt'), 'w').write('\n'.join(txt) + '\n')

    r = raw(os.path.join(OUT, 'r2f_c3b_ka.ncu-rep'))

    def val(k):
        u, v = r[k]
        return float(v.replace(',', '')) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1e-3, 'us': 1e-6,
                                             'ns': 1e-9}.get(u, 1)
    rec = {'kernel': 'atm_moment_kernel<BC05,double>',
           'capture': 'ncu --set full --clock-control none, bench_micro/prof_case.py c3b (8 distinct 37x721x1440 fp64 cubes, LMAX 60)',
           'fields_per_launch': 8, 'gpu_time_s': val('gpu__time_duration.sum'),
           'dram_bytes_read': val('dram__bytes_read.sum'), 'dram_bytes_write': val('dram__bytes_write.sum'),
           'algorithmic_bytes': 8 * 2 * 8 * 37 * 721 * 1440,
           'pipe_fp64_pct': val('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
           'pipe_dmma_pct': val('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'),
           'issue_active_pct': val('smsp__issue_active.avg.pct_of_peak_sustained_active'),
           'warps_active_pct': val('sm__warps_active.avg.pct_of_peak_sustained_active'),
           'registers_per_thread': val('launch__registers_per_thread')}
    rec['pipe_fp64_plus_dmma_pct'] = rec['pipe_fp64_pct'] + rec['pipe_dmma_pct']
    json.dump(rec, open(os.path.join(PROF, 'r02_atm_moment_batch8.json'), 'w'), indent=1)
    shutil.copy(os.path.join(OUT, 'r2f_bench_n1.json'), os.path.join(PROF, 'r02_bench_n1.json'))
    print(json.dumps(rec))


if __name__ == '__main__':
    main()
