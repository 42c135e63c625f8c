"""The reference arm of bench.py runs without a GPU: check the output contract on it (exactly one JSON
line on stdout, the keys the driver reads)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ)
    env.pop('RANK', None)
    env.pop('WORLD_SIZE', None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '0'], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['metric'] == 'fields_to_stokes_per_second' and d['unit'] == 'fields/s'
    assert d['steps'] == 1 and d['n_gpus'] == 1 and d['higher_is_better'] is True
    assert d['dtype'] == 'f64' and d['vs_baseline'] is None
    assert d['config']['workload'] == 'atmosphere_c3'
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['sample'] and cb['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'fields/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['value'] > 0.0
