#!/usr/bin/env python
"""
Generates ``tests/golden/*.npz`` by running the UNMODIFIED reference hot-path files from
``/root/reference`` (through ``oracle/ref_loader.py``) on seeded synthetic inputs.

Run in the build container only (the reference tree does not travel):
    python tests/golden/make_golden.py                 # the small cases
    python tests/golden/make_golden.py --big [names]   # BASELINE-size cases C2/C3/C4/C5 (minutes each)
Each fixture stores the case parameters, a SHA-256 of the regenerated inputs, and the
reference's clm/slm.  Inputs are not stored: ``tests/golden_cases.py`` rebuilds them from
the same seeds.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import ref_loader          # noqa: E402
import golden_cases                    # noqa: E402


def main(names=None, table=None):
    ref = ref_loader.load()
    outdir = os.path.dirname(os.path.abspath(__file__))
    table = golden_cases.CASES if table is None else table
    for name in (names or table):
        func, args, kwargs = table[name]()
        Y = getattr(ref, func)(*args, **kwargs)
        path = os.path.join(outdir, name + '.npz')
        np.savez_compressed(path, clm=Y.clm, slm=Y.slm, lmax=int(Y.lmax), mmax=int(Y.mmax),
                            input_sha256=golden_cases.digest(args, kwargs), func=func,
                            real_gravity_toolkit=bool(ref.real_gravity_toolkit))
        print('%-28s %-22s max|clm|=%.3e  %s' % (name, func, np.abs(Y.clm).max(), Y.clm.shape))


def hypsometric():
    """tests/golden/hypsometric.npz: the reference's own loop lines executed on seeded inputs."""
    from oracle import hypsometric_oracle as ho
    case = dict(nlev=11, nlat=10, nlon=18, seed=20)
    Z, D = ho.reference_loop(*ho.synthetic_model_levels(**case))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hypsometric.npz'), Z=Z, DIFF=D,
                        source='reference lines 349-383 executed by oracle/hypsometric_oracle.reference_loop')


def driver_ops():
    """tests/golden/driver_ops.npz: the drivers' own pre-step lines and tile loop executed on seeded inputs
    (oracle/driver_ops_oracle.py reference_* functions; inputs rebuilt by tests/golden_cases.driver_case)."""
    from oracle import driver_ops_oracle as oo
    from oracle.gravtk_standin import harmonics
    ref = ref_loader.load()
    lon, lat, pressure, mean_p, ii, jj, dP, weight, tiles, L, LOVE = golden_cases.driver_case()
    out = {}
    for name, w in (('plain', None), ('weighted', weight)):
        out['pressure_' + name] = np.stack([oo.reference_pressure_step(pressure, mean_p, t, lat, ii, jj, weight=w)
                                            for t in range(pressure.shape[0])])
        out['levels_' + name] = oo.reference_atmosphere_step(dP, lat, ii, jj, weight=w)
    Y = oo.reference_tile_accumulation(*tiles, L, L, LOVE, ref.gen_point_pressure, harmonics)
    out['tiles_clm'], out['tiles_slm'] = Y.clm, Y.slm
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'driver_ops.npz'),
                        source='reference driver lines executed by oracle/driver_ops_oracle.py', **out)
    print('driver_ops', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--big':
        # BASELINE-size cases (minutes of CPU each): all of them, or the ones named after --big
        main(sys.argv[2:] or None, golden_cases.BIG_CASES)
    else:
        main()
        hypsometric()
        driver_ops()
