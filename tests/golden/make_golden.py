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


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--big':
        # BASELINE-size cases (minutes of CPU each): all of them, or the ones named after --big
        main(sys.argv[2:] or None, golden_cases.BIG_CASES)
    else:
        main()
        hypsometric()
