"""
Parity tests proper (GPU): the CUDA path, called through the Python drop-in operators and the
C ABI, against (a) the committed golden fixtures produced by the unmodified reference files and
(b) the NumPy oracle run on the same seeded inputs.  Gate: max|delta| / max|ref| <= 1e-12 over
clm and slm jointly; structural properties exact.
"""
import os

import numpy as np
import pytest

import golden_cases
from parity import TOL, rel_to_max, assert_structure

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _ops():
    import model_harmonics_b200 as mh
    return mh


@pytest.mark.parametrize('name', sorted(golden_cases.CASES))
def test_cuda_matches_golden_and_oracle(name):
    from oracle import stokes_oracle
    mh = _ops()
    func, args, kwargs = golden_cases.build(name)
    gold = np.load(os.path.join(GOLDEN, name + '.npz'))
    Y = getattr(mh, func)(*args, **kwargs)
    lmax, mmax = int(gold['lmax']), int(gold['mmax'])
    assert int(Y.lmax) == lmax and int(Y.mmax) == mmax
    assert_structure(Y.clm, Y.slm, lmax, mmax)
    e_gold = rel_to_max(Y.clm, Y.slm, gold['clm'], gold['slm'])
    O = getattr(stokes_oracle, func)(*args, **kwargs)
    e_orc = rel_to_max(Y.clm, Y.slm, O.clm, O.slm)
    print('%s: vs golden %.2e, vs oracle %.2e' % (name, e_gold, e_orc))
    assert e_gold <= TOL, 'vs golden: %.3e' % e_gold
    assert e_orc <= TOL, 'vs oracle: %.3e' % e_orc


def test_inputs_not_modified_and_rerun_bitwise():
    mh = _ops()
    func, args, kwargs = golden_cases.build('atmosphere_bc05')
    before = [np.copy(a) for a in args]
    A = mh.gen_atmosphere_stokes(*args, **kwargs)
    B = mh.gen_atmosphere_stokes(*args, **kwargs)
    for a, b in zip(args, before):
        assert np.array_equal(a, b)
    # fixed-order reductions: run-to-run bit reproducible
    assert np.array_equal(A.clm, B.clm) and np.array_equal(A.slm, B.slm)


def test_plm_argument_paths(monkeypatch):
    """PLM given: standard table -> recomputed on the fly; foreign table -> used as given."""
    from oracle import stokes_oracle
    from oracle.gravtk_standin import plm_holmes
    mh = _ops()
    func, args, kwargs = golden_cases.build('pressure_options')
    lat = args[4]
    L = kwargs['LMAX']
    PLM, _ = plm_holmes(L, np.cos(np.radians(90.0 - lat)))
    base = mh.gen_pressure_stokes(*args, **kwargs)
    with_tab = mh.gen_pressure_stokes(*args, PLM=PLM, **kwargs)
    assert rel_to_max(with_tab.clm, with_tab.slm, base.clm, base.slm) <= 1e-14
    monkeypatch.setenv('MHB200_PLM_MODE', 'table')
    forced = mh.gen_pressure_stokes(*args, PLM=PLM.copy(), **kwargs)
    assert rel_to_max(forced.clm, forced.slm, base.clm, base.slm) <= TOL
    monkeypatch.delenv('MHB200_PLM_MODE')
    # a table from different colatitudes (e.g. geocentric) must be honoured
    PLM2, _ = plm_holmes(L, np.cos(np.radians(90.0 - 0.995 * lat)))
    got = mh.gen_pressure_stokes(*args, PLM=PLM2, **kwargs)
    want = stokes_oracle.gen_pressure_stokes(*args, PLM=PLM2, **kwargs)
    assert rel_to_max(got.clm, got.slm, want.clm, want.slm) <= TOL


def test_float32_cubes_match_float64_of_same_values():
    mh = _ops()
    # nlon = 120: float32 and float64 rows are both 16-byte multiples, so both storage types take the same
    # (TMA-fed) path and the widening is exact -> bit-identical results
    func, args, kwargs = golden_cases.build('atmosphere_weight')
    GPH32, dP32 = args[0].astype(np.float32), args[1].astype(np.float32)
    A = mh.gen_atmosphere_stokes(GPH32, dP32, *args[2:], **kwargs)
    B = mh.gen_atmosphere_stokes(GPH32.astype(np.float64), dP32.astype(np.float64), *args[2:], **kwargs)
    assert np.array_equal(A.clm, B.clm) and np.array_equal(A.slm, B.slm)
    # nlon = 90: float32 rows are not 16-byte multiples -> direct kernels; same values to rounding
    func, args, kwargs = golden_cases.build('atmosphere_levels37')
    GPH32, dP32 = args[0].astype(np.float32), args[1].astype(np.float32)
    A = mh.gen_atmosphere_stokes(GPH32, dP32, *args[2:], **kwargs)
    B = mh.gen_atmosphere_stokes(GPH32.astype(np.float64), dP32.astype(np.float64), *args[2:], **kwargs)
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= 1e-14


def test_batch_entry_points_match_single_calls():
    import torch
    from model_harmonics_b200 import testing as syn
    mh = _ops()
    # pressure: static geometry, 5 fields
    lon, lat = syn.grid_axes(46, 90)
    G, R = syn.surface_geometry(lon, lat, seed=3)
    P = syn.pressure_fields(46, 90, 5, seed=10)
    LOVE = syn.love_numbers(30)
    clm, slm = mh.pressure_stokes_batch(P, G, R, lon, lat, LMAX=30, LOVE=LOVE)
    for t in range(5):
        Y = mh.gen_pressure_stokes(P[t], G, R, lon, lat, LMAX=30, LOVE=LOVE)
        # a different batch size is a different static schedule, i.e. a different (fixed) summation order
        assert rel_to_max(clm[t], slm[t], Y.clm, Y.slm) <= 1e-14
    # atmosphere: shared cube + per-field scale factors == explicit scaled cubes
    GPH, dP, geoid = syn.atmosphere_cube(6, 46, 90, seed=5)
    scales = np.array([syn.atmosphere_field_scalars(t) for t in (0, 7, 123)])
    kw = dict(LMAX=30, ELLIPSOID='WGS84', GEOID=geoid, LOVE=LOVE)
    clm, slm = mh.atmosphere_stokes_batch(GPH, dP, lon, lat, field_scale=scales, **kw)
    for i in range(3):
        Y = mh.gen_atmosphere_stokes(GPH * scales[i, 0], dP * scales[i, 1], lon, lat, **kw)
        assert rel_to_max(clm[i], slm[i], Y.clm, Y.slm) <= 1e-14
    # device-resident inputs are used in place
    Pd = torch.from_numpy(P).cuda()
    c2, s2 = mh.pressure_stokes_batch(Pd, torch.from_numpy(G).cuda(), torch.from_numpy(R).cuda(), lon, lat,
                                      LMAX=30, LOVE=LOVE, as_numpy=False)
    assert c2.is_cuda and np.array_equal(c2.cpu().numpy(), mh.pressure_stokes_batch(P, G, R, lon, lat, LMAX=30, LOVE=LOVE)[0])


def test_error_conventions():
    mh = _ops()
    func, args, kwargs = golden_cases.build('atmosphere_weight')
    bad = dict(kwargs)
    bad['METHOD'] = 'XX'
    with pytest.raises(ValueError):
        mh.gen_atmosphere_stokes(*args, **bad)
    bad = dict(kwargs)
    bad['ELLIPSOID'] = None
    with pytest.raises(AttributeError):
        mh.gen_atmosphere_stokes(*args, **bad)
    bad = dict(kwargs)
    bad['ELLIPSOID'] = 'NOPE'
    with pytest.raises(AssertionError):
        mh.gen_atmosphere_stokes(*args, **bad)
    bad = dict(kwargs)
    bad['GEOID'] = None
    with pytest.raises(TypeError):
        mh.gen_atmosphere_stokes(*args, **bad)
    bad = dict(kwargs)
    bad['LOVE'] = None
    with pytest.raises(TypeError):
        mh.gen_atmosphere_stokes(*args, **bad)
    func, args, kwargs = golden_cases.build('points_mmax0')
    bad = dict(kwargs)
    bad['AREA'] = None
    with pytest.raises(AttributeError):
        mh.gen_point_pressure(*args, **bad)


def test_host_band_pipeline_matches_device_path():
    """NumPy cubes take the latitude-band pipelined C-ABI entry; CUDA tensors take the device entry."""
    import torch
    mh = _ops()
    func, args, kwargs = golden_cases.build('atmosphere_bc05')
    GPH, dP = args[0], args[1]
    for dt in (np.float64, np.float32):
        g, d = np.ascontiguousarray(GPH.astype(dt)), np.ascontiguousarray(dP.astype(dt))
        host = mh.gen_atmosphere_stokes(g, d, *args[2:], **kwargs)
        dev = mh.gen_atmosphere_stokes(torch.from_numpy(g).cuda(), torch.from_numpy(d).cuda(), *args[2:], **kwargs)
        # different static schedules (8 bands vs one launch): fixed but different summation orders
        assert rel_to_max(host.clm, host.slm, dev.clm, dev.slm) <= 1e-14
    # non-contiguous and masked inputs go through the same entry
    gm = np.ma.masked_array(GPH, mask=np.zeros(GPH.shape, dtype=bool))
    A = mh.gen_atmosphere_stokes(gm, dP[:, :, ::1], *args[2:], **kwargs)
    B = mh.gen_atmosphere_stokes(GPH, dP, *args[2:], **kwargs)
    assert np.array_equal(A.clm, B.clm) and np.array_equal(A.slm, B.slm)


@pytest.mark.parametrize('name', ['pressure_options', 'pressure_mean_field', 'pressure_c1'])
def test_condition_aware_bound_and_truth(name):
    """
    Secondary gate: |delta_lm| <= 1e-12 * A_lm per coefficient, A_lm = the sum of the
    absolute values of the terms of that coefficient.  Informational: literal per-coefficient
    relative error of the CUDA path and of the reference against an extended-precision evaluation.
    """
    from oracle import error_scale as es, stokes_oracle
    mh = _ops()
    func, args, kwargs = golden_cases.build(name)
    A = es.pressure_error_scale(*args, **kwargs)
    tc, ts = es.pressure_truth(*args, **kwargs)
    tc, ts = tc.astype(np.float64), ts.astype(np.float64)
    Y = mh.gen_pressure_stokes(*args, **kwargs)
    O = stokes_oracle.gen_pressure_stokes(*args, **kwargs)
    ok = A > 0
    for got, ref in ((Y.clm, O.clm), (Y.slm, O.slm)):
        assert np.all(np.abs(got - ref)[ok] <= 1e-12 * A[ok])
    nz = ok & (np.abs(tc) > 0)
    lit_gpu = (np.abs(Y.clm - tc)[nz] / np.abs(tc[nz])).max()
    lit_ref = (np.abs(O.clm - tc)[nz] / np.abs(tc[nz])).max()
    cond_gpu = (np.abs(Y.clm - tc)[ok] / A[ok]).max()
    cond_ref = (np.abs(O.clm - tc)[ok] / A[ok]).max()
    print('%s: vs extended-precision truth -- |d|/A: cuda %.2e, reference %.2e; literal per-coefficient: '
          'cuda %.2e, reference %.2e' % (name, cond_gpu, cond_ref, lit_gpu, lit_ref))
    # the CUDA path is as close to the truth as the reference is (within a small factor)
    assert cond_gpu <= max(4.0 * cond_ref, 1e-14)
