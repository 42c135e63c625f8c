"""
Parity metrics.  Primary gate: max|X_a - X_b| / max|X_b| over clm and slm
jointly <= 1e-12.  Structural checks are exact.
"""
import numpy as np

TOL = 1.0e-12


def rel_to_max(a_clm, a_slm, b_clm, b_slm):
    scale = max(np.abs(b_clm).max(), np.abs(b_slm).max())
    err = max(np.abs(a_clm - b_clm).max(), np.abs(a_slm - b_slm).max())
    return err / scale


def assert_structure(clm, slm, lmax, mmax):
    assert clm.shape == (lmax + 1, mmax + 1) and slm.shape == (lmax + 1, mmax + 1)
    assert clm.dtype == np.float64 and slm.dtype == np.float64
    assert np.all(slm[:, 0] == 0.0)
    iu = np.triu_indices(lmax + 1, k=1, m=mmax + 1)
    assert np.all(clm[iu] == 0.0) and np.all(slm[iu] == 0.0)
    assert np.all(np.isfinite(clm)) and np.all(np.isfinite(slm))


def assert_parity(Y, clm_ref, slm_ref, tol=TOL, what=''):
    e = rel_to_max(Y.clm, Y.slm, clm_ref, slm_ref)
    assert e <= tol, '%s: relative-to-max error %.3e > %.1e' % (what, e, tol)
    return e
