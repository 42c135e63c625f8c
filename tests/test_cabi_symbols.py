"""CPU: the C-ABI library loads and exports every symbol include/mhb200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    from model_harmonics_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'mhb200.h')).read()
    declared = set(re.findall(r'\b(mhb200_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 10
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), 'symbol %s missing from libmhb200.so' % name
    # the ctypes prototype table covers the header one to one
    assert declared == set(_lib.PROTOTYPES)


def test_no_cpu_fallback_without_device():
    import numpy as np
    import pytest
    import torch
    import model_harmonics_b200 as mh
    from model_harmonics_b200 import _lib
    assert _lib.lib().mhb200_version() >= 100
    if torch.cuda.is_available():
        pytest.skip('device present')
    assert _lib.lib().mhb200_device_check(0) != 0
    with pytest.raises(RuntimeError):
        mh.gen_pressure_stokes(np.zeros((3, 4)), 1.0, 1.0, np.arange(4.0), np.arange(3.0), LMAX=2,
                               LOVE=(np.zeros(3),) * 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'model_harmonics_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'import oracle' not in src and 'from oracle' not in src, fn
