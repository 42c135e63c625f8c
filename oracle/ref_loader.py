"""
Loads the reference hot-path files UNMODIFIED from ``/root/reference`` (TEST
INFRASTRUCTURE -- see ``oracle/__init__.py``).

``import model_harmonics`` fails in this image (its ``__init__`` pulls netCDF4, h5py,
pyproj, gravity_toolkit ...), but the three hot-path modules import only ``numpy``,
``gravity_toolkit`` and ``model_harmonics.datum``.  Registering the stand-in as
``gravity_toolkit`` and a stub package whose ``__path__`` points at the reference tree
(so the reference ``__init__`` never runs) lets them execute byte-for-byte as shipped.

The reference tree exists only in the build container; on the GPU box ``available()``
is False and the committed fixtures under ``tests/golden/`` stand in for it.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('MHB200_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'model_harmonics', 'gen_pressure_stokes.py'))


def load():
    """Returns a namespace with the reference's gen_* functions and ``datum`` class."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    try:
        import gravity_toolkit  # noqa: F401  (real package wins when importable)
        real = not getattr(gravity_toolkit, '_mhb200_standin', False)
    except ImportError:
        from oracle import gravtk_standin as g
        mod = types.ModuleType('gravity_toolkit')
        mod.plm_holmes, mod.legendre = g.plm_holmes, g.legendre
        mod.units, mod.harmonics = g.units, g.harmonics
        mod._mhb200_standin = True
        sys.modules['gravity_toolkit'] = mod
        real = False
    if 'model_harmonics' not in sys.modules or \
            not getattr(sys.modules['model_harmonics'], '_mhb200_stub', False):
        pkg = types.ModuleType('model_harmonics')
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'model_harmonics')]
        pkg._mhb200_stub = True
        sys.modules['model_harmonics'] = pkg
    ns = types.SimpleNamespace()
    ns.gen_pressure_stokes = importlib.import_module(
        'model_harmonics.gen_pressure_stokes').gen_pressure_stokes
    ns.gen_atmosphere_stokes = importlib.import_module(
        'model_harmonics.gen_atmosphere_stokes').gen_atmosphere_stokes
    ns.gen_point_pressure = importlib.import_module(
        'model_harmonics.gen_point_pressure').gen_point_pressure
    ns.datum = importlib.import_module('model_harmonics.datum').datum
    ns.real_gravity_toolkit = real
    return ns
