"""
model_harmonics_b200 -- B200-native (sm_100a) fields -> Stokes coefficients.

Drop-in for the harmonic-analysis operators of tsutterley/model-harmonics
(``model_harmonics/__init__.py:19-21``): ``gen_pressure_stokes``, ``gen_atmosphere_stokes``,
``gen_point_pressure`` keep the reference's call surface and return ``harmonics`` objects;
the arithmetic runs in hand-written CUDA kernels behind the C ABI in ``include/mhb200.h``.
There is no CPU fallback.
"""
from .operators import (gen_pressure_stokes, gen_atmosphere_stokes, gen_point_pressure, gen_stokes,
                        pressure_stokes_batch, atmosphere_stokes_batch, point_pressure_batch, stokes_batch,
                        gen_point_pressure_tiles)
from .hypsometric import geopotential_levels
from ._compat import harmonics, units

__version__ = '0.1.0'
__all__ = ['gen_pressure_stokes', 'gen_atmosphere_stokes', 'gen_point_pressure', 'gen_stokes',
           'pressure_stokes_batch', 'atmosphere_stokes_batch', 'point_pressure_batch', 'stokes_batch',
           'gen_point_pressure_tiles', 'geopotential_levels', 'harmonics', 'units']
