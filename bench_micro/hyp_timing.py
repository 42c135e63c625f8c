import sys, ctypes
import numpy as np, torch
sys.path.insert(0, '.')
import model_harmonics_b200 as mh
from model_harmonics_b200 import testing as syn, _lib
T, Q, SP, zs, A, B, AI, BI = syn.model_level_inputs(37, 721, 1440, seed=11)
dv = [torch.from_numpy(a).cuda() for a in (T, Q, SP, zs)]
f = lambda: mh.geopotential_levels(dv[0], dv[1], dv[2], dv[3], A, B, AI, BI, divide_by=9.80665, as_numpy=False)
L = _lib.lib()
for _ in range(3): f()
torch.cuda.synchronize(); L.mhb200_profile_enable(1)
for _ in range(10): f()
n = ctypes.c_int(0); ms = L.mhb200_profile_mean_ms(ctypes.byref(n)); L.mhb200_profile_enable(0)
moved = 16 * 37 * 721 * 1440
print('kernel ms %.4f  GB/s %.0f  (n=%d)' % (ms, moved / ms / 1e6, n.value))
