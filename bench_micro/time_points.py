"""C2 timing (1e5 points -> LMAX 60), device-resident, CUDA events; parity against the reference golden."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import model_harmonics_b200 as mh
from model_harmonics_b200 import testing as syn
LOVE = syn.love_numbers(60)
pts = syn.point_cloud(100000)
dv = [torch.from_numpy(a).cuda() for a in pts]
f = lambda: mh.point_pressure_batch(dv[0][None], dv[1], dv[2], dv[3], dv[4], AREA=dv[5], LMAX=60, LOVE=LOVE, as_numpy=False)
c, s = f()
gold = np.load('tests/golden/points_c2.npz')
sc = max(np.abs(gold['clm']).max(), np.abs(gold['slm']).max())
err = max(np.abs(c[0].cpu().numpy() - gold['clm']).max(), np.abs(s[0].cpu().numpy() - gold['slm']).max()) / sc
for _ in range(3): f()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in ev:
    a.record(); f(); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
print('C2 ms median %.4f best %.4f  rel-to-max vs golden %.2e' % (ms[10], ms[0], err))
# kernel-only times from the library's profile hook, if it is on
try:
    from model_harmonics_b200 import _lib
    L = _lib.lib()
    if hasattr(L, 'mhb200_profile_enable'):
        pass
except Exception:
    pass
# back-to-back stream of 20 calls
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): f()
b.record(); torch.cuda.synchronize()
print('C2 back-to-back ms per call %.4f' % (a.elapsed_time(b) / 20))
