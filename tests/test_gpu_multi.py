"""
Multi-GPU (NCCL) test of the time-sharded driver: runs only when >= 2 CUDA devices are visible
(`gpurun --gpus 2`); the CPU suite covers the same logic with gloo (tests/test_sharding_gloo.py).
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import model_harmonics_b200 as mh
    from model_harmonics_b200 import batch, testing as syn
    L = 30
    T = 11                                     # uneven over 2 ranks
    lon, lat = syn.grid_axes(46, 90)
    GPH, dP, geoid = syn.atmosphere_cube(6, 46, 90, seed=50)
    scales = np.array([syn.atmosphere_field_scalars(t) for t in range(T)])
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(L))
    g, d = torch.from_numpy(GPH).cuda(), torch.from_numpy(dP).cuda()
    clm, slm = batch.atmosphere_stokes_sharded(g, d, lon, lat, scales, **kw)
    assert clm.shape[0] == T and clm.is_cuda
    if rank == 0:
        want_c, want_s = mh.atmosphere_stokes_batch(g, d, lon, lat, field_scale=scales, **kw)
        sc = np.abs(want_c).max()
        err = max(np.abs(clm.cpu().numpy() - want_c).max(), np.abs(slm.cpu().numpy() - want_s).max()) / sc
        open(os.path.join(outdir, 'err'), 'w').write(repr(float(err)))
    dist.barrier()
    dist.destroy_process_group()


def test_time_sharded_atmosphere_under_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 CUDA devices (gpurun --gpus 2)')
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    # different batch sizes are different (fixed) summation orders: agreement to rounding
    assert float(open(os.path.join(str(tmp_path), 'err')).read()) <= 1e-14
