"""
CPU (gloo, world_size 2): the time-shard + gather logic of the multi-GPU driver.  The compute
step is injected (NumPy oracle) because there is no GPU here; on the GPU box the same code runs
with the CUDA operator and the NCCL backend (tests/test_gpu_parity.py covers world_size 1).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from model_harmonics_b200.batch import shard_range, gather_coefficients, atmosphere_stokes_sharded


def test_shard_range_partitions():
    for T in (1, 2, 5, 7, 480):
        for W in (1, 2, 3, 4, 8):
            spans = [shard_range(T, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == T
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, T, outdir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import stokes_oracle
    from model_harmonics_b200 import testing as syn
    L = 8
    lon, lat = syn.grid_axes(13, 24)
    GPH, dP, geoid = syn.atmosphere_cube(3, 13, 24, seed=2)
    scales = np.array([syn.atmosphere_field_scalars(t) for t in range(T)])
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(L))

    def compute(fs):
        Y = [stokes_oracle.gen_atmosphere_stokes(GPH * a, dP * b, lon, lat, **kw) for a, b in fs]
        return (torch.from_numpy(np.stack([y.clm for y in Y])), torch.from_numpy(np.stack([y.slm for y in Y])))

    clm, slm = atmosphere_stokes_sharded(GPH, dP, lon, lat, scales, compute=compute)
    if rank == 0:
        want = compute(scales)
        ok = torch.equal(clm, want[0]) and torch.equal(slm, want[1]) and clm.shape[0] == T
        open(os.path.join(outdir, 'ok'), 'w').write('1' if ok else '0')
    # uneven blocks through the bare gather
    t0, t1 = shard_range(T, rank, world)
    block = torch.arange(t0, t1, dtype=torch.float64).reshape(-1, 1, 1).repeat(1, 2, 3)
    full = gather_coefficients(block, T)
    assert torch.equal(full[:, 0, 0], torch.arange(T, dtype=torch.float64))
    dist.destroy_process_group()


@pytest.mark.parametrize('T', [5, 1])
def test_time_shard_and_gather_world2(tmp_path, T):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, T, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), 'ok')).read() == '1'
