"""
Time-sharded multi-GPU driver: one process per GPU (torchrun), independent time steps split
across ranks, one final gather of the coefficient blocks (NCCL over NVLink on the GPU box, gloo
in the CPU tests).

The reference processes time steps in a serial loop with no cross-iteration state
(``reanalysis/reanalysis_atmospheric_harmonics.py:355-421``), so the path shards without any
data-path collective; the only exchange is the gather of (T, 2, LMAX+1, MMAX+1) coefficients.
"""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ['shard_range', 'gather_coefficients', 'atmosphere_stokes_sharded']


def shard_range(nfields, rank, world_size):
    """Contiguous block of time steps owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(nfields), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_coefficients(local_block, nfields, group=None):
    """
    All-gathers per-rank coefficient blocks (T_local, ...) into the full (T, ...) tensor, in time
    order, on every rank.  Blocks may differ in length by one: they are padded to a common
    length for ``all_gather_into_tensor`` and trimmed afterwards.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return local_block
    world = dist.get_world_size(group)
    counts = [shard_range(nfields, r, world)[1] - shard_range(nfields, r, world)[0] for r in range(world)]
    tmax = max(counts)
    tail = tuple(local_block.shape[1:])
    padded = local_block.new_zeros((tmax,) + tail)
    padded[:local_block.shape[0]] = local_block
    out = local_block.new_empty((world * tmax,) + tail)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    parts = [out[r * tmax:r * tmax + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def atmosphere_stokes_sharded(GPH, pressure, lon, lat, field_scale, compute=None, group=None, **kwargs):
    """
    Fields t = 0..T-1 derived from one base cube (``field_scale`` (T, 2), see
    ``atmosphere_stokes_batch``) are split over the ranks; returns the gathered (clm, slm) of
    shape (T, LMAX+1, MMAX+1) as tensors on the compute device.  ``compute`` defaults to the
    CUDA operator; the CPU tests inject a host implementation to exercise the sharding logic.
    """
    field_scale = np.asarray(field_scale, dtype=np.float64)
    nfields = field_scale.shape[0]
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    t0, t1 = shard_range(nfields, rank, world)
    if compute is None:
        from .operators import atmosphere_stokes_batch
        compute = lambda fs: atmosphere_stokes_batch(GPH, pressure, lon, lat, field_scale=fs, as_numpy=False,
                                                     **kwargs)
    if t1 > t0:
        clm, slm = compute(field_scale[t0:t1])
        local = torch.stack([clm, slm], dim=1)          # (T_local, 2, L+1, M+1)
    else:
        local = None
    if world == 1:
        return local[:, 0], local[:, 1]
    if local is None:
        # ranks beyond the number of fields still take part in the collective
        ref = compute(field_scale[:1])
        local = torch.stack(ref, dim=1)[:0]
    full = gather_coefficients(local, nfields, group=group)
    return full[:, 0], full[:, 1]
