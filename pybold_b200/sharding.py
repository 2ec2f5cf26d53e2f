"""Voxel-range sharding across the GPUs of one box.

Voxels are independent problems, so the solve needs no collective: rank r of R owns the
contiguous rows ``[floor(r V / R), floor((r + 1) V / R))`` of ``y[V, T]`` and the only exchange is
the final gather of the per-rank output slabs (``torch.distributed``: NCCL over NVLink on the
GPU box, gloo in the CPU tests).  Nothing here computes; it only partitions and gathers.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def voxel_range(n_voxels, rank, world_size):
    """Half-open row range owned by ``rank`` (balanced to within one voxel)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return (rank * n_voxels) // world_size, ((rank + 1) * n_voxels) // world_size


def gather_rows(local, n_voxels, group=None):
    """All-gather row slabs of unequal height into the full ``[V, ...]`` tensor on every rank.

    Slabs are padded to the largest height so that one ``all_gather_into_tensor`` suffices.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    heights = [voxel_range(n_voxels, r, world)[1] - voxel_range(n_voxels, r, world)[0]
               for r in range(world)]
    if local.shape[0] != heights[rank]:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local.shape[0], heights[rank]))
    hmax = max(heights)
    tail = tuple(local.shape[1:])
    padded = local.new_zeros((hmax,) + tail)
    padded[:local.shape[0]] = local
    full = local.new_empty((world * hmax,) + tail)
    dist.all_gather_into_tensor(full, padded.contiguous(), group=group)
    full = full.reshape((world, hmax) + tail)
    return torch.cat([full[r, :heights[r]] for r in range(world)], dim=0)


def bd_sharded(y_full_or_local, n_voxels, solve, gather=True, group=None):
    """Run ``solve(y_local) -> dict of [v_local, ...] tensors`` on this rank's rows.

    ``y_full_or_local`` is either the full ``[V, T]`` matrix (every rank slices its rows) or this
    rank's slab already.  With ``gather`` the outputs are all-gathered to ``[V, ...]`` everywhere.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = voxel_range(n_voxels, rank, world)
    y_local = y_full_or_local[lo:hi] if y_full_or_local.shape[0] == n_voxels and world > 1 \
        else y_full_or_local
    out = solve(y_local)
    if not gather or world == 1:
        return out
    return {k: gather_rows(v, n_voxels, group) for k, v in out.items()}
