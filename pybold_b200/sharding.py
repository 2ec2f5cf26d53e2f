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


def row_heights(n_voxels, world_size):
    return [voxel_range(n_voxels, r, world_size)[1] - voxel_range(n_voxels, r, world_size)[0]
            for r in range(world_size)]


def gather_rows(local, n_voxels, group=None, out=None):
    """All-gather the per-rank row slabs into the full ``[V, ...]`` tensor on every rank.

    Equal slabs (the usual case: V divisible by the world size) are ONE ``all_gather_into_tensor``
    straight into the result -- no padding, no concatenation, the result is written once.  Unequal
    slabs (they differ by at most one row) go as one grouped batch of sends / receives, every
    receive landing in its final rows.  ``out`` lets a caller reuse the result tensor.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    heights = row_heights(n_voxels, world)
    if local.shape[0] != heights[rank]:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local.shape[0], heights[rank]))
    tail = tuple(local.shape[1:])
    if out is None:
        out = local.new_empty((n_voxels,) + tail)
    elif tuple(out.shape) != (n_voxels,) + tail or out.dtype != local.dtype or not out.is_contiguous():
        raise ValueError("out must be a contiguous %s tensor of shape %s" % (local.dtype, (n_voxels,) + tail))
    local = local.contiguous()
    if world == 1:
        out.copy_(local)
        return out
    if min(heights) == max(heights):
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    ops = []
    for r in range(world):
        lo, hi = voxel_range(n_voxels, r, world)
        if r == rank:
            out[lo:hi].copy_(local)
            continue
        peer = dist.get_global_rank(group, r) if group is not None else r
        if heights[rank] > 0:
            ops.append(dist.P2POp(dist.isend, local, peer, group))
        if hi > lo:
            ops.append(dist.P2POp(dist.irecv, out[lo:hi], peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


ALL_OUTPUTS = ("x", "z", "diff_z", "h", "theta", "J", "r", "g")


def gather_outputs(out, n_voxels, keys=ALL_OUTPUTS, group=None, into=None):
    """Gather the named outputs of a ``bd_batch`` result dict; ``into`` (a dict from a previous call)
    reuses the full-size tensors."""
    res = {} if into is None else into
    for k in keys:
        res[k] = gather_rows(out[k], n_voxels, group, out=res.get(k))
    return res


class PeerGather:
    """All-gather of per-rank row slabs by COPY-ENGINE peer writes over NVLink.

    Every rank owns the full-size result tensors ``full[name]`` of shape ``[V, ...]`` and exports them once
    through CUDA IPC; a gather is then, on every rank, one ``cudaMemcpyAsync`` of its slab into its rows of
    every peer's tensor (device-to-device over NVLink / NVSwitch, executed by the copy engines: no SM is
    taken from the solver that is running the next step), followed by a one-element all-reduce that is
    stream-ordered after the copies on every rank -- when it completes, every slab has landed everywhere.
    Compared with ``all_gather_into_tensor`` on a side stream: an NCCL all-gather kernel that loses the race
    for SMs against the next persistent solve makes the kernels of ALL ranks spin for a whole step on the
    SMs they hold (measured on 8 GPUs: 642 instead of 590 ms per step); the copies need no SM at all and the
    fence kernel is one CTA.

    ``spec``: name -> (tail shape, dtype).  One process per GPU on ONE node, every GPU visible to every
    process under the same index (what ``torch.distributed.run`` does).

    ``mapping``: how a rank gets at its peers' tensors.  ``"symm"`` (default): the result tensors live in
    one symmetric-memory allocation (``torch.distributed._symmetric_memory``: cuMem handles exchanged at
    rendezvous), peer writes go over NVLink.  ``"ipc"``: legacy ``cudaIpc`` handles of ordinary tensors --
    works everywhere, also for two processes on one GPU (the test), but peer writes through such mappings
    were measured at PCIe speed on the NVSwitch box (25 GB/s).
    """

    def __init__(self, spec, n_voxels, device, group=None, mapping="symm"):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_voxels = n_voxels
        self.device = torch.device(device)
        self.lo, self.hi = voxel_range(n_voxels, self.rank, self.world)
        self.mapping = mapping
        self.peers = []
        err = None
        try:
            if mapping == "symm":
                self._map_symmetric(spec)
            else:
                self._map_ipc(spec)
            for r in range(self.world):                     # touch every mapping once
                for k in self.full:
                    if self.hi > self.lo:
                        self.peers[r][k][self.lo:self.lo + 1].copy_(self.full[k][self.lo:self.lo + 1])
            torch.cuda.synchronize(self.device)
        except Exception as exc:                            # no symmetric memory / IPC / peer access here
            err = exc
            if not hasattr(self, "full"):
                self.full = {}
        # every rank takes the same decision (this is also the barrier that ends the setup)
        self._nccl = dist.get_backend(group) == "nccl"
        ok = torch.tensor([0.0 if err is None else 1.0], device=self.device if self._nccl else "cpu")
        dist.all_reduce(ok, group=group)
        if float(ok) > 0:
            raise RuntimeError("PeerGather: CUDA IPC / peer access is not available on %d rank(s)%s"
                               % (int(ok), "" if err is None else ": %r" % (err,)))
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)

    def _map_ipc(self, spec):
        from torch.multiprocessing.reductions import rebuild_cuda_tensor, reduce_tensor
        self.full = {k: torch.empty((self.n_voxels,) + tuple(tail), dtype=dt, device=self.device)
                     for k, (tail, dt) in spec.items()}
        mine = {k: reduce_tensor(t)[1] for k, t in self.full.items()}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        for r in range(self.world):
            if r == self.rank:
                self.peers.append(self.full)
            else:
                self.peers.append({k: rebuild_cuda_tensor(*args) for k, args in everyone[r].items()})

    def _map_symmetric(self, spec):
        import torch.distributed._symmetric_memory as symm_mem
        # one allocation for all outputs, every output 256-byte aligned, viewed as bytes
        offsets, total = {}, 0
        for k, (tail, dt) in spec.items():
            n = self.n_voxels
            for d in tail:
                n *= d
            offsets[k] = total
            total += (n * torch.empty((), dtype=dt).element_size() + 255) // 256 * 256
        with torch.cuda.device(self.device):
            self._buf = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
            grp = self.group if self.group is not None else dist.group.WORLD
            self._hdl = symm_mem.rendezvous(self._buf, grp)

        def views(r):
            out = {}
            for k, (tail, dt) in spec.items():
                shape = (self.n_voxels,) + tuple(tail)
                esz = torch.empty((), dtype=dt).element_size()
                out[k] = self._hdl.get_buffer(r, shape, dt, offsets[k] // esz)
            return out

        self.full = views(self.rank)
        for r in range(self.world):
            self.peers.append(self.full if r == self.rank else views(r))

    def gather(self, local, keys=None):
        """Push this rank's slabs (``local[name]`` of shape ``[hi - lo, ...]``) to every rank, on the
        current stream; returns ``self.full`` (valid once the stream has passed the fence)."""
        keys = list(self.full) if keys is None else keys
        for k in keys:
            if local[k].shape[0] != self.hi - self.lo:
                raise ValueError("rank %d holds %d rows of %s, expected %d"
                                 % (self.rank, local[k].shape[0], k, self.hi - self.lo))
        from . import _lib
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            for step in range(self.world):
                r = (self.rank + step) % self.world      # start with the own copy, then round the ring
                for k in keys:
                    src = local[k].contiguous()
                    dst = self.peers[r][k][self.lo:self.hi]
                    # one cudaMemcpyAsync per slab, issued from this rank's device (torch's cross-device
                    # copy_ of an IPC mapping was measured at PCIe speed, 24 GB/s)
                    _lib.check(_lib.lib.pb_copy_async(dst.data_ptr(), src.data_ptr(),
                                                      src.numel() * src.element_size(), stream), "pb_copy_async")
        self.fence()
        return self.full

    def fence(self):
        if self._nccl:
            dist.all_reduce(self._flag, group=self.group)    # ordered after the copies on this stream
        else:
            torch.cuda.current_stream(self.device).synchronize()
            dist.barrier(group=self.group)


_peer_gatherers = {}


def _peer_gatherer(out, n_voxels, group, mapping):
    """The cached :class:`PeerGather` for outputs of these shapes (building one is a collective: symmetric
    allocation + rendezvous; every rank arrives here with the same shapes in the same call)."""
    spec = {k: (tuple(v.shape[1:]), v.dtype) for k, v in out.items()}
    dev = next(iter(out.values())).device
    key = (tuple((k, s, str(d)) for k, (s, d) in spec.items()), n_voxels, str(dev), id(group), mapping)
    pg = _peer_gatherers.get(key)
    if pg is None:
        if len(_peer_gatherers) >= 4:       # each one holds full-size result tensors
            _peer_gatherers.clear()
        pg = _peer_gatherers[key] = PeerGather(spec, n_voxels, dev, group, mapping=mapping)
    return pg


def bd_sharded(y_full_or_local, n_voxels, solve, gather=True, group=None, peer_mapping="symm"):
    """Run ``solve(y_local) -> dict of [v_local, ...] tensors`` on this rank's rows.

    ``y_full_or_local`` is either the full ``[V, T]`` matrix (every rank slices its rows) or this
    rank's slab already.  With ``gather`` the outputs are all-gathered to ``[V, ...]`` everywhere:
    ``True`` over the process group's backend (NCCL / gloo, :func:`gather_rows`), ``"peer"`` by copy-engine
    peer writes (:class:`PeerGather`, CUDA tensors on one node).  The peer path returns tensors OWNED by the
    cached gatherer: they are overwritten by the next call with the same shapes -- clone what must survive it.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = voxel_range(n_voxels, rank, world)
    y_local = y_full_or_local[lo:hi] if y_full_or_local.shape[0] == n_voxels and world > 1 \
        else y_full_or_local
    out = solve(y_local)
    if not gather or world == 1:
        return out
    if gather == "peer":
        pg = _peer_gatherer(out, n_voxels, group, peer_mapping)
        pg.fence()          # every rank is done with the previous result before anyone overwrites it
        return dict(pg.gather(out))
    return {k: gather_rows(v, n_voxels, group) for k, v in out.items()}
