"""World-size-2 test of the multi-GPU host logic on CPU (gloo): voxel-range partition and the
final gather.  The solve itself needs a GPU; here each rank runs a stand-in that tags its rows."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pybold_b200.sharding import ALL_OUTPUTS, bd_sharded, gather_outputs, gather_rows, voxel_range


def test_voxel_range_partitions_exactly():
    for V in (0, 1, 7, 100, 230000):
        for R in (1, 2, 3, 8):
            spans = [voxel_range(V, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == V
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        voxel_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, V, T, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(V * T, dtype=torch.float32).reshape(V, T)

        def solve(y_local):     # stand-in for bd_batch: per-row results that depend on the row only
            return {"z": y_local * 2.0, "theta": y_local[:, 0] + 0.5}

        out = bd_sharded(full, V, solve, gather=True)
        ok = torch.equal(out["z"], full * 2.0) and torch.equal(out["theta"], full[:, 0] + 0.5)
        lo, hi = voxel_range(V, rank, world)
        out2 = gather_rows(full[lo:hi].clone(), V)
        ok = ok and torch.equal(out2, full)
        # every output of the solve, into reused result tensors, twice
        local = {k: (full[lo:hi, :3] + i if k not in ("theta",) else full[lo:hi, 0] + i)
                 for i, k in enumerate(ALL_OUTPUTS)}
        into = None
        for _ in range(2):
            into = gather_outputs(local, V, into=into)
        for i, k in enumerate(ALL_OUTPUTS):
            want = full[:, :3] + i if k != "theta" else full[:, 0] + i
            ok = ok and torch.equal(into[k], want)
        try:
            gather_rows(full[lo:hi].clone(), V, out=torch.empty(V + 1, T))
            ok = False
        except ValueError:
            pass
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("V", [7, 64])
def test_sharded_solve_and_gather_world2(V):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, V, 5, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
