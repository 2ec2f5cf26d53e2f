"""PeerGather (pybold_b200/sharding.py): the copy-engine all-gather through CUDA IPC mappings, run as two
processes that share cuda:0 (the driver's GPU test box has one GPU).  The object exchange and the fence go
over gloo here (host barrier after a stream synchronise): no kernel of one process ever waits for the other."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, V, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pybold_b200.sharding import PeerGather, voxel_range
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        spec = {"z": ((5,), torch.float32), "theta": ((), torch.float32), "J": ((3,), torch.float64)}
        pg = PeerGather(spec, V, dev, mapping="ipc")
        lo, hi = voxel_range(V, rank, world)
        ok = True
        for rep in range(3):                                   # reused across steps
            rows = torch.arange(lo, hi, device=dev, dtype=torch.float32)
            local = {"z": (rows[:, None] * 10 + torch.arange(5, device=dev) + rep).contiguous(),
                     "theta": rows + 0.5 + rep,
                     "J": (rows[:, None].double() * 2 + torch.arange(3, device=dev) - rep).contiguous()}
            full = pg.gather(local)
            allr = torch.arange(V, device=dev, dtype=torch.float32)
            ok = ok and torch.equal(full["z"], allr[:, None] * 10 + torch.arange(5, device=dev) + rep)
            ok = ok and torch.equal(full["theta"], allr + 0.5 + rep)
            ok = ok and torch.equal(full["J"], allr[:, None].double() * 2 + torch.arange(3, device=dev) - rep)
            dist.barrier()                                     # nobody overwrites before everyone has checked
        try:
            pg.gather({"z": local["z"][:-1], "theta": local["theta"], "J": local["J"]})
            ok = False
        except ValueError:
            pass
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("V", [7, 64])
def test_peer_gather_two_processes_one_gpu(V):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), V, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _worker_sharded(rank, world, port, V, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from pybold_b200.bold_signal import bd_batch
        from pybold_b200.sharding import bd_sharded
        from pybold_b200.synth import gen_voxels
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        T = 96
        y = torch.as_tensor(gen_voxels(V, T, 1.0, 20.0, seed0=4100).astype(np.float32), device=dev)

        def solve(yl, scale=1.0):
            return bd_batch(yl * scale, 1.0, 1.2, 2.0, None, 20.0, [(0.6, 1.9)], 5, False, 4, 1e-12)
        ok = True
        for scale in (1.0, 0.5):                            # second call reuses the cached gatherer
            ref = solve(y, scale)
            full = bd_sharded(y, V, lambda yl: solve(yl, scale), gather="peer", peer_mapping="ipc")
            torch.cuda.synchronize()
            for k in ("x", "z", "diff_z", "h", "theta", "J", "r", "g"):
                ok = ok and full[k].shape[0] == V and torch.equal(full[k], ref[k])
            dist.barrier()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_bd_sharded_peer_gather_two_processes_one_gpu():
    """bd_sharded(gather="peer"): every rank solves its voxel range, the cached PeerGather assembles all
    outputs everywhere; bit-identical to the single-process solve (a voxel's arithmetic does not depend on
    its neighbours in the batch)."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_sharded, args=(world, _free_port(), 37, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
