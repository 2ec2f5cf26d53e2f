"""Developer measurement: gather of the cfg3 outputs (491 MB per rank) alone -- PeerGather vs NCCL all-gather.
Launch with torch.distributed.run."""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from pybold_b200.sharding import ALL_OUTPUTS, PeerGather, gather_outputs

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
V, T, K, n = 100000, 300, 20, 100
f32 = torch.float32
spec = {"x": ((T,), f32), "z": ((T,), f32), "diff_z": ((T,), f32), "h": ((K,), f32), "theta": ((), f32),
        "J": ((n + 2,), f32), "r": ((n + 2,), f32), "g": ((n + 2,), f32)}
local = {k: torch.randn((V,) + tuple(t), dtype=d, device=dev) for k, (t, d) in spec.items()}
nbytes = sum(v.numel() * 4 for v in local.values())
pg = PeerGather(spec, V * world, dev, mapping=os.environ.get("PB_PEER_MAPPING", "symm"))
into = {}


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
        dist.barrier(); torch.cuda.synchronize()
    return best


t_peer = timed(lambda: pg.gather(local))
t_nccl = timed(lambda: gather_outputs(local, V * world, ALL_OUTPUTS, into=into))
ok = all(torch.equal(pg.full[k], into[k]) for k in spec)
p2p = [torch.cuda.can_device_access_peer(lr, r) for r in range(world) if r != lr]
if rank == 0:
    print("world %d, %.0f MB per rank: peer copies %.2f ms (%.0f GB/s out of each GPU), NCCL all-gather %.2f ms; equal: %s; "
          "can_device_access_peer: %s" % (world, nbytes / 1e6, t_peer, nbytes * (world - 1) / t_peer / 1e6, t_nccl, ok, p2p))
dist.destroy_process_group()
