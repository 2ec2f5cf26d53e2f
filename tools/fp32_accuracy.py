"""Developer tool: FP32 build vs FP64 build of bd on the same voxels (relative errors)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200.synth import gen_voxels_chunked
def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())
for (V, T, t_r) in [(2048, 300, 1.0), (2048, 240, 0.75), (512, 600, 1.0), (512, 1200, 0.72)]:
    y = torch.as_tensor(gen_voxels_chunked(V, T, t_r), device="cuda")
    a = pb.bd(y, t_r, lbda=1.7, theta_0=2.0, nb_iter=100)
    b = pb.bd(y.double(), t_r, lbda=1.7, theta_0=2.0, nb_iter=100)
    per_voxel_z = ((a[1].double() - b[1]).abs().amax(dim=1) / b[1].abs().amax(dim=1))
    per_voxel_x = ((a[0].double() - b[0]).abs().amax(dim=1) / b[0].abs().amax(dim=1))
    per_voxel_J = ((a[4]["J"].double() - b[4]["J"]).abs().amax(dim=1) / b[4]["J"].abs().amax(dim=1))
    dth = (a[4]["theta"].double() - b[4]["theta"]).abs()
    print("T=%d V=%d: z rel err median %.2e max %.2e | x max %.2e | J max %.2e | theta abs err median %.2e max %.2e | h max %.2e"
          % (T, V, per_voxel_z.median(), per_voxel_z.max(), per_voxel_x.max(), per_voxel_J.max(), dth.median(), dth.max(), rel(a[3], b[3])))
