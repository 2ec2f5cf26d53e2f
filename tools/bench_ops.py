"""Developer measurement: HBM-bound operator kernels (GB/s against the measured copy bandwidth)."""
import json, sys
import torch
sys.path.insert(0, ".")
import pybold_b200 as pb
from pybold_b200 import _lib
from pybold_b200.io import voxels_from_timeseries
from pybold_b200.convolution import simple_convolve, simple_retro_convolve

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

peak = 6554.2
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
V, T, K = 230000, 1200, 28
x = torch.randn((V, T), device="cuda")
k = torch.rand(K, device="cuda")
D = pb.DiscretInteg()
H = pb.ConvAndLinear(D, k, dim_in=T)
xt = x.t().contiguous()     # [T, V]
byt = 2 * V * T * 4
for name, fn in [("integ op (cumsum)", lambda: D.op(x)), ("integ adj", lambda: D.adj(x)),
                 ("conv op (K=28)", lambda: simple_convolve(k, x)), ("conv adj (K=28)", lambda: simple_retro_convolve(k, x)),
                 ("hrfinteg op", lambda: H.op(x)), ("hrfinteg adj", lambda: H.adj(x)),
                 ("layout adapter [T,V]->[V,T]", lambda: voxels_from_timeseries(xt)),
                 ("torch copy (reference)", lambda: x.clone())]:
    ms = timed(fn)
    print("%-32s %8.3f ms  %7.1f GB/s  %.2f of measured HBM copy (%.0f GB/s)" % (name, ms, byt / ms / 1e6, byt / ms / 1e6 / peak, peak))

# tap-count sweep of the convolution: at full HBM rate a K-tap row kernel must also sustain K FFMA per 8 bytes
# (K = 28: 46 Tflop/s, 0.62 of the nominal FP32 pipe) -- fewer taps show what the memory path alone delivers
for Kx in (4, 8, 12, 20, 28):
    kx = torch.rand(Kx, device="cuda")
    ms = timed(lambda: simple_convolve(kx, x))
    print("%-32s %8.3f ms  %7.1f GB/s  %.2f of measured HBM copy  (%.1f Tflop/s of FFMA)"
          % ("conv op (K=%d)" % Kx, ms, byt / ms / 1e6, byt / ms / 1e6 / peak, 2.0 * V * T * Kx / ms / 1e9))

# the cfg3 shape (T = 300, K = 20), same bytes
del x, xt
V, T, K = 920000, 300, 20
x = torch.randn((V, T), device="cuda")
k = torch.rand(K, device="cuda")
H = pb.ConvAndLinear(D, k, dim_in=T)
xt = x.t().contiguous()
byt = 2 * V * T * 4
print("-- %d voxels x %d scans, K = %d --" % (V, T, K))
for name, fn in [("integ op (cumsum)", lambda: D.op(x)), ("integ adj", lambda: D.adj(x)),
                 ("conv op (K=20)", lambda: simple_convolve(k, x)), ("conv adj (K=20)", lambda: simple_retro_convolve(k, x)),
                 ("hrfinteg op", lambda: H.op(x)), ("hrfinteg adj", lambda: H.adj(x)),
                 ("layout adapter [T,V]->[V,T]", lambda: voxels_from_timeseries(xt)),
                 ("torch transposed copy", lambda: xt.t().contiguous()),
                 ("torch copy (reference)", lambda: x.clone())]:
    ms = timed(fn)
    print("%-32s %8.3f ms  %7.1f GB/s  %.2f of measured HBM copy (%.0f GB/s)" % (name, ms, byt / ms / 1e6, byt / ms / 1e6 / peak, peak))
