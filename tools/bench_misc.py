"""Developer measurement: throughput of the fallback paths (generic kernel, FP64 variants)."""
import sys, json
import torch
sys.path.insert(0, ".")
from tools.bench_configs import run_bd
run_bd("T=700 K=28 FP32 (CTA R12 NW2)", 8000, 700, 0.72)
run_bd("T=700 K=28 FP64 (generic kernel)", 2000, 700, 0.72, dtype=torch.float64)
run_bd("T=1200 K=28 FP64 (CTA R20 NW2, double)", 4000, 1200, 0.72, dtype=torch.float64)
run_bd("T=200 K=40 (TR=0.5) FP32 (generic kernel)", 4000, 200, 0.5)
run_bd("T=100 K=20 FP32 (old one-warp kernel R10)", 40000, 100, 1.0)
run_bd("T=400 K=20 FP32 (CTA R16 NW1)", 20000, 400, 1.0)
run_bd("T=900 K=28 FP32 (CTA R16 NW2)", 8000, 900, 0.72)
