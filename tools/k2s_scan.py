"""time wd_d_batch for single large j on a 512-angle slab (streaming kernel) and report F_alg TFLOP/s"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wignerd.jl_b200 as w
eng = w.Engine(0)
nb = 512
beta = torch.linspace(0.0, np.pi, nb, dtype=torch.float64, device="cuda:0")
for j in (30, 60, 100, 111, 112, 150, 200, 256, 384, 512):
    N = 2 * j + 1
    out = torch.empty((nb, N, N), dtype=torch.float64, device="cuda:0")
    for _ in range(3):
        eng.wignerd_batch(j, beta, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.wignerd_batch(j, beta, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    falg = 2.0 * (j + 1) ** 3 * nb
    print(f"j={j:4d}  {ms:8.3f} ms  F_alg {falg / ms / 1e9:6.2f} TFLOP/s  write {nb * N * N * 8 / ms / 1e6:7.1f} GB/s", flush=True)
    del out
