"""probe: is the slow first streaming-kernel j of tools/k2s_scan.py a first-use effect or a property of that j?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wignerd.jl_b200 as w
eng = w.Engine(0)
nb = 512
beta = torch.linspace(0.0, np.pi, nb, dtype=torch.float64, device="cuda:0")
for j in (150, 112, 113, 112, 120, 128, 150):
    N = 2 * j + 1
    out = torch.empty((nb, N, N), dtype=torch.float64, device="cuda:0")
    for _ in range(3):
        eng.wignerd_batch(j, beta, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.wignerd_batch(j, beta, out=out)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"j={j:4d}  " + "  ".join(f"{t:8.3f}" for t in ts) + " ms", flush=True)
    del out
