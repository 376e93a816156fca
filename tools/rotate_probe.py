"""wd_rotate_sh timing at lmax = 512 for several batch sizes (device buffers)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wignerd.jl_b200 as w
e = w.Engine(0)
lmax = 512; L2 = (lmax + 1) ** 2
flops1 = 8.0 * sum((2 * l + 1) ** 2 for l in range(lmax + 1))
alm = torch.complex(torch.randn(L2, dtype=torch.float64, device="cuda"), torch.randn(L2, dtype=torch.float64, device="cuda"))
for nb in (128, 512, 2048, 4096):
    a = torch.rand(nb, dtype=torch.float64, device="cuda") * 6; b = torch.linspace(0, np.pi, nb, dtype=torch.float64, device="cuda"); g = torch.rand(nb, dtype=torch.float64, device="cuda") * 6
    out = torch.empty((nb, L2), dtype=torch.complex128, device="cuda")
    e.rotate_sh(lmax, alm, a, b, g, out=out); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): e.rotate_sh(lmax, alm, a, b, g, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(nb, "rotations: %.1f ms, %.1f TFLOP/s" % (ms, flops1 * nb / ms / 1e9))
