import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import wignerd.jl_b200 as w
e = w.Engine(0)
nb = 16384
b = np.linspace(0, np.pi, nb)
pg = np.empty((nb, 201, 201)); pg[...] = 0
pin = torch.empty((nb, 201, 201), dtype=torch.float64, pin_memory=True)
for name, out in (("pageable", pg), ("pinned", pin.numpy())):
    e.wignerd_batch(100, b, out=out)
    t0 = time.perf_counter()
    for _ in range(3): e.wignerd_batch(100, b, out=out)
    dt = (time.perf_counter() - t0) / 3
    print(name, "%.2f GB/s" % (out.nbytes / dt / 1e9), "%.3e el/s" % (out.size / dt))
assert np.array_equal(pg, pin.numpy())
