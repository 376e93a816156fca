"""end-to-end (host buffers) rate of wd_d_batch at cfg-2's j for several host-path slab sizes (wd_set_slab_bytes; two slabs are
double-buffered on the device) -- VERDICT r1 item 6"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wignerd.jl_b200 as w
e = w.Engine(0)
nb = 32768
b = np.linspace(0, np.pi, nb)
pin = torch.empty((nb, 201, 201), dtype=torch.float64, pin_memory=True)
for mb in (32, 128, 512, 2048):
    e.set_slab_bytes(mb << 20)
    e.wignerd_batch(100, b, out=pin.numpy())
    t0 = time.perf_counter()
    for _ in range(3): e.wignerd_batch(100, b, out=pin.numpy())
    dt = (time.perf_counter() - t0) / 3
    print("slab %4d MB: %.2f GB/s  %.3e elements/s" % (mb, pin.numpy().nbytes / dt / 1e9, pin.numel() / dt), flush=True)
