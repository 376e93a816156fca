"""fast GPU sanity check used between kernel edits (bounded by the caller's `timeout`)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wignerd.jl_b200 as w
from oracle import wignerd_oracle as o
e = w.Engine(0)
bad = 0
def rep(tag, err):
    global bad
    print(tag, err, flush=True)
    if not err < 1e-12: bad += 1
for j, nb in [(100, 40), (10, 1000), (49.5, 33), (2, 5), (0, 3), (0.5, 17), (60, 50), (1, 16), (7.5, 100), (150, 40), (200.5, 19), (300, 17), (103, 9), (63.5, 20), (64.5, 20), (31.5, 77)]:
    b = np.linspace(0, 3.1, nb)
    rep(("unpaired", j, nb), np.abs(e.wignerd_batch(j, b) - o.wignerd_batch(j, b)).max())
# symmetric grids: the beta <-> pi - beta pairing path (even / odd / ragged sizes, several waves)
for j, nb in [(100, 64), (100, 37), (10, 1000), (10, 999), (49.5, 48), (49.5, 33), (3, 16), (3, 17), (0, 20), (0.5, 31), (20, 2500), (100, 16), (100, 18), (101, 23)]:
    b = np.linspace(0, np.pi, nb)
    got = e.wignerd_batch(j, b)
    rep(("paired", j, nb), np.abs(got - o.wignerd_batch(j, b)).max())
# large symmetric grid, many items per CTA: against the unpaired path
import torch
for j, nb in [(20, 148 * 8 * 5 + 13), (100, 148 * 8 * 2 + 5)]:
    b = torch.linspace(0, np.pi, nb, dtype=torch.float64, device="cuda")
    e.set_variant(0); x = e.wignerd_batch(j, b).clone()
    e.set_variant(5); y = e.wignerd_batch(j, b).clone()
    e.set_variant(0)
    torch.cuda.synchronize()
    rep(("paired-vs-unpaired", j, nb), float((x - y).abs().max()))
    pick = np.array([0, 1, 7, 8, nb // 2 - 1, nb // 2, nb // 2 + 1, nb - 9, nb - 8, nb - 1])
    rep(("paired-big-vs-oracle", j, nb), np.abs(x[pick].cpu().numpy() - o.wignerd_batch(j, b[pick].cpu().numpy())).max())
a = np.linspace(0, 6, 33)
rep("D", np.abs(e.wignerD_batch(49.5, a, np.linspace(0, 3.1, 33), a) - o.wignerD_batch(49.5, a, np.linspace(0, 3.1, 33), a)).max())
rep("D", np.abs(e.wignerD_batch(10, a, np.linspace(0, 3.1, 33), a) - o.wignerD_batch(10, a, np.linspace(0, 3.1, 33), a)).max())
print("FAILED" if bad else "ALL OK", bad)
sys.exit(1 if bad else 0)
