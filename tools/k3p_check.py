"""GPU parity check of k3_tmap_kernel (variant 8) against k3_tma_kernel (variant 7: bit for bit, same DMMA order and the same
epilogue arithmetic) and against the oracle -- run under a short timeout before any timing"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wignerd.jl_b200 as w
from oracle import cref
e = w.Engine(0)
rng = np.random.default_rng(0)
bad = 0
for j, nb in [(2.5, 9), (3, 40), (10, 33), (16.5, 17), (20.5, 64), (33, 20), (49.5, 33), (49.5, 8 * 148 * 2 + 5), (50, 19), (63, 24), (63.5, 17), (7, 1200)]:
    a, g, b = rng.uniform(0, 6.3, nb), rng.uniform(0, 6.3, nb), rng.uniform(0, 3.14, nb)
    e.set_variant(8); x = e.wignerD_batch(j, a, b, g).copy()
    e.set_variant(7); y = e.wignerD_batch(j, a, b, g).copy()
    e.set_variant(0)
    pick = np.unique(np.concatenate([[0, nb - 1, nb // 2], rng.integers(0, nb, 10)]))
    d78 = float(np.abs(x - y).max())
    dor = float(np.abs(x[pick] - cref.wignerD_batch(int(round(2 * j)), a[pick], b[pick], g[pick])).max())
    print((j, nb), "8 vs 7:", d78, " vs oracle:", dor, flush=True)
    if not (d78 == 0.0 and dor < 1e-12): bad += 1
print("FAILED" if bad else "ALL OK", bad)
sys.exit(1 if bad else 0)
