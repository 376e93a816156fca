"""fast GPU sanity check used between kernel edits (bounded by the caller's `timeout`)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wignerd.jl_b200 as w
from oracle import wignerd_oracle as o
e = w.Engine(0)
for j, nb in [(100, 40), (10, 1000), (49.5, 33), (2, 5), (0, 3), (0.5, 17), (60, 50), (1, 16), (7.5, 100), (150, 40), (200.5, 19), (300, 17)]:
    b = np.linspace(0, 3.1, nb)
    print(j, nb, np.abs(e.wignerd_batch(j, b) - o.wignerd_batch(j, b)).max(), flush=True)
a = np.linspace(0, 6, 33)
print("D", np.abs(e.wignerD_batch(49.5, a, np.linspace(0, 3.1, 33), a) - o.wignerD_batch(49.5, a, np.linspace(0, 3.1, 33), a)).max())
print("D", np.abs(e.wignerD_batch(10, a, np.linspace(0, 3.1, 33), a) - o.wignerD_batch(10, a, np.linspace(0, 3.1, 33), a)).max())
