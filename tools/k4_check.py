"""parity + timing of wignerdjmn_batch (K4) at cfg-5 and at ragged / large-j tuples"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wignerd.jl_b200 as w
from oracle import cref
e = w.Engine(0)
rng = np.random.default_rng(5)
bad = 0
for tjmax, n in [(400, 20000), (1024, 4000), (33, 5000), (7, 3000)]:
    twoj = rng.integers(0, tjmax + 1, n).astype(np.int32)
    i = (rng.random(n) * (twoj + 1)).astype(np.int32); k = (rng.random(n) * (twoj + 1)).astype(np.int32)
    twom = 2 * i - twoj; twon = 2 * k - twoj
    beta = rng.uniform(-7, 7, n); beta[:5] = [0, np.pi, np.pi / 2, -np.pi, 2 * np.pi]
    got = e.wignerdjmn_batch(twoj, twom, twon, beta)          # the batch API takes 2j, 2m, 2n
    ref = cref.wignerdjmn_batch(twoj, twom, twon, beta)
    err = np.abs(got - ref).max()
    print("jmn 2j<=%d n=%d maxerr %.3e" % (tjmax, n, err), flush=True)
    bad += not err < 1e-12
sys.exit(1 if bad else 0)
