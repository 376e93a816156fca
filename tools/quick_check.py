"""fast GPU sanity check used between kernel edits (bounded by the caller's `timeout`)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wignerd.jl_b200 as w
from oracle import wignerd_oracle as o
from oracle import cref
e = w.Engine(0)
def ref_d(j, b):      # numpy oracle for small j, its (bit-identical, threaded) C port for large j
    return o.wignerd_batch(j, b) if 2 * j <= 120 else cref.wignerd_batch(int(round(2 * j)), b)
bad = 0
def rep(tag, err):
    global bad
    print(tag, err, flush=True)
    if not err < 1e-12: bad += 1
for j, nb in [(100, 40), (10, 1000), (49.5, 33), (2, 5), (0, 3), (0.5, 17), (60, 50), (1, 16), (7.5, 100), (150, 40), (200.5, 19), (300, 17), (103, 9), (63.5, 20), (64.5, 20), (31.5, 77)]:
    b = np.linspace(0, 3.1, nb)
    rep(("unpaired", j, nb), np.abs(e.wignerd_batch(j, b) - ref_d(j, b)).max())
# symmetric grids: the beta <-> pi - beta pairing path (even / odd / ragged sizes, several waves)
for j, nb in [(100, 64), (100, 37), (10, 1000), (10, 999), (49.5, 48), (49.5, 33), (3, 16), (3, 17), (0, 20), (0.5, 31), (20, 2500), (100, 16), (100, 18), (101, 23)]:
    b = np.linspace(0, np.pi, nb)
    got = e.wignerd_batch(j, b)
    rep(("paired", j, nb), np.abs(got - o.wignerd_batch(j, b)).max())
# streaming kernel (table not resident) on symmetric grids: the store warps also feed the mirror matrices
for j, nb in [(150, 64), (150, 37), (120, 33), (200.5, 40), (200.5, 51), (300, 32), (300, 45), (512, 34)]:
    b = np.linspace(0, np.pi, nb)
    rep(("paired-stream", j, nb), np.abs(e.wignerd_batch(j, b) - ref_d(j, b)).max())
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
# complex D: column-staged kernel (table resident, Q <= 64 / 112), k2_cplx_kernel beyond, plain kernel for tiny j
rng = np.random.default_rng(0)
for j, nb in [(49.5, 33), (10, 33), (49.5, 8 * 148 * 2 + 5), (1.5, 9), (2, 40), (1, 5), (0.5, 3), (31.5, 100), (63.5, 17), (64, 20), (100, 19), (7, 1200), (20.5, 64), (3, 2500)]:
    a, g, b = rng.uniform(0, 6.3, nb), rng.uniform(0, 6.3, nb), rng.uniform(0, 3.14, nb)
    got = e.wignerD_batch(j, a, b, g)
    pick = np.unique(np.concatenate([[0, nb - 1, nb // 2], rng.integers(0, nb, 12)]))
    rep(("D", j, nb), np.abs(got[pick] - cref.wignerD_batch(int(round(2 * j)), a[pick], b[pick], g[pick])).max())
    if nb > 1000:   # every matrix: unitarity
        import torch
        G = torch.from_numpy(np.ascontiguousarray(got)).cuda()
        rep(("D unitarity", j, nb), float((G.conj().transpose(1, 2) @ G - torch.eye(G.shape[1], dtype=G.dtype, device="cuda")).abs().max()))
print("FAILED" if bad else "ALL OK", bad)
sys.exit(1 if bad else 0)
