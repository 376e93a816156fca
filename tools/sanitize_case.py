"""small ragged inputs through every kernel of the library, for compute-sanitizer (one --tool per gpurun call):
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
    compute-sanitizer --tool synccheck python tools/sanitize_case.py
Host (numpy) buffers only; results are compared with the oracle so that a run that "passes" the tool also computed the right thing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wignerd.jl_b200 as w
from oracle import cref

e = w.Engine(0)
bad = 0
def check(tag, got, ref, tol=1e-12):
    global bad
    err = float(np.abs(got - ref).max())
    print(tag, err, flush=True)
    bad += not err < tol

big = int(os.environ.get("WD_SANITIZE_BIG", "0"))
rng = np.random.default_rng(0)
for twoj, nbs in ((9, (5, 37, 2373 if big else 300)), (41, (37, 200)), (99, (21, 40)), (200, (19, 34)), (241, (18, 35))):
    for nb in nbs:
        sym = np.linspace(0, np.pi, nb)
        gen = rng.uniform(0, np.pi, nb)
        ref_s, ref_g = cref.wignerd_batch(twoj, sym), cref.wignerd_batch(twoj, gen)
        for variant, grid, ref, name in ((0, sym, ref_s, "auto/symmetric"), (0, gen, ref_g, "auto/generic"), (2, gen, ref_g, "k2_real"),
                                         (4, sym, ref_s, "k2_col paired"), (5, gen, ref_g, "k2_col unpaired"), (3, sym, ref_s, "k2_stream paired"),
                                         (3, gen, ref_g, "k2_stream")):
            if twoj > 223 and variant in (2, 4, 5):
                continue
            if twoj > 200 and variant in (4, 5):
                continue
            try:
                e.set_variant(variant)
                check((twoj, nb, name), e.wignerd_batch(twoj / 2, grid), ref)
            finally:
                e.set_variant(0)
        if twoj <= 223:
            a, g = rng.uniform(0, 6, nb), rng.uniform(0, 6, nb)
            check((twoj, nb, "k2_cplx"), e.wignerD_batch(twoj / 2, a, gen, g), cref.wignerD_batch(twoj, a, gen, g))
n = 3000
tj = rng.integers(0, 61, n).astype(np.int32)
tm = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
tn = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
b = rng.uniform(0, np.pi, n)
check("k4_jmn", e.wignerdjmn_batch(tj, tm, tn, b), cref.wignerdjmn_batch(tj, tm, tn, b))
lmax, nb = 6, 21
alm = rng.normal(size=(lmax + 1) ** 2) + 1j * rng.normal(size=(lmax + 1) ** 2)
a, bb, g = rng.uniform(0, 6, nb), rng.uniform(0, 3, nb), rng.uniform(0, 6, nb)
out = e.rotate_sh(lmax, alm, a, bb, g)
ref = np.concatenate([np.einsum("bmn,m->bn", cref.wignerD_batch(2 * l, a, bb, g), alm[l * l:(l + 1) ** 2]) for l in range(lmax + 1)], axis=1)
check("k5_rotate", out, ref, 1e-11)
print("FAILED" if bad else "ALL OK", bad)
e.close()
sys.exit(1 if bad else 0)
