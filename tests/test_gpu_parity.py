"""Parity of the CUDA path (through the C ABI / the reference-named host API) against the CPU oracle.
All tests need a B200: run with ``pytest -m gpu``.  Tolerance: BASELINE.json north_star -- 1e-12 absolute
per element, identical signs and phase conventions; most checks are held to a tighter figure."""
import json
import os
import threading
from fractions import Fraction as F
from math import pi

import numpy as np
import pytest

import reference_suite as rs
import wignerd.jl_b200 as w
from oracle import wignerd_oracle as o

pytestmark = pytest.mark.gpu
TOL = 1e-12          # north_star tolerance (absolute, per element)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class GpuImpl:
    """the product's reference-named API on the session engine"""
    def __init__(self, eng):
        self.wignerd, self.wignerD = eng.wignerd, eng.wignerD
        self.wignerdjmn, self.wignerDjmn = eng.wignerdjmn, eng.wignerDjmn
    Equator, NorthPole, SouthPole = w.Equator, w.NorthPole, w.SouthPole


# ---- the reference's own test-suite, against the GPU -------------------------------------------------
def test_readme_known_answers(engine):
    g = json.load(open(os.path.join(GOLD, "readme_vectors.json")))
    assert np.allclose(engine.wignerd(0.5, 0), np.array(g["wignerd_0.5_0"]), atol=1e-15)
    assert np.allclose(engine.wignerd(1, pi / 3), np.array(g["wignerd_1_pi_3"]), atol=5e-7)
    assert abs(engine.wignerdjmn(1, 1, 1, pi / 3) - g["wignerdjmn_1_1_1_pi_3"]) < 1e-15
    D = engine.wignerD(1, 0, pi / 3, pi / 2)
    ref = np.array(g["wignerD_1_0_pi_3_pi_2"])
    assert np.allclose(D.real, ref[..., 0], atol=5e-7) and np.allclose(D.imag, ref[..., 1], atol=5e-7)
    z = engine.wignerDjmn(1, 1, 1, 0, pi / 3, pi / 2)
    assert abs(z - complex(*g["wignerDjmn_1_1_1_0_pi_3_pi_2"])) < 1e-15


def test_ref_special_jmn(engine):
    rs.check_special_jmn(GpuImpl(engine))


def test_ref_Djmn_vs_djmn(engine):
    rs.check_Djmn_vs_djmn(GpuImpl(engine), nbeta=5)


def test_ref_closed_forms(engine):
    rs.check_closed_forms(GpuImpl(engine), ntheta=100)


def test_ref_symmetries(engine):
    rs.check_symmetries(GpuImpl(engine), nbeta=4)


def test_ref_wignerD(engine):
    rs.check_wignerD(GpuImpl(engine), ngrid=4)


def test_ref_cartesian(engine):
    rs.check_cartesian(GpuImpl(engine), ngrid=5)


# ---- spectral setup (K1) --------------------------------------------------------------------------------
@pytest.mark.parametrize("twoj", [0, 1, 2, 3, 20, 99, 200, 401, 1024, 1100, 2001])
def test_eigvecs_residual_orthogonality(engine, twoj):
    u = engine.eigvecs(F(twoj, 2))
    N = twoj + 1
    k = np.arange(1, N)
    b = 0.5 * np.sqrt(k * (N - k))
    T = np.diag(b, 1) + np.diag(b, -1)
    lam = np.arange(N) - twoj / 2
    assert np.abs(T @ u - u * lam).max() < 2e-13 * max(1.0, twoj / 200)      # ~ eps * ||T||
    assert np.abs(u.T @ u - np.eye(N)).max() < 1e-13
    # same subspaces as LAPACK zheevr on the reference's own J_y: |<m|mu>| agree
    if twoj <= 200:
        vs = o.jy_eigen(twoj)                          # vs[mu, m]
        assert np.abs(np.abs(vs.T) - np.abs(u)).max() < 1e-13


# ---- matrices (K2/K3) vs oracle ----------------------------------------------------------------------------
def _cmp_d(engine, twoj, betas, tol=TOL):
    got = engine.wignerd_batch(F(twoj, 2), betas)
    ref = o.wignerd_batch(F(twoj, 2), betas)
    err = np.abs(got - ref).max()
    assert err < tol, (twoj, err)
    return err


def test_cfg1_j10_1000_angles(engine):
    betas = np.linspace(0, pi, 1000)
    assert _cmp_d(engine, 20, betas, 1e-13) < 1e-13


@pytest.mark.parametrize("twoj,nb", [(0, 5), (1, 33), (2, 16), (3, 17), (4, 1), (7, 40), (31, 50), (32, 31),
                                     (33, 64), (63, 20), (64, 20), (99, 48), (127, 19), (200, 37), (255, 18)])
def test_d_batch_vs_oracle(engine, twoj, nb):
    rng = np.random.default_rng(twoj)
    betas = np.concatenate([rng.uniform(0, pi, nb - 1), [rng.uniform(-31, 31)]]) if nb > 1 else np.array([0.3])
    _cmp_d(engine, twoj, betas)


@pytest.mark.parametrize("twoj", [5, 20, 99, 200])
def test_variants_agree(engine, twoj):
    """the DMMA kernel and the plain FP64-ALU kernel are two independent implementations"""
    betas = np.random.default_rng(1).uniform(0, pi, 35)
    try:
        engine.set_variant(1)
        a = engine.wignerd_batch(F(twoj, 2), betas).copy()
        engine.set_variant(2)
        b = engine.wignerd_batch(F(twoj, 2), betas).copy()
    finally:
        engine.set_variant(0)
    assert np.abs(a - b).max() < 1e-13
    assert np.abs(a - o.wignerd_batch(F(twoj, 2), betas)).max() < TOL


@pytest.mark.parametrize("twoj,nb", [(241, 33), (300, 16), (401, 21), (512, 50)])
def test_streaming_kernel_vs_plain_and_oracle(engine, twoj, nb):
    """tables that do not fit in shared memory take the mu-streaming DMMA kernel (auto variant); the plain
    FP64-ALU kernel (variant 1) is the independent cross-check"""
    betas = np.random.default_rng(twoj).uniform(-1.0, 4.0, nb)
    try:
        engine.set_variant(1)
        a = engine.wignerd_batch(F(twoj, 2), betas).copy()
    finally:
        engine.set_variant(0)
    b = engine.wignerd_batch(F(twoj, 2), betas)
    assert np.abs(a - b).max() < 1e-13
    assert np.abs(b - o.wignerd_batch(F(twoj, 2), betas)).max() < TOL


@pytest.mark.parametrize("twoj", [9, 12, 41])
def test_multi_wave_ragged_batches(engine, twoj):
    """more 16-angle chunks than SMs, last chunk ragged: several items per CTA (trig-table double buffer, ring phases
    across items) and a split last wave in the real-d kernel; complex D alongside.  Every matrix against the plain
    FP64-ALU kernel, a sample against the oracle."""
    nb = 148 * 16 * 2 + 16 * 3 + 5
    rng = np.random.default_rng(twoj)
    b = rng.uniform(-0.5, pi + 0.5, nb)
    a, g = rng.uniform(0, 2 * pi, nb), rng.uniform(0, 2 * pi, nb)
    try:
        engine.set_variant(1)
        d1 = engine.wignerd_batch(F(twoj, 2), b).copy()
        D1 = engine.wignerD_batch(F(twoj, 2), a, b, g).copy()
    finally:
        engine.set_variant(0)
    d0 = engine.wignerd_batch(F(twoj, 2), b)
    D0 = engine.wignerD_batch(F(twoj, 2), a, b, g)
    assert np.abs(d0 - d1).max() < 1e-13
    assert np.abs(D0 - D1).max() < 1e-12
    pick = np.concatenate([rng.integers(0, nb, 40), [0, nb - 1, nb - 5, nb - 6, 148 * 16, 148 * 16 - 1]])
    assert np.abs(d0[pick] - o.wignerd_batch(F(twoj, 2), b[pick])).max() < TOL
    assert np.abs(D0[pick] - o.wignerD_batch(F(twoj, 2), a[pick], b[pick], g[pick])).max() < TOL


@pytest.mark.parametrize("twoj", [300, 600, 1024])
def test_large_j_vs_oracle(engine, twoj):
    betas = np.array([0.0, 0.4, pi / 2, 2.9, pi])
    got = engine.wignerd_batch(F(twoj, 2), betas)
    ref = o.wignerd_batch(F(twoj, 2), betas)
    assert np.abs(got - ref).max() < TOL


@pytest.mark.parametrize("twoj,nb", [(0, 3), (1, 20), (2, 33), (5, 16), (20, 30), (99, 48), (100, 21), (200, 18)])
def test_D_batch_vs_oracle(engine, twoj, nb):
    rng = np.random.default_rng(100 + twoj)
    a, g = rng.uniform(0, 2 * pi, nb), rng.uniform(0, 2 * pi, nb)
    b = rng.uniform(0, pi, nb)
    got = engine.wignerD_batch(F(twoj, 2), a, b, g)
    ref = o.wignerD_batch(F(twoj, 2), a, b, g)
    assert np.abs(got - ref).max() < TOL
    try:                                  # plain kernel uses the reference's exact cis(-(m a + n g))
        engine.set_variant(1)
        got1 = engine.wignerD_batch(F(twoj, 2), a, b, g)
    finally:
        engine.set_variant(0)
    assert np.abs(got1 - ref).max() < TOL


def test_mp_exact_samples(engine):
    """exact multiprecision sums (fixtures): triangulates GPU vs the absent Julia run"""
    g = json.load(open(os.path.join(GOLD, "mp_exact_samples.json")))
    s = g["samples"]
    tj = np.array([x["twoj"] for x in s], np.int32)
    tm = np.array([x["twom"] for x in s], np.int32)
    tn = np.array([x["twon"] for x in s], np.int32)
    beta = np.array([float.fromhex(x["beta"]) for x in s])
    exact = np.array([float.fromhex(x["d"]) for x in s])
    got = engine.wignerdjmn_batch(tj, tm, tn, beta)
    assert np.abs(got - exact).max() < 2e-13
    # and through the matrix path
    for x in s[::7]:
        d = engine.wignerd(F(x["twoj"], 2), float.fromhex(x["beta"]))
        i, k = (x["twom"] + x["twoj"]) // 2, (x["twon"] + x["twoj"]) // 2
        assert abs(d[i, k] - float.fromhex(x["d"])) < 2e-13


def test_scalable_invariants_j100(engine):
    """size-independent properties at the headline j: trace, orthogonality, det, the symmetry identities"""
    betas = np.random.default_rng(7).uniform(0, pi, 24)
    d = engine.wignerd_batch(100, betas)
    m = np.arange(-100, 101)
    for b, db in zip(betas, d):
        assert abs(np.trace(db) - np.cos(m * b).sum()) < 1e-11
        assert np.abs(db.T @ db - np.eye(201)).max() < 1e-12
    dm = engine.wignerd_batch(100, -betas)
    assert np.abs(dm.transpose(0, 2, 1) - d).max() < 1e-12                       # d(-b)' = d(b)
    dp = engine.wignerd_batch(100, pi - betas)
    sg = (-1.0) ** (100 + m)
    assert np.abs(dp - sg[None, :, None] * d[:, :, ::-1]).max() < 1e-11        # d(pi-b)[m,n] = (-1)^(j+m) d(b)[m,-n]


def test_special_angles_matrix(engine):
    for j in [0, 0.5, 1, 1.5, 2, 3.5, 10, 49.5]:
        for sp in (w.Equator(), w.NorthPole(), w.SouthPole()):
            got = engine.wignerd(j, sp)
            assert np.abs(got - o.wignerd(j, {w.EQUATOR: o.Equator, w.NORTH: o.NorthPole, w.SOUTH: o.SouthPole}[sp.kind]())).max() < 1e-13
            D = engine.wignerD(j, 0.7, sp, -1.3)
            Dref = o.wignerD(j, 0.7, {w.EQUATOR: o.Equator, w.NORTH: o.NorthPole, w.SOUTH: o.SouthPole}[sp.kind](), -1.3)
            assert np.abs(D - Dref).max() < 1e-13
    # Equator exact zeros (:230-237)
    d = engine.wignerd(4, w.Equator())
    assert d[4 + 1, 4] == 0.0 and d[4, 4 + 3] == 0.0
    # in-place semantics: the poles only touch the (anti-)diagonal (:247-264)
    buf = np.full((3, 3), 7.0, order="F")
    engine.wignerd_(buf, 1, w.NorthPole())
    assert np.array_equal(buf, np.where(np.eye(3) == 1, 1.0, 7.0))


# ---- single elements (K4) -----------------------------------------------------------------------------------
def test_jmn_batch_vs_oracle(engine):
    rng = np.random.default_rng(3)
    n = 20000
    tj = rng.integers(0, 81, n).astype(np.int32)
    tm = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
    tn = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
    b = rng.uniform(0, pi, n)
    got = engine.wignerdjmn_batch(tj, tm, tn, b)
    ref = o.wignerdjmn_batch(tj, tm, tn, b)
    assert np.abs(got - ref).max() < TOL
    a, g = rng.uniform(0, 2 * pi, n), rng.uniform(0, 2 * pi, n)
    gotD = engine.wignerDjmn_batch(tj, tm, tn, a, b, g)
    x = -(a * tm / 2.0 + g * tn / 2.0)
    assert np.abs(gotD - ref * (np.cos(x) + 1j * np.sin(x))).max() < TOL


def test_jmn_large_j(engine):
    rng = np.random.default_rng(5)
    for twoj in (400, 399, 1024):
        n = 64
        tj = np.full(n, twoj, np.int32)
        tm = (-tj + 2 * rng.integers(0, twoj + 1, n)).astype(np.int32)
        tn = (-tj + 2 * rng.integers(0, twoj + 1, n)).astype(np.int32)
        b = rng.uniform(0, pi, n)
        assert np.abs(engine.wignerdjmn_batch(tj, tm, tn, b) - o.wignerdjmn_batch(tj, tm, tn, b)).max() < TOL


def test_jmn_matches_matrix_path(engine):
    betas = np.array([0.37, 2.2])
    for twoj in (7, 20, 99):
        d = engine.wignerd_batch(F(twoj, 2), betas)
        N = twoj + 1
        mi, ni = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
        for k, b in enumerate(betas):
            e = engine.wignerdjmn_batch(np.full(N * N, twoj), 2 * mi.ravel() - twoj, 2 * ni.ravel() - twoj, np.full(N * N, b))
            assert np.abs(e.reshape(N, N) - d[k]).max() < 1e-13


# ---- error behaviour, in-place API, Jy, threads, device memory ------------------------------------------
def test_errors(engine):
    with pytest.raises(w.ArgumentError):
        engine.wignerdjmn(1, 2, 0, 0.3)                     # :191
    with pytest.raises(w.ArgumentError):
        engine.wignerdjmn(1, 0.5, 0, 0.3)
    with pytest.raises(w.InexactError):
        engine.wignerd(0.3, 0.1)                            # Int(2j+1), :130
    with pytest.raises(AssertionError):
        engine.wignerd(-1, 0.1)
    with pytest.raises(AssertionError):
        engine.wignerd_(np.zeros((2, 2)), 1, 0.1)           # :267
    with pytest.raises(w.ArgumentError):
        engine.wignerdjmn_batch(np.array([4, 4]), np.array([0, 6]), np.array([0, 0]), np.array([0.1, 0.2]))
    # poles do not check indices, like the reference (:218-228)
    assert engine.wignerdjmn(1, 5, 5, w.NorthPole()) == 1.0


def test_Jy_argument(engine):
    eng = w.Engine(0)
    try:
        with pytest.raises(AssertionError):
            eng.wignerd(2, 0.3, np.zeros((5, 5)))                       # eltype must be ComplexF64 (:148)
        with pytest.raises(AssertionError):
            eng.wignerd(2, 0.3, np.zeros((4, 4), np.complex128))        # size (:149)
        Jy = np.zeros((5, 5), np.complex128)
        a = eng.wignerd(2, 0.3, Jy)
        assert np.array_equal(a, eng.wignerd(2, 0.3))                   # with/without Jy exactly equal (:101)
        eng.wignerd(2, 0.3, np.zeros((4, 4)))                           # cache hit: asserts not evaluated
        for m in (-2, 0, 1):
            assert eng.wignerDjmn(2, m, 1, 0, 0.9, 0, Jy) == eng.wignerDjmn(2, m, 1, 0, 0.9, 0)
    finally:
        eng.close()


def test_inplace_and_offsetmatrix(engine):
    d = np.zeros((4, 4), order="F")
    r = engine.wignerd_(d, 1.5, 0.8)
    assert r is d
    dw = w._offsetmatrix(1.5, d)
    assert abs(dw[-1.5, -1.5] - np.cos(0.4) ** 3) < 1e-14
    c = np.zeros((4, 4))                                               # C-ordered buffer
    engine.wignerd_(c, 1.5, 0.8)
    assert np.array_equal(c, d)
    D = np.zeros((3, 3), np.complex128, order="F")
    assert engine.wignerD_(D, 1, 0.1, 0.2, 0.3) is D
    assert np.abs(D - o.wignerD(1, 0.1, 0.2, 0.3)).max() < 1e-14


def test_threads_each_with_own_engine():
    """the reference is safe under Threads.@threads via its per-Task cache (test/runtests.jl:96,135)"""
    errs = []

    def work(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(20):
                tj = int(rng.integers(0, 12))
                b = float(rng.uniform(0, pi))
                if np.abs(w.wignerd(F(tj, 2), b) - o.wignerd_batch(F(tj, 2), [b])[0]).max() > 1e-13:
                    errs.append((tj, b))
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=work, args=(s,)) for s in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_device_memory_api(engine):
    import torch
    betas = torch.linspace(0.0, pi, 50, dtype=torch.float64, device="cuda:0")
    d = engine.wignerd_batch(10, betas)
    torch.cuda.synchronize()
    assert d.is_cuda and tuple(d.shape) == (50, 21, 21)
    ref = o.wignerd_batch(10, betas.cpu().numpy())
    assert np.abs(d.cpu().numpy() - ref).max() < 1e-13
    a = torch.rand(50, dtype=torch.float64, device="cuda:0")
    D = engine.wignerD_batch(10, a, betas, a)
    torch.cuda.synchronize()
    assert np.abs(D.cpu().numpy() - o.wignerD_batch(10, a.cpu().numpy(), betas.cpu().numpy(), a.cpu().numpy())).max() < TOL
    assert engine.launch_count() > 0


def test_multi_j(engine):
    betas = np.linspace(0, pi, 40)
    js = [F(t, 2) for t in range(0, 24)]
    _, offsets, views = engine.wignerd_multi(js, betas)
    for j, v in zip(js, views):
        assert np.abs(v - o.wignerd_batch(j, betas)).max() < 1e-13


def test_full_size_cfg2_properties(engine):
    """cfg-2 at FULL size (j=100, 10^5 angles, 32 GB resident): checked through size-independent
    properties -- trace identity on every matrix, plus oracle parity on a random sample of matrices."""
    import torch
    nb, N = 100000, 201
    betas = torch.linspace(0.0, pi, nb, dtype=torch.float64, device="cuda:0")
    d = engine.wignerd_batch(100, betas)
    torch.cuda.synchronize()
    tr = torch.diagonal(d, dim1=1, dim2=2).sum(-1)
    m = torch.arange(-100, 101, dtype=torch.float64, device="cuda:0")
    ref_tr = torch.cos(betas[:, None] * m[None, :]).sum(-1)
    assert float((tr - ref_tr).abs().max()) < 1e-11
    # d(pi - b)[m, n] = (-1)^(j+m) d(b)[m, -n]: the linspace grid is symmetric, so matrix b pairs with nb-1-b
    sg = torch.where(torch.arange(N, device="cuda:0") % 2 == 0, 1.0, -1.0).to(torch.float64)   # (-1)^(j+m), j even
    idx = torch.randint(0, nb, (256,), device="cuda:0")
    lhs = d[nb - 1 - idx]
    rhs = sg[None, :, None] * torch.flip(d[idx], dims=[2])
    assert float((lhs - rhs).abs().max()) < 1e-10
    pick = np.random.default_rng(11).integers(0, nb, 6)
    ref = o.wignerd_batch(100, betas[pick].cpu().numpy())
    assert np.abs(d[pick].cpu().numpy() - ref).max() < TOL
    del d
    torch.cuda.empty_cache()


# ---- edge cases -------------------------------------------------------------------------------------------
def test_empty_and_single_batches(engine):
    assert engine.wignerd_batch(3, np.zeros(0)).shape == (0, 7, 7)
    assert engine.wignerD_batch(1.5, np.zeros(0), np.zeros(0), np.zeros(0)).shape == (0, 4, 4)
    assert engine.wignerdjmn_batch(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0)).shape == (0,)
    for twoj in (0, 1, 2, 201, 300):
        d = engine.wignerd_batch(F(twoj, 2), np.array([1.234]))
        assert np.abs(d[0] - o.wignerd_batch(F(twoj, 2), [1.234])[0]).max() < TOL


def test_large_and_special_float_angles(engine):
    # |mu*beta| ~ 1e7: both sides form fl(mu*beta) and call a full-range sincos, so they still agree to the north_star tolerance
    for beta in (1.0e6, -3.3e5, 12345.678):
        for j in (2, 10.5):
            assert np.abs(engine.wignerd_batch(j, [beta])[0] - o.wignerd_batch(j, [beta])[0]).max() < TOL
    d = engine.wignerd_batch(2, np.array([np.nan, np.inf, 0.5]))
    assert np.isnan(d[0]).all() and np.isnan(d[1]).all() and np.isfinite(d[2]).all()
    assert np.abs(engine.wignerd(2, 0.0) - np.eye(5)).max() < 1e-15
    assert np.array_equal(engine.wignerd(2, -0.0), engine.wignerd(2, 0.0))


def test_shared_engine_from_threads(engine):
    """one handle used from several threads: calls are serialised by the handle's mutex"""
    errs = []

    def work(seed):
        rng = np.random.default_rng(seed)
        for _ in range(10):
            tj = int(rng.integers(0, 30))
            b = rng.uniform(0, pi, 5)
            if np.abs(engine.wignerd_batch(F(tj, 2), b) - o.wignerd_batch(F(tj, 2), b)).max() > 1e-13:
                errs.append(tj)

    ts = [threading.Thread(target=work, args=(s,)) for s in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs


def test_device_memory_jmn_and_multi(engine):
    import torch
    dev = "cuda:0"
    rng = np.random.default_rng(9)
    n = 5000
    tj = rng.integers(0, 51, n).astype(np.int32)
    tm = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
    tn = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
    b = rng.uniform(0, pi, n)
    out = engine.wignerdjmn_batch(torch.from_numpy(tj).to(dev), torch.from_numpy(tm).to(dev), torch.from_numpy(tn).to(dev),
                                  torch.from_numpy(b).to(dev))
    assert np.abs(out.cpu().numpy() - o.wignerdjmn_batch(tj, tm, tn, b)).max() < TOL
    betas = torch.linspace(0.1, 3.0, 33, dtype=torch.float64, device=dev)
    js = [0, 0.5, 1, 7, 10.5, 40]
    _, _, views = engine.wignerd_multi(js, betas)
    torch.cuda.synchronize()
    for j, v in zip(js, views):
        assert np.abs(v.cpu().numpy() - o.wignerd_batch(j, betas.cpu().numpy())).max() < 1e-13


def test_prepare_and_cache_clear():
    eng = w.Engine(0)
    try:
        eng.prepare(range(0, 41))
        n0 = eng.launch_count()
        a = eng.wignerd(10, 0.7)
        eng.cache_clear()
        b = eng.wignerd(10, 0.7)
        assert np.array_equal(a, b) and eng.launch_count() > n0
        u = eng.eigvecs(10)
        assert np.abs(u.T @ u - np.eye(21)).max() < 1e-14
    finally:
        eng.close()


def test_in_place_complex_poles(engine):
    """wignerD! with NorthPole/SouthPole multiplies the caller's matrix by the phases after the pole fill (:335-342)"""
    for sp, osp in ((w.NorthPole(), o.NorthPole()), (w.SouthPole(), o.SouthPole())):
        D = engine.wignerD(2.5, 0.3, sp, 1.1)
        assert np.abs(D - o.wignerD(2.5, 0.3, osp, 1.1)).max() < 1e-14


@pytest.mark.parametrize("twoj", [1100, 1401, 2000])
def test_very_large_j_invariants(engine, twoj):
    """beyond 2j ~ 1024 the eigenvector tails underflow 2^-j: the twist index must stay in the classically allowed
    region (regression test).  The oracle needs minutes here, so the size-independent properties are checked."""
    j = F(twoj, 2)
    N = twoj + 1
    betas = np.array([0.7, 2.9])
    d = engine.wignerd_batch(j, betas)
    m = np.arange(N) - twoj / 2
    for b, db in zip(betas, d):
        assert np.abs(db.T @ db - np.eye(N)).max() < 1e-12
        assert abs(np.trace(db) - np.cos(m * b).sum()) < 1e-11
    dm = engine.wignerd_batch(j, -betas)
    assert np.abs(dm.transpose(0, 2, 1) - d).max() < 1e-12


# ---- complex D through the tensor-store kernel (k3_tma_kernel, variant 7 = forced) ----------------------------------
@pytest.mark.parametrize("twoj,nb", [(5, 9), (6, 40), (20, 8), (33, 7), (63, 100), (64, 17), (99, 33), (127, 24), (126, 9), (99, 1)])
def test_tensor_store_kernel_vs_oracle_and_other_kernels(engine, twoj, nb):
    """k3_tma_kernel (phase in registers, TMA tensor stores) against the oracle, the plain FP64-ALU kernel (the reference's exact
    cis(-(m alpha + n gamma))) and k2_cplx_kernel (variant 2): integer and half-integer j over its whole range 5 <= 2j <= 127,
    batches that end inside an 8-angle item"""
    rng = np.random.default_rng(700 + twoj)
    a, g = rng.uniform(-2 * pi, 2 * pi, nb), rng.uniform(0, 2 * pi, nb)
    b = rng.uniform(0, pi, nb)
    got = {}
    try:
        for v in (7, 2, 1):
            engine.set_variant(v)
            got[v] = engine.wignerD_batch(F(twoj, 2), a, b, g).copy()
    finally:
        engine.set_variant(0)
    ref = o.wignerD_batch(F(twoj, 2), a, b, g)
    assert np.abs(got[7] - ref).max() < TOL
    assert np.abs(got[7] - got[1]).max() < TOL
    assert np.abs(got[7] - got[2]).max() < 4e-13       # k3_tma_kernel makes its phase tables by anchored rotations (<= 6e-14 per factor from one sincos per entry)
    assert np.abs(engine.wignerD_batch(F(twoj, 2), a, b, g) - got[7]).max() == 0.0      # auto mode takes the same kernel


@pytest.mark.parametrize("twoj", [41, 99, 100])
def test_tensor_store_kernel_many_items(engine, twoj):
    """more 8-angle items than SMs (table double buffer across items, rotating column shares, split last wave), ragged end:
    every matrix against k2_cplx_kernel, unitarity of every matrix, a sample against the oracle"""
    import torch
    nb = 148 * 8 * 3 + 8 * 5 + 3
    gen = torch.Generator(device="cuda").manual_seed(twoj)
    a = torch.rand(nb, dtype=torch.float64, device="cuda", generator=gen) * (2 * pi)
    g = torch.rand(nb, dtype=torch.float64, device="cuda", generator=gen) * (2 * pi)
    b = torch.rand(nb, dtype=torch.float64, device="cuda", generator=gen) * pi
    try:
        engine.set_variant(7)
        D7 = engine.wignerD_batch(F(twoj, 2), a, b, g).clone()
        engine.set_variant(2)
        D2 = engine.wignerD_batch(F(twoj, 2), a, b, g).clone()
    finally:
        engine.set_variant(0)
    torch.cuda.synchronize()
    assert float((D7 - D2).abs().max()) < 4e-13
    eye = torch.eye(twoj + 1, dtype=D7.dtype, device="cuda")
    assert float((D7.conj().transpose(1, 2) @ D7 - eye).abs().max()) < 1e-12
    pick = np.array([0, 1, 7, 8, 9, 148 * 8 - 1, 148 * 8, nb - 12, nb - 4, nb - 3, nb - 1])
    an, bn, gn = (x[pick].cpu().numpy() for x in (a, b, g))
    assert np.abs(D7[pick].cpu().numpy() - o.wignerD_batch(F(twoj, 2), an, bn, gn)).max() < TOL


def test_tensor_store_kernel_range_and_alignment(engine):
    """forced, the kernel refuses what it cannot do (2j > 127); in auto mode those
    requests take k2_cplx_kernel / the streaming path and give the same matrices"""
    import ctypes
    import torch
    rng = np.random.default_rng(5)
    nb = 20
    a, b, g = rng.uniform(0, 6, nb), rng.uniform(0, 3, nb), rng.uniform(0, 6, nb)
    try:
        engine.set_variant(7)
        with pytest.raises(NotImplementedError):
            engine.wignerD_batch(64, a, b, g)
        with pytest.raises(NotImplementedError):
            engine.wignerD_batch(100, a, b, g)
        # real d is dispatched as in auto mode
        assert np.abs(engine.wignerd_batch(10, b) - o.wignerd_batch(10, b)).max() < TOL
    finally:
        engine.set_variant(0)
    assert np.abs(engine.wignerD_batch(64, a, b, g) - o.wignerD_batch(64, a, b, g)).max() < TOL
    # raw ABI: a complex device output must be 16-byte aligned (16-byte stores, TMA tensor): refused as an argument error
    twoj, N = 21, 22
    ang = torch.from_numpy(np.stack([a, b, g])).cuda()
    buf = torch.zeros(2 * nb * N * N + 2, dtype=torch.float64, device="cuda")
    L, h = engine._lib, engine._h
    vp = ctypes.c_void_p
    args = lambda off: (h, twoj, vp(ang[0].data_ptr()), vp(ang[1].data_ptr()), vp(ang[2].data_ptr()), nb, vp(buf.data_ptr() + off), w.MEM_DEVICE)  # noqa: E731
    assert L.wd_D_batch(*args(8)) == w.api.WD_ERR_ASSERT
    assert L.wd_D_batch(*args(0)) == 0
    torch.cuda.synchronize()
    off0 = buf[:2 * nb * N * N].cpu().numpy()
    ref = o.wignerD_batch(F(twoj, 2), a, b, g).transpose(0, 2, 1)
    assert np.abs(off0.view(np.complex128).reshape(nb, N, N) - ref).max() < TOL
