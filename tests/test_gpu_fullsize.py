"""BASELINE.json's configurations at their FULL sizes on the GPU (run with ``pytest -m gpu``): every matrix / tuple is
checked through size-independent properties (unitarity, trace, inverse, symmetry identities), a sample against the CPU
oracle (numpy restatement / its C port) to the north_star tolerance 1e-12.  Mirrors test/runtests.jl:119-125 (trace),
:330-334 (pi - beta), :338-359 (unitarity, inverse) at the sizes the reference's own suite never reaches (it stops at j <= 5)."""
import os
from fractions import Fraction as F
from math import pi

import numpy as np
import pytest

import wignerd.jl_b200 as w
from oracle import cref
from oracle import wignerd_oracle as o

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _trace_ref(torch, betas, twoj):
    """tr d^j(beta) = sum_m cos(m beta) (test/runtests.jl:123), m = -j..j in steps of 1"""
    m = (torch.arange(twoj + 1, dtype=torch.float64, device=betas.device) - twoj / 2.0)
    return torch.cos(betas[:, None] * m[None, :]).sum(-1)


# ---- cfg-3: wignerD(49.5, alpha, beta, gamma) on 10^5 Euler triples (16 GB) -------------------------------------
def test_full_size_cfg3_unitarity_inverse_and_samples(engine):
    import torch
    dev = "cuda:0"
    nb, twoj, N = 100000, 99, 100
    gen = torch.Generator(device=dev).manual_seed(2)
    alpha = torch.rand(nb, dtype=torch.float64, device=dev, generator=gen) * (2 * pi)
    gamma = torch.rand(nb, dtype=torch.float64, device=dev, generator=gen) * (2 * pi)
    beta = torch.rand(nb, dtype=torch.float64, device=dev, generator=gen) * pi
    D = engine.wignerD_batch(F(twoj, 2), alpha, beta, gamma)               # [b, m, n] view of the library's layout
    torch.cuda.synchronize()
    eye = torch.eye(N, dtype=torch.complex128, device=dev)
    worst_u, worst_i = 0.0, 0.0
    for lo in range(0, nb, 10000):
        Dc = D[lo:lo + 10000]
        worst_u = max(worst_u, float((Dc.conj().transpose(1, 2) @ Dc - eye).abs().max()))     # D'D = I (:351)
        # inv(D) = D' = wignerD(j, -gamma, -beta, -alpha) (:353-355): on every matrix
        Di = engine.wignerD_batch(F(twoj, 2), -gamma[lo:lo + 10000], -beta[lo:lo + 10000], -alpha[lo:lo + 10000])
        torch.cuda.synchronize()
        worst_i = max(worst_i, float((Di - Dc.conj().transpose(1, 2)).abs().max()))
        del Di
    assert worst_u < 5e-12, worst_u
    assert worst_i < TOL, worst_i
    # D(alpha, beta, gamma)[m, n] = exp(-i(m alpha + n gamma)) d(beta)[m, n]: the real-d kernel is an independent path
    pick = torch.from_numpy(np.random.default_rng(12).integers(0, nb, 512)).to(dev)
    d = engine.wignerd_batch(F(twoj, 2), beta[pick].contiguous())
    m = torch.arange(N, dtype=torch.float64, device=dev) - twoj / 2.0
    ph = torch.exp(-1j * (alpha[pick][:, None, None] * m[None, :, None] + gamma[pick][:, None, None] * m[None, None, :]))
    assert float((D[pick] - ph * d).abs().max()) < TOL
    # 96 sampled matrices against the oracle's C port (first, last, ragged tail and random)
    pk = np.concatenate([[0, 1, 15, 16, nb - 1, nb - 2, nb - 17], np.random.default_rng(13).integers(0, nb, 89)])
    ref = cref.wignerD_batch(twoj, alpha[pk].cpu().numpy(), beta[pk].cpu().numpy(), gamma[pk].cpu().numpy())
    assert np.abs(D[torch.from_numpy(pk).to(dev)].cpu().numpy() - ref).max() < TOL
    del D
    torch.cuda.empty_cache()


# ---- cfg-4: every integer j <= 512 on a 512-angle slab through wd_d_multi, output ring-buffered ----------------------
def test_full_size_cfg4_multi_ring_buffer(engine):
    """The device path of wd_d_multi for j = 0..512 on one GPU's 512-angle share of the 4096-point grid.  The 0.74 TB
    of output per GPU cannot be kept: groups of j share ONE reused region (ring buffer): inside a group the offsets are
    distinct, from group to group they alias -- and the last call aliases EVERY j onto offset 0, exactly what
    bench.py's cfg-4 does.  Every matrix of every j: trace and orthogonality; sampled j incl. the kernel hand-over
    points against the oracle."""
    import torch
    dev = "cuda:0"
    nb, jmax = 512, 512
    rank, world = 3, 8                                    # an inner slab of the 4096-point grid
    beta = (torch.arange(rank * nb, rank * nb + nb, dtype=torch.float64, device=dev) * (pi / (nb * world - 1))).contiguous()
    engine.prepare([2 * j for j in range(jmax + 1)])
    region_elems = int(2.2e9)                             # 17.6 GB ring region
    region = torch.empty(region_elems, dtype=torch.float64, device=dev)
    sample_js = {0, 1, 2, 50, 100, 101, 103, 104, 111, 112, 113, 200, 255, 256, 384, 511, 512}
    bh = beta.cpu().numpy()
    pick = np.array([0, 7, 8, 255, 256, 505, 511])
    j = 0
    worst = {"trace": 0.0, "orth": 0.0, "oracle": 0.0}
    while j <= jmax:
        group, offs, used = [], [], 0
        while j <= jmax and used + nb * (2 * j + 1) ** 2 <= region_elems:
            group.append(j); offs.append(used); used += nb * (2 * j + 1) ** 2; j += 1
        assert group, "region too small"
        _, _, views = engine.wignerd_multi(group, beta, out=region, offsets=np.array(offs, dtype=np.int64))
        torch.cuda.synchronize()
        for jj, v in zip(group, views):
            N = 2 * jj + 1
            tr = torch.diagonal(v, dim1=1, dim2=2).sum(-1)
            worst["trace"] = max(worst["trace"], float((tr - _trace_ref(torch, beta, 2 * jj)).abs().max()))
            for lo in range(0, nb, 128):
                vc = v[lo:lo + 128]
                worst["orth"] = max(worst["orth"], float((vc.transpose(1, 2) @ vc - torch.eye(N, dtype=torch.float64, device=dev)).abs().max()))
            if jj in sample_js:
                ref = cref.wignerd_batch(2 * jj, bh[pick])
                worst["oracle"] = max(worst["oracle"], float(np.abs(v[torch.from_numpy(pick).to(dev)].cpu().numpy() - ref).max()))
    assert worst["trace"] < 2e-10, worst            # N terms of size <= 1, each to ~1e-13
    assert worst["orth"] < 5e-12, worst
    assert worst["oracle"] < TOL, worst
    # bench.py's form: every j aliased onto offset 0 of one buffer; afterwards the buffer holds the last j's matrices
    js = list(range(jmax + 1))
    _, _, views = engine.wignerd_multi(js, beta, out=region, offsets=np.zeros(len(js), dtype=np.int64))
    torch.cuda.synchronize()
    last = views[-1]
    ref = cref.wignerd_batch(2 * jmax, bh[pick])
    assert np.abs(last[torch.from_numpy(pick).to(dev)].cpu().numpy() - ref).max() < TOL
    tr = torch.diagonal(last, dim1=1, dim2=2).sum(-1)
    assert float((tr - _trace_ref(torch, beta, 2 * jmax)).abs().max()) < 2e-10
    del region
    torch.cuda.empty_cache()


# ---- cfg-5: 10^7 random (j <= 200, m, n, beta) tuples ----------------------------------------------------------------
def test_full_size_cfg5_tuples(engine):
    import torch
    dev = "cuda:0"
    n, tjmax = 10_000_000, 400
    gen = torch.Generator(device=dev).manual_seed(3)
    tj = torch.randint(0, tjmax + 1, (n,), device=dev, generator=gen, dtype=torch.int32)
    u1 = torch.rand(n, device=dev, generator=gen, dtype=torch.float64)
    u2 = torch.rand(n, device=dev, generator=gen, dtype=torch.float64)
    tm = (-tj + 2 * torch.floor(u1 * (tj + 1).double()).to(torch.int32)).contiguous()
    tn = (-tj + 2 * torch.floor(u2 * (tj + 1).double()).to(torch.int32)).contiguous()
    beta = torch.rand(n, device=dev, generator=gen, dtype=torch.float64) * pi
    d = engine.wignerdjmn_batch(tj, tm, tn, beta)
    # symmetry fills of the reference, on every tuple (:277-292): d[m,n] = (-1)^(m-n) d[n,m] = d[-n,-m]
    d_t = engine.wignerdjmn_batch(tj, tn, tm, beta)
    d_r = engine.wignerdjmn_batch(tj, (-tn).contiguous(), (-tm).contiguous(), beta)
    torch.cuda.synchronize()
    sg = torch.where(((tm - tn) // 2) % 2 == 0, 1.0, -1.0).to(torch.float64)
    assert float((d - sg * d_t).abs().max()) < 1e-12
    assert float((d - d_r).abs().max()) < 1e-12
    assert float(d.abs().max()) <= 1.0 + 1e-12
    # 10^5-tuple sample against the C port of the reference loop
    pk = torch.from_numpy(np.random.default_rng(14).integers(0, n, 100000)).to(dev)
    ref = cref.wignerdjmn_batch(tj[pk].cpu().numpy(), tm[pk].cpu().numpy(), tn[pk].cpu().numpy(), beta[pk].cpu().numpy())
    assert np.abs(d[pk].cpu().numpy() - ref).max() < TOL
    # complex elements of the same tuples: D = d cis(-(alpha m + gamma n)) (:245)
    a = torch.rand(n, device=dev, generator=gen, dtype=torch.float64) * (2 * pi)
    g = torch.rand(n, device=dev, generator=gen, dtype=torch.float64) * (2 * pi)
    Dz = engine.wignerDjmn_batch(tj, tm, tn, a, beta, g)
    torch.cuda.synchronize()
    x = -(a * tm.double() / 2 + g * tn.double() / 2)
    assert float((Dz - d * torch.complex(torch.cos(x), torch.sin(x))).abs().max()) < TOL


# ---- beta <-> pi - beta pairing of symmetric grids (test/runtests.jl:330-334) --------------------------
@pytest.mark.parametrize("twoj,nb", [(200, 64), (200, 37), (200, 148 * 16 + 9), (20, 1000), (20, 999), (99, 48), (99, 33),
                                     (5, 16), (6, 17), (63, 250), (127, 40), (8, 5000)])
def test_pairing_symmetric_grids_vs_oracle(engine, twoj, nb):
    """linspace(0, pi, nb) is symmetric about pi/2: the column kernel writes matrix nb-1-b from the accumulators of b"""
    betas = np.linspace(0, pi, nb)
    got = engine.wignerd_batch(F(twoj, 2), betas)
    pick = np.unique(np.concatenate([[0, 1, nb // 2 - 1, nb // 2, (nb + 1) // 2, nb - 2, nb - 1], np.random.default_rng(twoj).integers(0, nb, 24)]))
    ref = cref.wignerd_batch(twoj, betas[pick])
    assert np.abs(got[pick] - ref).max() < TOL
    try:                                            # every matrix against the unpaired path
        engine.set_variant(5)
        unp = engine.wignerd_batch(F(twoj, 2), betas)
    finally:
        engine.set_variant(0)
    assert np.abs(got - unp).max() < 3e-13          # j * tol = 2e-13 is the bound of the pairing error (wd_k2_col.cuh)


def test_pairing_is_refused_off_tolerance(engine):
    """a grid that is symmetric only to 1e-9 must NOT be paired: results equal the unpaired path bit for bit"""
    nb = 400
    betas = np.linspace(0, pi, nb)
    betas[nb // 3] += 1e-9
    a = engine.wignerd_batch(100, betas)
    try:
        engine.set_variant(5)
        b = engine.wignerd_batch(100, betas)
    finally:
        engine.set_variant(0)
    assert np.array_equal(a, b)
    assert np.abs(a[[nb // 3, nb - 1 - nb // 3]] - cref.wignerd_batch(200, betas[[nb // 3, nb - 1 - nb // 3]])).max() < TOL
    # NaN in the grid: no pairing, NaN matrices only where the angle is NaN
    betas = np.linspace(0, pi, 64)
    betas[5] = np.nan
    d = engine.wignerd_batch(10, betas)
    assert np.isnan(d[5]).all() and np.isfinite(np.delete(d, 5, axis=0)).all()


@pytest.mark.parametrize("shift", [1, 2, 3, 5])
def test_output_at_odd_sector_phase(engine, shift):
    """the column kernel lays its staging rows out in the sector phase of the OUTPUT address: any 8-byte aligned
    destination must work (wd_d_multi offsets are arbitrary), and nothing outside the block may be touched"""
    import torch
    dev = "cuda:0"
    for twoj, nb in ((200, 40), (41, 70), (99, 33), (12, 300)):
        N = twoj + 1
        betas = torch.linspace(0.0, pi, nb, dtype=torch.float64, device=dev)
        buf = torch.full((nb * N * N + 64,), 777.0, dtype=torch.float64, device=dev)
        engine.wignerd_multi([F(twoj, 2)], betas, out=buf, offsets=np.array([shift], dtype=np.int64))
        torch.cuda.synchronize()
        assert bool((buf[:shift] == 777.0).all()) and bool((buf[shift + nb * N * N:] == 777.0).all())
        got = buf[shift:shift + nb * N * N].reshape(nb, N, N).transpose(1, 2).cpu().numpy()
        assert np.abs(got - o.wignerd_batch(F(twoj, 2), betas.cpu().numpy())).max() < TOL


# ---- multi-GPU handle, pageable hosts, persisted tables, sparse j ----------------------------------------------------
def test_multi_handle_shards_host_batches():
    import torch
    ndev = torch.cuda.device_count()
    devices = [0, 1] if ndev >= 2 else [0, 0]           # one GPU: two child handles on the same device
    eng = w.Engine(devices=devices)
    try:
        assert eng.is_multi
        rng = np.random.default_rng(21)
        betas = rng.uniform(0, pi, 203)
        assert np.abs(eng.wignerd_batch(20, betas) - o.wignerd_batch(20, betas)).max() < TOL
        betas = np.linspace(0, pi, 77)
        assert np.abs(eng.wignerd_batch(100, betas) - o.wignerd_batch(100, betas)).max() < TOL
        a, g = rng.uniform(0, 2 * pi, 77), rng.uniform(0, 2 * pi, 77)
        assert np.abs(eng.wignerD_batch(49.5, a, betas, g) - o.wignerD_batch(49.5, a, betas, g)).max() < TOL
        assert eng.wignerd_batch(3, np.zeros(0)).shape == (0, 7, 7)
        assert np.abs(eng.wignerd_batch(3, [0.4]) - o.wignerd_batch(3, [0.4])).max() < TOL       # fewer angles than devices
        n = 5001
        tj = rng.integers(0, 61, n).astype(np.int32)
        tm = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
        tn = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
        b = rng.uniform(0, pi, n)
        assert np.abs(eng.wignerdjmn_batch(tj, tm, tn, b) - o.wignerdjmn_batch(tj, tm, tn, b)).max() < TOL
        js = [0, 0.5, 3, 10.5, 40]
        betas = np.linspace(0.1, 3.0, 50)
        _, _, views = eng.wignerd_multi(js, betas)
        for j, v in zip(js, views):
            assert np.abs(v - o.wignerd_batch(j, betas)).max() < 1e-13
        assert np.abs(eng.wignerd(2, 0.3) - o.wignerd(2, 0.3)).max() < 1e-14
        with pytest.raises(w.ArgumentError):
            eng.wignerdjmn_batch(np.array([4, 4, 4, 4]), np.array([0, 0, 0, 6]), np.array([0, 0, 0, 0]), np.array([0.1, 0.2, 0.3, 0.4]))
        with pytest.raises(AssertionError):
            eng.wignerd_batch(2, torch.zeros(4, dtype=torch.float64, device="cuda:0"))       # device tensors: per-device engines
    finally:
        eng.close()


def test_pageable_and_pinned_hosts_agree(engine):
    """outputs above 8 MB into pageable memory go through the pinned bounce ring + host copy threads"""
    import torch
    betas = np.linspace(0, 3.0, 300)                      # 300 x 201 x 201 x 8 = 97 MB: several 32 MB pieces
    pageable = np.empty((300, 201, 201))
    pinned = torch.empty((300, 201, 201), dtype=torch.float64, pin_memory=True)
    engine.wignerd_batch(100, betas, out=pageable)
    engine.wignerd_batch(100, betas, out=pinned.numpy())
    assert np.array_equal(pageable, pinned.numpy())
    assert np.abs(pageable[[0, 150, 299]].transpose(0, 2, 1) - o.wignerd_batch(100, betas[[0, 150, 299]])).max() < TOL
    engine.set_slab_bytes(16 << 20)                       # several slabs AND several pieces
    try:
        again = np.empty((300, 201, 201))
        engine.wignerd_batch(100, betas, out=again)
        assert np.array_equal(again, pageable)
    finally:
        engine.set_slab_bytes(512 << 20)


def test_tables_save_load_skip_the_eigensolver(tmp_path):
    path = str(tmp_path / "tables.wdtb")
    a = w.Engine(0)
    b = w.Engine(0)
    try:
        a.prepare([0, 1, 7, 20, 99, 200, 300])
        ref = a.wignerd_batch(100, [0.3, 2.0])
        a.tables_save(path)
        assert os.path.getsize(path) > 8 * (101 * 106)
        n0 = b.launch_count()
        b.tables_load(path)
        assert b.launch_count() == n0                     # no eigensolver launch
        got = b.wignerd_batch(100, [0.3, 2.0])
        assert np.array_equal(got, ref)                   # the same tables, bit for bit
        n1 = b.launch_count()
        b.wignerd_batch(150, [0.3])                       # 2j = 300 comes from the file as well: one contraction call, no K1
        b.wignerd_batch(3.5, [0.3])
        assert np.abs(b.wignerd_batch(10, [1.1]) - o.wignerd_batch(10, [1.1])).max() < 1e-13
        assert b.launch_count() - n1 <= 12
        with open(path, "r+b") as f:                      # corrupted header: refused, nothing crashes
            f.write(b"garbage!")
        with pytest.raises(w.ArgumentError):
            b.tables_load(path)
        with pytest.raises(w.ArgumentError):
            b.tables_load(str(tmp_path / "missing.wdtb"))
    finally:
        a.close()
        b.close()


def test_jmn_prepares_only_the_j_present():
    """a single element at 2j = 2000 must not build (and cache) every table below it (21 GB of scratch)"""
    import torch
    eng = w.Engine(0)
    try:
        torch.cuda.synchronize()
        free0, _ = torch.cuda.mem_get_info(0)
        v = eng.wignerdjmn(1000, 3, -2, 0.7)
        free1, _ = torch.cuda.mem_get_info(0)
        assert free0 - free1 < 400e6, (free0 - free1)     # one table of 8 MB + workspace, not gigabytes
        d = eng.wignerd_batch(1000, [0.7])[0]
        assert abs(v - d[1000 + 3, 1000 - 2]) < 1e-13
        # mixed sparse j values in one batch
        tj = np.array([4000, 2, 4000, 1999, 0], np.int32)
        tm = np.array([-4000, 0, 10, 1, 0], np.int32)
        tn = np.array([4000, 2, -10, -1999, 0], np.int32)
        b = np.array([0.2, 0.3, 1.4, 2.0, 0.9])
        got = eng.wignerdjmn_batch(tj, tm, tn, b)
        assert abs(got[1] - o.wignerdjmn(1, 0, 1, 0.3)) < 1e-14 and got[4] == 1.0 and np.isfinite(got).all()
        assert abs(got[0] - np.sin(0.1) ** 4000) < 1e-13     # d^j_{-j,j} = (-1)^(2j) sin(beta/2)^(2j) ... tiny
    finally:
        eng.close()


def test_api_argument_checks(engine):
    """torch tuples default to int64; device / dtype / size of user buffers are validated"""
    import torch
    dev = "cuda:0"
    tj = torch.tensor([4, 6, 7], device=dev)                 # int64 on purpose
    tm = torch.tensor([0, -2, 1], device=dev)
    tn = torch.tensor([2, 6, -7], device=dev)
    b = torch.tensor([0.3, 1.0, 2.0], dtype=torch.float64, device=dev)
    got = engine.wignerdjmn_batch(tj, tm, tn, b).cpu().numpy()
    ref = o.wignerdjmn_batch(np.array([4, 6, 7]), np.array([0, -2, 1]), np.array([2, 6, -7]), np.array([0.3, 1.0, 2.0]))
    assert np.abs(got - ref).max() < 1e-14
    with pytest.raises(AssertionError):
        engine.wignerdjmn_batch(tj, tm, tn, b.cpu().numpy())                     # mixed memory spaces
    with pytest.raises(AssertionError):
        engine.wignerDjmn_batch(tj, tm, tn, b.cpu().numpy(), b, b)               # host alpha with device tuples
    with pytest.raises(AssertionError):
        engine.wignerdjmn_batch(tj, tm, tn, b[:2].contiguous())                   # length mismatch
    with pytest.raises(AssertionError):
        engine.wignerd_batch(2, b, out=torch.empty((3, 5, 5), dtype=torch.float32, device=dev))
    with pytest.raises(AssertionError):
        engine.wignerd_batch(2, b, out=torch.empty((2, 5, 5), dtype=torch.float64, device=dev))      # too small
    with pytest.raises(AssertionError):
        engine.wignerd_batch(2, np.array([0.1, 0.2]), out=np.empty((2, 5, 5), dtype=np.float32))
    with pytest.raises(AssertionError):
        engine.wignerd_batch(2, b.float())
    with pytest.raises(AssertionError):
        engine.wignerdjmn_batch(tj.double(), tm, tn, b)


# ---- fused consumer: rotation of spherical-harmonic coefficients (docs/src/index.md:76-84) -------------
def _rotate_ref(lmax, alm, a, b, g, ls=None):
    """a'_{l m'} = sum_m D^l_{m m'} a_{l m} from the oracle's wignerD, for the l in ls (all by default)"""
    out = {}
    for l in (range(lmax + 1) if ls is None else ls):
        D = cref.wignerD_batch(2 * l, a, b, g) if l > 8 else o.wignerD_batch(l, a, b, g)      # [rot, m, m']
        al = alm[..., l * l:(l + 1) * (l + 1)]
        out[l] = np.einsum("bmn,bm->bn", D, np.broadcast_to(al, (len(b), 2 * l + 1)))
    return out


@pytest.mark.parametrize("lmax,nb,shared", [(0, 5, True), (1, 7, False), (8, 33, True), (8, 40, False), (64, 50, True), (64, 19, False)])
def test_rotate_sh_vs_oracle(engine, lmax, nb, shared):
    rng = np.random.default_rng(100 + lmax + nb)
    L2 = (lmax + 1) ** 2
    alm = (rng.normal(size=(L2,) if shared else (nb, L2)) + 1j * rng.normal(size=(L2,) if shared else (nb, L2))) / np.sqrt(2)
    a, g = rng.uniform(0, 2 * pi, nb), rng.uniform(0, 2 * pi, nb)
    b = rng.uniform(0, pi, nb)
    got = engine.rotate_sh(lmax, alm, a, b, g)
    assert got.shape == (nb, L2)
    ref = _rotate_ref(lmax, alm, a, b, g)
    for l, r in ref.items():
        assert np.abs(got[:, l * l:(l + 1) * (l + 1)] - r).max() < 2e-12 * np.sqrt(2 * l + 1), l      # sums of 2l+1 terms of size ~1
    # a rotation preserves the norm of every l block (unitarity of D^l)
    for l in range(lmax + 1):
        n_in = np.linalg.norm(np.broadcast_to(alm[..., l * l:(l + 1) ** 2], (nb, 2 * l + 1)), axis=1)
        assert np.abs(np.linalg.norm(got[:, l * l:(l + 1) ** 2], axis=1) - n_in).max() < 1e-11 * (1 + n_in.max())


def test_rotate_sh_lmax512_device_path(engine):
    """cfg-4's consumer at its size: lmax = 512 on a 512-rotation slab, device buffers; sampled l against the oracle, norm
    preservation and the composition rule R(-gamma, -beta, -alpha) R(alpha, beta, gamma) = 1 on everything"""
    import torch
    dev = "cuda:0"
    lmax, nb = 512, 512
    L2 = (lmax + 1) ** 2
    gen = torch.Generator(device=dev).manual_seed(5)
    alm = torch.complex(torch.randn(L2, dtype=torch.float64, device=dev, generator=gen), torch.randn(L2, dtype=torch.float64, device=dev, generator=gen))
    a = torch.rand(nb, dtype=torch.float64, device=dev, generator=gen) * (2 * pi)
    g = torch.rand(nb, dtype=torch.float64, device=dev, generator=gen) * (2 * pi)
    b = torch.linspace(0.0, pi, nb, dtype=torch.float64, device=dev)
    out = engine.rotate_sh(lmax, alm, a, b, g)
    back = engine.rotate_sh(lmax, out, -g, -b, -a)                 # one coefficient set per rotation: the inverse rotation
    torch.cuda.synchronize()
    assert float((back - alm[None, :]).abs().max()) < 2e-10        # 1025 terms of size ~3, twice
    lidx = torch.repeat_interleave(torch.arange(lmax + 1, device=dev), 2 * torch.arange(lmax + 1, device=dev) + 1)
    n_in = torch.zeros(lmax + 1, dtype=torch.float64, device=dev).index_add_(0, lidx, alm.abs() ** 2)
    n_out = torch.zeros((nb, lmax + 1), dtype=torch.float64, device=dev).index_add_(1, lidx, out.abs() ** 2)
    assert float(((n_out - n_in[None, :]).abs() / (1 + n_in[None, :])).max()) < 1e-11
    pick = np.array([0, 1, 17, 255, 256, 511])
    ah, bh, gh, almh = a.cpu().numpy()[pick], b.cpu().numpy()[pick], g.cpu().numpy()[pick], alm.cpu().numpy()
    ref = _rotate_ref(lmax, almh, ah, bh, gh, ls=[0, 1, 2, 100, 111, 112, 256, 511, 512])
    oh = out[torch.from_numpy(pick).to(dev)].cpu().numpy()
    for l, r in ref.items():
        assert np.abs(oh[:, l * l:(l + 1) * (l + 1)] - r).max() < 1e-11 * np.sqrt(2 * l + 1), l
    del out, back
    torch.cuda.empty_cache()


def test_rotate_sh_multi_handle_and_errors(engine):
    import torch
    rng = np.random.default_rng(7)
    lmax, nb = 12, 37
    L2 = (lmax + 1) ** 2
    alm = rng.normal(size=(nb, L2)) + 1j * rng.normal(size=(nb, L2))
    a, b, g = rng.uniform(0, 6, nb), rng.uniform(0, 3, nb), rng.uniform(0, 6, nb)
    ref = engine.rotate_sh(lmax, alm, a, b, g)
    eng = w.Engine(devices=[0, 1] if torch.cuda.device_count() >= 2 else [0, 0])
    try:
        assert np.abs(eng.rotate_sh(lmax, alm, a, b, g) - ref).max() < 1e-13
        assert np.abs(eng.rotate_sh(lmax, alm[0], a, b, g) - engine.rotate_sh(lmax, alm[0], a, b, g)).max() < 1e-13
    finally:
        eng.close()
    with pytest.raises(AssertionError):
        engine.rotate_sh(lmax, alm[:, :-1], a, b, g)
    with pytest.raises(AssertionError):
        engine.rotate_sh(-1, alm, a, b, g)
    assert engine.rotate_sh(3, np.zeros(16, complex), np.zeros(0), np.zeros(0), np.zeros(0)).shape == (0, 16)


@pytest.mark.parametrize("twoj,nb", [(241, 40), (300, 45), (600, 33), (1024, 34), (401, 64)])
def test_pairing_in_the_streaming_kernel(engine, twoj, nb):
    """2j > 223: the mu-streaming kernel contracts only the first half of a symmetric grid, its store warps write the mirror matrices"""
    betas = np.linspace(0, pi, nb)
    got = engine.wignerd_batch(F(twoj, 2), betas)
    pick = np.array([0, 1, nb // 2 - 1, nb // 2, (nb + 1) // 2, nb - 2, nb - 1])
    assert np.abs(got[pick] - cref.wignerd_batch(twoj, betas[pick])).max() < TOL
    try:
        engine.set_variant(6)                       # the same kernel without pairing: every matrix
        unp = engine.wignerd_batch(F(twoj, 2), betas)
    finally:
        engine.set_variant(0)
    assert np.abs(got - unp).max() < 6e-13          # j * tol <= 5e-13 bounds the pairing error (pair_tol in wd_api.cu)
    assert np.array_equal(got[: (nb + 1) // 2], unp[: (nb + 1) // 2])      # the first half is computed exactly as before


@pytest.mark.parametrize("twoj,nb", [(241, 21), (300, 16), (401, 9)])
def test_complex_D_beyond_the_resident_kernel(engine, twoj, nb):
    """2j > 223: real d through the streaming DMMA kernel, then the phase pass with the reference's exact cis(-(m alpha + n gamma))"""
    rng = np.random.default_rng(twoj)
    a, g, b = rng.uniform(0, 2 * pi, nb), rng.uniform(0, 2 * pi, nb), rng.uniform(0, pi, nb)
    got = engine.wignerD_batch(F(twoj, 2), a, b, g)
    assert np.abs(got - cref.wignerD_batch(twoj, a, b, g)).max() < TOL
    try:
        engine.set_variant(1)                       # FP64-ALU kernel: independent implementation
        ref = engine.wignerD_batch(F(twoj, 2), a, b, g)
    finally:
        engine.set_variant(0)
    assert np.abs(got - ref).max() < 1e-12


def test_engines_on_two_devices_in_one_thread():
    """kernel attributes (opt-in shared memory) are per device: a second engine on another GPU must set them again"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    e0, e1 = w.Engine(0), w.Engine(1)
    try:
        betas = np.linspace(0, pi, 300)
        for eng in (e0, e1, e0):
            assert np.abs(eng.wignerd_batch(100, betas)[[0, 150, 299]] - o.wignerd_batch(100, betas[[0, 150, 299]])).max() < TOL
            assert np.abs(eng.wignerd_batch(150, betas[:20]) - cref.wignerd_batch(300, betas[:20])).max() < TOL
    finally:
        e0.close(); e1.close()
