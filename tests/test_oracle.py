"""Pins the CPU oracle (oracle/) before anything is compared against it: README known answers,
the reference's closed forms / invariants (test/runtests.jl), exact multiprecision sums, and
C port == numpy restatement.  CPU only."""
import json
import os
from fractions import Fraction as F
from math import pi

import numpy as np
import pytest

import reference_suite as rs
from oracle import cref
from oracle import wignerd_oracle as o

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class OracleImpl:
    wignerd = staticmethod(o.wignerd)
    wignerD = staticmethod(o.wignerD)
    wignerdjmn = staticmethod(o.wignerdjmn)
    wignerDjmn = staticmethod(o.wignerDjmn)
    Equator, NorthPole, SouthPole = o.Equator, o.NorthPole, o.SouthPole


def test_readme_known_answers():
    g = json.load(open(os.path.join(GOLD, "readme_vectors.json")))
    d = o.wignerd(0.5, 0)
    assert np.allclose(d, np.array(g["wignerd_0.5_0"]), atol=1e-15)
    assert np.signbit(d[0, 1]) and not np.signbit(d[1, 0])            # the printed -0.0 / 0.0 (README.md:15-16)
    assert np.allclose(o.wignerd(1, pi / 3), np.array(g["wignerd_1_pi_3"]), atol=5e-7)   # printed to 6 digits
    assert o.wignerdjmn(1, 1, 1, pi / 3) == g["wignerdjmn_1_1_1_pi_3"]                   # bit-exact
    D = o.wignerD(1, 0, pi / 3, pi / 2)
    ref = np.array(g["wignerD_1_0_pi_3_pi_2"])
    assert np.allclose(D.real, ref[..., 0], atol=5e-7, rtol=1e-5) and np.allclose(D.imag, ref[..., 1], atol=5e-7)
    assert np.allclose(D.real[ref[..., 0] < 1e-10], ref[..., 0][ref[..., 0] < 1e-10], rtol=1e-5, atol=0)
    z = o.wignerDjmn(1, 1, 1, 0, pi / 3, pi / 2)
    assert [z.real, z.imag] == g["wignerDjmn_1_1_1_0_pi_3_pi_2"]                          # bit-exact


def test_mp_exact_samples():
    g = json.load(open(os.path.join(GOLD, "mp_exact_samples.json")))
    worst = {}
    for s in g["samples"]:
        beta, exact = float.fromhex(s["beta"]), float.fromhex(s["d"])
        got = float(o._wignerdjmn(s["twoj"], s["twom"], s["twon"], beta, o.jy_eigen(s["twoj"])))
        worst[s["twoj"]] = max(worst.get(s["twoj"], 0.0), abs(got - exact))
    for tj, e in worst.items():
        assert e < (5e-13 if tj > 400 else 1e-13), (tj, e)


def test_special_jmn():
    rs.check_special_jmn(OracleImpl)


def test_Djmn_vs_djmn():
    rs.check_Djmn_vs_djmn(OracleImpl)


def test_closed_forms():
    rs.check_closed_forms(OracleImpl)


def test_symmetries():
    rs.check_symmetries(OracleImpl, nbeta=4)


def test_wignerD():
    rs.check_wignerD(OracleImpl, ngrid=4)


def test_cartesian():
    rs.check_cartesian(OracleImpl, ngrid=6)


def test_sincos_equator():
    for a in range(-10, 11):
        for x in (a, F(a, 2)):
            s, c = o.sincos_equator(x)
            assert abs(s - np.sin(float(x) * pi / 2)) < 1e-14 and abs(c - np.cos(float(x) * pi / 2)) < 1e-14


def test_errors():
    with pytest.raises(ValueError):
        o.wignerd(0.3, 0.1)            # InexactError (:130)
    with pytest.raises(ValueError):
        o.wignerdjmn(1, 2, 0, 0.1)     # ArgumentError (:191)
    with pytest.raises(ValueError):
        o.wignerdjmn(1, 0.5, 0, 0.1)


@pytest.mark.parametrize("twoj", [0, 1, 2, 5, 20, 99])
def test_c_port_equals_numpy_oracle(twoj):
    rng = np.random.default_rng(twoj)
    b = rng.uniform(-7, 7, 5)
    a, g = rng.uniform(0, 6, 5), rng.uniform(0, 6, 5)
    c = cref.wignerd_batch(twoj, b)
    p = np.stack([o.wignerd(F(twoj, 2), x) for x in b])
    assert np.array_equal(c, p)
    C = cref.wignerD_batch(twoj, a, b, g)
    P = np.stack([o.wignerD(F(twoj, 2), a[i], b[i], g[i]) for i in range(5)])
    assert np.abs(C - P).max() < 1e-15
    # the vectorised batch helper is the same arithmetic up to summation-order-free identities
    assert np.abs(o.wignerd_batch(F(twoj, 2), b) - p).max() < 1e-13


def test_c_port_jmn():
    rng = np.random.default_rng(3)
    n = 300
    tj = rng.integers(0, 41, n)
    tm = -tj + 2 * rng.integers(0, tj + 1)
    tn = -tj + 2 * rng.integers(0, tj + 1)
    b = rng.uniform(0, pi, n)
    got = cref.wignerdjmn_batch(tj, tm, tn, b)
    ref = o.wignerdjmn_batch(tj, tm, tn, b)
    assert np.abs(got - ref).max() < 1e-14
