"""The reference's own test strategy (test/runtests.jl), restated once and run against any
implementation exposing the reference API: the CPU oracle (tests/test_oracle.py) and the CUDA
product through the C ABI (tests/test_gpu_parity.py, -m gpu).  ``impl`` needs:
wignerd, wignerD, wignerdjmn, wignerDjmn, Equator, NorthPole, SouthPole."""
from fractions import Fraction as F
from math import cos, pi, sqrt

import numpy as np

from closed_forms import CLOSED, isapprox

SQRT_EPS = sqrt(np.finfo(float).eps)
h = F(1, 2)


def mrange(j):
    j = F(j)
    return [-j + k for k in range(int(2 * j) + 1)]


def idx(j, m):
    return int(F(j) + F(m))


def check_special_jmn(impl, js=None):
    """test/runtests.jl:46-93  pi/2, 0, pi  ==  Equator(), NorthPole(), SouthPole()."""
    js = js if js is not None else [F(k) for k in range(4)] + [F(k, 2) for k in (1, 3, 5, 7)]
    for j in js:
        for m in mrange(j):
            for n in mrange(j):
                for beta, sp in ((pi / 2, impl.Equator()), (0, impl.NorthPole()), (pi, impl.SouthPole())):
                    a = impl.wignerdjmn(j, m, n, beta)
                    b = impl.wignerdjmn(j, m, n, sp)
                    assert isapprox(a, b, atol=1e-14, rtol=SQRT_EPS), (j, m, n, sp, a, b)


def check_Djmn_vs_djmn(impl, js=None, nbeta=10):
    """test/runtests.jl:95-116"""
    js = js if js is not None else [F(k) for k in range(4)] + [F(k, 2) for k in (1, 3, 5, 7)]
    for j in js:
        for beta in np.linspace(0, pi, nbeta):
            for m in mrange(j):
                for n in mrange(j):
                    D = impl.wignerDjmn(j, m, n, 0, beta, 0)
                    d = impl.wignerdjmn(j, m, n, beta)
                    assert isapprox(D, d, atol=1e-14, rtol=SQRT_EPS), (j, m, n, beta, D, d)


def check_closed_forms(impl, ntheta=100):
    """test/runtests.jl:119-310: every element's closed form at 100 theta + the special angles,
    trace = sum cos(m theta), det = 1."""
    for j, form in CLOSED.items():
        thetas = list(np.linspace(0, pi, ntheta)) + [impl.Equator(), impl.NorthPole(), impl.SouthPole()]
        for theta in thetas:
            d = impl.wignerd(j, theta)
            for (m, n), val in form(float(theta)).items():
                got = d[idx(j, m), idx(j, n)]
                assert isapprox(got, val, atol=1e-14, rtol=1e-8), (j, m, n, float(theta), got, val)
            tr = sum(cos(float(m) * float(theta)) for m in mrange(j))
            assert isapprox(np.trace(d), tr, atol=1e-14, rtol=SQRT_EPS)
            assert abs(np.linalg.det(d) - 1) < 1e-8


def check_symmetries(impl, jmax2=10, nbeta=10):
    """test/runtests.jl:312-336"""
    for beta in np.linspace(0, pi, nbeta):
        for tj in range(0, jmax2 + 1):
            j = F(tj, 2)
            N = tj + 1
            d = impl.wignerd(j, beta)
            assert np.allclose(impl.wignerd(j, -beta).T, d, atol=1e-13, rtol=1e-8)
            for k in range(1, 5):
                sg = (-1.0) ** (tj * k)
                assert np.allclose(impl.wignerd(j, beta + 2 * pi * k), sg * d, atol=1e-12, rtol=1e-8)
                assert np.allclose(impl.wignerd(j, beta - 2 * pi * k), sg * d, atol=1e-12, rtol=1e-8)
                for sign in (+1, -1):
                    d2 = impl.wignerd(j, beta + sign * (2 * k + 1) * pi)
                    for m in mrange(j):
                        for n in mrange(j):
                            e = sign * (2 * k + 1) * j - n          # integer
                            ph = -1.0 if int(e) % 2 else 1.0
                            assert isapprox(d2[idx(j, m), idx(j, n)], ph * d[idx(j, m), idx(j, -n)],
                                            atol=1e-13, rtol=1e-8), (j, m, n, beta, k, sign)
            d3 = impl.wignerd(j, pi - beta)
            for m in mrange(j):
                for n in mrange(j):
                    p1 = -1.0 if int(j + m) % 2 else 1.0
                    p2 = -1.0 if int(j - n) % 2 else 1.0
                    assert isapprox(d3[idx(j, m), idx(j, n)], p1 * d[idx(j, m), idx(j, -n)], atol=1e-13, rtol=1e-8)
                    assert isapprox(d3[idx(j, m), idx(j, n)], p2 * d[idx(j, -m), idx(j, n)], atol=1e-13, rtol=1e-8)
            assert d.shape == (N, N)


def check_wignerD(impl, jmax2=6, ngrid=10, elementwise=True):
    """test/runtests.jl:338-359"""
    for beta in np.linspace(0, pi, ngrid):
        for tj in range(0, jmax2 + 1):
            j = F(tj, 2)
            N = tj + 1
            d = impl.wignerd(j, beta)
            D0 = impl.wignerD(j, 0, beta, 0)
            assert np.array_equal(d, D0), (j, beta)                   # d == D exactly (:343)
            for alpha in np.linspace(0, 2 * pi, ngrid):
                for gamma in np.linspace(0, 2 * pi, ngrid):
                    D = impl.wignerD(j, alpha, beta, gamma)
                    if elementwise:
                        for n in mrange(j):
                            for m in mrange(j):
                                e = impl.wignerDjmn(j, m, n, alpha, beta, gamma)
                                assert isapprox(D[idx(j, m), idx(j, n)], e, atol=1e-14, rtol=SQRT_EPS)
                    assert abs(np.linalg.det(D) - 1) < 1e-8
                    assert np.allclose(D.conj().T @ D, np.eye(N), atol=1e-12)
                    Dinv = impl.wignerD(j, -gamma, -beta, -alpha)
                    assert np.allclose(np.linalg.inv(D), Dinv, atol=1e-10)
                    assert np.allclose(D.conj().T, Dinv, atol=1e-12)


def rot_zyz(a, b, g):
    def rz(t):
        return np.array([[np.cos(t), -np.sin(t), 0], [np.sin(t), np.cos(t), 0], [0, 0, 1]])

    def ry(t):
        return np.array([[np.cos(t), 0, np.sin(t)], [0, 1, 0], [-np.sin(t), 0, np.cos(t)]])
    return rz(a) @ ry(b) @ rz(g)


def check_cartesian(impl, ngrid=10):
    """test/runtests.jl:367-388: conj(D^1(a,b,g)) == U RotZYZ(a,b,g) U' pins the Euler/sign convention."""
    r = 1 / sqrt(2)
    U = np.array([[r, -1j * r, 0], [0, 0, 1], [-r, -1j * r, 0]])
    for a in np.linspace(0, 2 * pi, ngrid):
        for b in np.linspace(0, 2 * pi, ngrid):
            for g in np.linspace(0, 2 * pi, ngrid):
                D = impl.wignerD(1, a, b, g)
                assert np.allclose(D.conj(), U @ rot_zyz(a, b, g) @ U.conj().T, atol=1e-12)
