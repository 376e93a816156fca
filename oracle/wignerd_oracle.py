"""CPU oracle for the Wigner-d/D hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain numpy/scipy restatement of the algorithm of the reference
(jishnub/WignerD.jl v0.1.4, ``src/WignerD.jl``).  It is the checker the parity
tests compare the CUDA path against.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it;
the product path (``wignerd.jl_b200``) never does.

Parity pinning: Julia is not installed in this image, so the reference itself
cannot be executed here.  The oracle is pinned instead against every golden
value the reference publishes for this path -- README.md:13-34 -- and against
every closed form / invariant of test/runtests.jl (see tests/test_oracle.py),
plus exact multiprecision Wigner sums (oracle/mp_exact.py) at the BASELINE j.

Third-party arithmetic the reference reaches outside /root/reference:
``LinearAlgebra.eigen!(::Hermitian{ComplexF64})`` = LAPACK ``zheevr`` (Julia
stdlib, no pinned patch version; call site src/WignerD.jl:161).  The same
LAPACK routine is reached here through ``scipy.linalg.eigh(driver="evr")``.

All ``file:line`` references are to /root/reference/src/WignerD.jl.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np
import scipy.linalg as _sl

# --------------------------------------------------------------------------
# special co-latitudes (src/WignerD.jl:14-46)
# --------------------------------------------------------------------------
GENERIC, NORTH, SOUTH, EQUATOR = 0, 1, 2, 3


class SpecialCoLatitude(float):
    """A float subclass so that it can be used wherever a Real beta is."""
    kind = GENERIC

    def __new__(cls):
        return super().__new__(cls, cls._value)


class NorthPole(SpecialCoLatitude):      # :35-37
    _value = 0.0
    kind = NORTH


class SouthPole(SpecialCoLatitude):      # :38
    _value = float(np.pi)
    kind = SOUTH


class Equator(SpecialCoLatitude):        # :39
    _value = float(np.pi / 2)
    kind = EQUATOR


def beta_kind(beta) -> int:
    return getattr(beta, "kind", GENERIC)


def sincos_equator(alpha):
    """``_sincos(alpha, Equator())`` (:49-84): exact (sin, cos) of alpha*pi/2
    for integer alpha (mod 4 table) and half-integer alpha (mod 8 table)."""
    two_a = Fraction(alpha) * 2
    if two_a.denominator != 1:
        raise ValueError("alpha must be an integer or a half-integer")
    two_a = int(two_a)
    r = 1.0 / math.sqrt(2.0)
    if two_a % 2 == 0:
        table = [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)]
        c, s = table[(two_a // 2) % 4]
    else:
        table = [(1.0, 0.0), (r, r), (0.0, 1.0), (-r, r),
                 (-1.0, 0.0), (-r, -r), (0.0, -1.0), (r, -r)]
        c, s = table[two_a % 8]
    return s, c


# --------------------------------------------------------------------------
# j plumbing: everything is keyed by the integer 2j, like UInt(2j+1) at :164
# --------------------------------------------------------------------------
def twoj_of(j) -> int:
    """Int(2j): raises like the reference's ``Int(2j + 1)`` (:130, InexactError)
    when 2j is not an integer, and when j < 0 (:96)."""
    try:
        tj = Fraction(j) * 2
    except TypeError:
        tj = Fraction(float(j)) * 2
    if tj.denominator != 1:
        raise ValueError(f"InexactError: 2j+1 is not an integer for j = {j}")
    tj = int(tj)
    if tj < 0:
        raise ValueError("j must be >= 0")
    return tj


def X(j: float, n: float) -> float:
    """:145  X(j, n) = sqrt((j + n)*(j - n + 1))."""
    return math.sqrt((j + n) * (j - n + 1))


def coeffi(twoj: int) -> np.ndarray:
    """:147-156  J_y in the J_z basis: zero matrix with super-diagonal
    ``Jy[k, k+1] = -X(j, -j+k)/2im`` (1-based k = 1..2j); only the upper
    triangle is meaningful (``Hermitian(Jy)`` wraps the upper triangle)."""
    N = twoj + 1
    j = twoj / 2.0
    Jy = np.zeros((N, N), dtype=np.complex128)
    for k in range(1, N):
        Jy[k - 1, k] = -X(j, -j + k) / 2j
    return Jy


_EIGEN_CACHE: dict[int, np.ndarray] = {1: np.ones((1, 1), dtype=np.complex128)}   # :137-138


def jy_eigen(twoj: int) -> np.ndarray:
    """:163-176  eigenvectors of J_y, ascending eigenvalues (= -j..j), stored
    along ROWS: ``vs[mu_idx, m_idx] = <m|mu>`` -- ``permutedims`` (:173), i.e. a
    plain transpose, NOT an adjoint.  Cached by 2j+1 like the reference."""
    key = twoj + 1
    if key in _EIGEN_CACHE:
        return _EIGEN_CACHE[key]
    Jy = coeffi(twoj)
    # eigen!(Hermitian(Jy), sortby=identity)  ->  LAPACK zheevr('V','A','U')
    _, v = _sl.eigh(Jy, lower=False, driver="evr")
    vs = np.ascontiguousarray(v.T)
    _EIGEN_CACHE[key] = vs
    return vs


def _terms(vs: np.ndarray, mind, nind):
    """T1, T2 of :200-201 for index arrays mind, nind (0-based)."""
    re, im = vs.real, vs.imag
    a_re, a_im = re[:, mind], im[:, mind]
    b_re, b_im = re[:, nind], im[:, nind]
    T1 = a_re * b_re + a_im * b_im
    T2 = a_im * b_re - a_re * b_im
    return T1, T2


def _wignerdjmn(twoj: int, twom, twon, beta: float, vs: np.ndarray):
    """:187-206  d^j_mn(beta) = sum_mu [cos(-mu*beta)*T1 - sin(-mu*beta)*T2],
    accumulated sequentially over mu = -j..j exactly as the reference loop
    (vectorised over the (m, n) arrays only)."""
    N = twoj + 1
    twom = np.asarray(twom)
    twon = np.asarray(twon)
    if np.any((twom + twoj) % 2) or np.any((twon + twoj) % 2):
        raise ValueError("ArgumentError: m/n inconsistent with j")
    mind = (twom + twoj) // 2
    nind = (twon + twoj) // 2
    if np.any(mind < 0) or np.any(mind >= N) or np.any(nind < 0) or np.any(nind >= N):
        raise ValueError("ArgumentError: m/n is inconsistent with axes(v, 2)")   # :190-193
    T1, T2 = _terms(vs, mind, nind)
    acc = np.zeros(np.broadcast(mind, nind).shape, dtype=np.float64)
    b = float(beta)
    for k in range(N):
        mu = (2 * k - twoj) / 2.0
        x = -mu * b                      # :199  sincos(-mu * float(beta))
        s, c = math.sin(x), math.cos(x)
        acc = acc + (c * T1[k] - s * T2[k])     # :202
    return acc


def wignerdjmn(j, m, n, beta) -> float:
    """:214-237  single element, with the special-angle dispatch."""
    twoj = twoj_of(j)
    twom = int(Fraction(m) * 2)
    twon = int(Fraction(n) * 2)
    kind = beta_kind(beta)
    if kind == NORTH:                                  # :218-222
        return 1.0 if twom == twon else 0.0
    if kind == SOUTH:                                  # :224-228
        if twom == -twon:
            return 1.0 if ((twoj - twon) // 2) % 2 == 0 else -1.0
        return 0.0
    vs = jy_eigen(twoj)
    if kind == EQUATOR:                                # :230-237
        jm_odd = ((twoj + twom) // 2) % 2 == 1
        jn_odd = ((twoj + twon) // 2) % 2 == 1
        if (jm_odd and twon == 0) or (jn_odd and twom == 0):
            # the index checks of :190-193 are skipped here by the reference too
            return 0.0
    return float(_wignerdjmn(twoj, twom, twon, float(beta), vs))


def wignerDjmn(j, m, n, alpha, beta, gamma) -> complex:
    """:245  d * cis(-(alpha*m + gamma*n))."""
    d = wignerdjmn(j, m, n, beta)
    x = -(float(alpha) * float(Fraction(m)) + float(gamma) * float(Fraction(n)))
    return d * complex(math.cos(x), math.sin(x))


def wignerd(j, beta, out: np.ndarray | None = None) -> np.ndarray:
    """:266-324  full matrix.  ``d[j+m, j+n] = d^j_mn`` (0-based here), built
    like the reference: direct wedge (:271-273), then the three symmetry
    passes in the reference's order (:277-292) -- which is what produces the
    ``-0.0`` of README.md:15."""
    twoj = twoj_of(j)
    N = twoj + 1
    d = np.zeros((N, N), dtype=np.float64) if out is None else out
    kind = beta_kind(beta)
    vs = jy_eigen(twoj)             # :307 runs the eigen step even for the poles
    if kind == NORTH:               # :247-253 (does not zero the off-diagonal)
        d[np.arange(N), np.arange(N)] = 1.0
        return d
    if kind == SOUTH:               # :255-264
        even = True
        for i in range(N):          # m = -j..j ; dj[m, -m]
            d[i, N - 1 - i] = 1.0 if even else -1.0
            even = not even
        return d

    # direct wedge: n in -j..0 (2n <= 0), m in n..-n
    tm, tn = [], []
    for twon in range(-twoj, 1, 2):
        for twom in range(twon, -twon + 1, 2):
            tm.append(twom)
            tn.append(twon)
    tm = np.array(tm, dtype=np.int64)
    tn = np.array(tn, dtype=np.int64)
    if len(tm):
        vals = _wignerdjmn(twoj, tm, tn, float(beta), vs)
        if kind == EQUATOR:         # :230-237
            jm_odd = ((twoj + tm) // 2) % 2 == 1
            jn_odd = ((twoj + tn) // 2) % 2 == 1
            vals = np.where((jm_odd & (tn == 0)) | (jn_odd & (tm == 0)), 0.0, vals)
        d[(tm + twoj) // 2, (tn + twoj) // 2] = vals

    def idx(t):
        return (t + twoj) // 2

    def sgn(tm_, tn_):              # (-1)^(m-n), m-n integer
        return -1.0 if ((tm_ - tn_) // 2) % 2 else 1.0

    start = 2 if twoj % 2 == 0 else 1        # _half_or_one_to(j): 1..j or 1/2..j
    # right wedge  dj[m,n] = (-1)^(m-n) dj[-m,-n]     :277-281
    for twon in range(start, twoj + 1, 2):
        for twom in range(-twon, twon + 1, 2):
            d[idx(twom), idx(twon)] = sgn(twom, twon) * d[idx(-twom), idx(-twon)]
    # bottom wedge dj[m,n] = (-1)^(m-n) dj[n,m]       :284-288
    for twom in range(start, twoj + 1, 2):
        for twon in range(-twom, twom + 1, 2):
            d[idx(twom), idx(twon)] = sgn(twom, twon) * d[idx(twon), idx(twom)]
    # top wedge    dj[m,n] = dj[-n,-m]                :290-292
    # (m in _unitrange(-j, -1): for half-integer j this stops at m = -3/2)
    for twom in range(-twoj, -1, 2):
        for twon in range(twom, -twom + 1, 2):
            d[idx(twom), idx(twon)] = d[idx(-twon), idx(-twom)]
    return d


def wignerD(j, alpha, beta, gamma) -> np.ndarray:
    """:335-357  D[m,n] = d[m,n] * cis(-(m*alpha + n*gamma))."""
    twoj = twoj_of(j)
    N = twoj + 1
    D = np.zeros((N, N), dtype=np.complex128)
    D.real[...] = 0.0
    dre = wignerd(j, beta)
    D[...] = dre
    a, g = float(alpha), float(gamma)
    for ni in range(N):
        n = (2 * ni - twoj) / 2.0
        for mi in range(N):
            m = (2 * mi - twoj) / 2.0
            x = -(m * a + n * g)                       # :339
            D[mi, ni] *= complex(math.cos(x), math.sin(x))
    return D


# --------------------------------------------------------------------------
# batched helpers used by the parity tests / bench (same arithmetic, numpy)
# --------------------------------------------------------------------------
def wignerd_batch(j, betas) -> np.ndarray:
    """(nb, N, N) array, ``out[b, j+m, j+n]``; generic-angle path only.
    Same per-element arithmetic as :func:`wignerd` but vectorised over the
    full matrix and the angle batch (sum over mu still sequential)."""
    twoj = twoj_of(j)
    N = twoj + 1
    betas = np.asarray(betas, dtype=np.float64).ravel()
    vs = jy_eigen(twoj)
    re, im = vs.real, vs.imag
    out = np.zeros((len(betas), N, N))
    for k in range(N):
        mu = (2 * k - twoj) / 2.0
        x = -mu * betas
        c, s = np.cos(x), np.sin(x)
        T1 = np.outer(re[k], re[k]) + np.outer(im[k], im[k])
        T2 = np.outer(im[k], re[k]) - np.outer(re[k], im[k])
        out += c[:, None, None] * T1[None] - s[:, None, None] * T2[None]
    return out


def wignerD_batch(j, alphas, betas, gammas) -> np.ndarray:
    twoj = twoj_of(j)
    N = twoj + 1
    d = wignerd_batch(j, betas)
    m = (2 * np.arange(N) - twoj) / 2.0
    a = np.asarray(alphas, dtype=np.float64).ravel()
    g = np.asarray(gammas, dtype=np.float64).ravel()
    x = -(m[None, :, None] * a[:, None, None] + m[None, None, :] * g[:, None, None])
    return d * (np.cos(x) + 1j * np.sin(x))


def wignerdjmn_batch(twoj, twom, twon, betas) -> np.ndarray:
    """vector of d^j_mn(beta) for tuple arrays (generic angles)."""
    twoj = np.asarray(twoj)
    out = np.zeros(len(twoj))
    for tj in np.unique(twoj):
        sel = np.nonzero(twoj == tj)[0]
        vs = jy_eigen(int(tj))
        N = int(tj) + 1
        mi = (np.asarray(twom)[sel] + tj) // 2
        ni = (np.asarray(twon)[sel] + tj) // 2
        T1, T2 = _terms(vs, mi, ni)
        b = np.asarray(betas, dtype=np.float64)[sel]
        acc = np.zeros(len(sel))
        for k in range(N):
            mu = (2 * k - tj) / 2.0
            x = -mu * b
            acc += np.cos(x) * T1[k] - np.sin(x) * T2[k]
        out[sel] = acc
    return out
