"""Exact (multiprecision) Wigner small-d elements -- TEST INFRASTRUCTURE ONLY.

Independent triangulation for the oracle at j far above what the reference's
own tests reach (test/runtests.jl stops at j = 5).  Uses the explicit Wigner
sum in the convention the reference documents (docs/src/index.md:11-33,
Varshalovich et al. 1988):

  d^j_{m'm}(b) = sum_k (-1)^(k-m+m') sqrt((j+m)!(j-m)!(j+m')!(j-m')!)
                 / ((j+m-k)! k! (j-k-m')! (k-m+m')!)
                 * cos(b/2)^(2j-2k+m-m') * sin(b/2)^(2k-m+m')

evaluated with mpmath at enough digits to survive the cancellation, with the
binomial-type coefficients formed as exact integers.  ``beta`` is taken as the
exact binary value of the float64 passed in.
"""
from __future__ import annotations

from math import factorial

import mpmath as mp


def wigner_d_exact(twoj: int, twomp: int, twom: int, beta: float) -> float:
    """d^j_{m' m}(beta) with j = twoj/2, m' = twomp/2 (row), m = twom/2 (col)."""
    assert (twoj + twomp) % 2 == 0 and (twoj + twom) % 2 == 0
    jpm = (twoj + twom) // 2      # j + m
    jmm = (twoj - twom) // 2      # j - m
    jpmp = (twoj + twomp) // 2    # j + m'
    jmmp = (twoj - twomp) // 2    # j - m'
    dm = (twomp - twom) // 2      # m' - m
    with mp.workdps(60 + int(0.7 * twoj)):
        b = mp.mpf(beta)
        c = mp.cos(b / 2)
        s = mp.sin(b / 2)
        pref = mp.sqrt(mp.mpf(factorial(jpm) * factorial(jmm) * factorial(jpmp) * factorial(jmmp)))
        kmin = max(0, -dm)
        kmax = min(jpm, jmmp)
        tot = mp.mpf(0)
        for k in range(kmin, kmax + 1):
            den = factorial(jpm - k) * factorial(k) * factorial(jmmp - k) * factorial(k + dm)
            term = pref / den * c ** (twoj - 2 * k - dm) * s ** (2 * k + dm)
            tot += -term if (k + dm) % 2 else term
        return float(tot)
