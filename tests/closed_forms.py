"""Closed forms of d^j_mn(theta) for j = 0, 1/2, 1, 3/2, 2, restated from the reference's
test/runtests.jl:129 (j=0), :154-158 (1/2), :183-193 (1), :218-236 (3/2), :261-289 (2).
Each returns {(m, n): value} with m, n as Fractions."""
from fractions import Fraction as F
from math import cos, sin, sqrt

h = F(1, 2)


def d0(t):
    return {(F(0), F(0)): 1.0}


def d_half(t):
    return {(-h, -h): cos(t / 2), (h, -h): -sin(t / 2), (-h, h): sin(t / 2), (h, h): cos(t / 2)}


def d1(t):
    r2 = sqrt(2)
    return {
        (F(-1), F(-1)): (1 + cos(t)) / 2, (F(0), F(-1)): -sin(t) / r2, (F(1), F(-1)): (1 - cos(t)) / 2,
        (F(-1), F(0)): sin(t) / r2, (F(0), F(0)): cos(t), (F(1), F(0)): -sin(t) / r2,
        (F(-1), F(1)): (1 - cos(t)) / 2, (F(0), F(1)): sin(t) / r2, (F(1), F(1)): (1 + cos(t)) / 2,
    }


def d3half(t):
    c, s, r3 = cos(t / 2), sin(t / 2), sqrt(3)
    a, b = 3 * h, h
    return {
        (-a, -a): c**3, (-b, -a): -r3 * s * c**2, (b, -a): r3 * s**2 * c, (a, -a): -s**3,
        (-a, -b): r3 * s * c**2, (-b, -b): c * (3 * c**2 - 2), (b, -b): s * (3 * s**2 - 2), (a, -b): r3 * s**2 * c,
        (-a, b): r3 * s**2 * c, (-b, b): -s * (3 * s**2 - 2), (b, b): c * (3 * c**2 - 2), (a, b): -r3 * s * c**2,
        (-a, a): s**3, (-b, a): r3 * s**2 * c, (b, a): r3 * s * c**2, (a, a): c**3,
    }


def d2(t):
    c, s = cos(t), sin(t)
    r32 = sqrt(3 / 2)
    return {
        (F(-2), F(-2)): (1 + c)**2 / 4, (F(-1), F(-2)): -s * (1 + c) / 2, (F(0), F(-2)): 0.5 * r32 * s**2,
        (F(1), F(-2)): -s * (1 - c) / 2, (F(2), F(-2)): (1 - c)**2 / 4,
        (F(-2), F(-1)): s * (1 + c) / 2, (F(-1), F(-1)): (2 * c**2 + c - 1) / 2, (F(0), F(-1)): -r32 * s * c,
        (F(1), F(-1)): -(2 * c**2 - c - 1) / 2, (F(2), F(-1)): -s * (1 - c) / 2,
        (F(-2), F(0)): 0.5 * r32 * s**2, (F(-1), F(0)): r32 * s * c, (F(0), F(0)): 0.5 * (3 * c**2 - 1),
        (F(1), F(0)): -r32 * s * c, (F(2), F(0)): 0.5 * r32 * s**2,
        (F(-2), F(1)): s * (1 - c) / 2, (F(-1), F(1)): -(2 * c**2 - c - 1) / 2, (F(0), F(1)): r32 * s * c,
        (F(1), F(1)): (2 * c**2 + c - 1) / 2, (F(2), F(1)): -s * (1 + c) / 2,
        (F(-2), F(2)): (1 - c)**2 / 4, (F(-1), F(2)): s * (1 - c) / 2, (F(0), F(2)): 0.5 * r32 * s**2,
        (F(1), F(2)): s * (1 + c) / 2, (F(2), F(2)): (1 + c)**2 / 4,
    }


CLOSED = {F(0): d0, h: d_half, F(1): d1, 3 * h: d3half, F(2): d2}


def isapprox(a, b, atol=1e-14, rtol=1e-8):
    """Julia's isapprox(a, b; atol, rtol): |a-b| <= max(atol, rtol*max(|a|,|b|))."""
    return abs(a - b) <= max(atol, rtol * max(abs(a), abs(b)))
