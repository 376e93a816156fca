"""ctypes binding of oracle/wignerd_ref.c (the C port of the reference hot
loop) -- TEST / BASELINE INFRASTRUCTURE ONLY.  Used by tests (port == numpy
oracle) and by bench.py's ``cpu_baseline`` / ``--impl reference`` legs."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import wignerd_oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwignerd_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "wignerd_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.wdref_num_threads.restype = ctypes.c_int
        L.wdref_wignerd_batch.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int64, dp, ctypes.c_int]
        L.wdref_wignerD_batch.argtypes = [ctypes.c_int, dp, dp, dp, dp, dp, ctypes.c_int64, dp, dp, ctypes.c_int]
        L.wdref_wignerdjmn_batch.argtypes = [ip, ip, ip, dp, ctypes.c_int64,
                                             ctypes.POINTER(dp), ctypes.POINTER(dp), dp, ctypes.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _soa(twoj):
    """the reference's cached StructArray (.re, .im), column-major [mu, m]."""
    vs = _o.jy_eigen(twoj)                     # vs[mu, m]
    re = np.asfortranarray(vs.real).ravel(order="F").copy()
    im = np.asfortranarray(vs.imag).ravel(order="F").copy()
    return re, im


def num_threads() -> int:
    return lib().wdref_num_threads()


def wignerd_batch(twoj: int, betas, nthreads: int = 0) -> np.ndarray:
    betas = np.ascontiguousarray(betas, dtype=np.float64)
    N = twoj + 1
    re, im = _soa(twoj)
    out = np.empty((len(betas), N, N))          # [b][n][m] in memory
    lib().wdref_wignerd_batch(twoj, _dp(re), _dp(im), _dp(betas), len(betas), _dp(out), nthreads)
    return out.transpose(0, 2, 1)               # view indexed [b, m, n]


def wignerD_batch(twoj: int, alphas, betas, gammas, nthreads: int = 0) -> np.ndarray:
    a = np.ascontiguousarray(alphas, dtype=np.float64)
    b = np.ascontiguousarray(betas, dtype=np.float64)
    g = np.ascontiguousarray(gammas, dtype=np.float64)
    N = twoj + 1
    re, im = _soa(twoj)
    out = np.empty((len(b), N, N), dtype=np.complex128)
    lib().wdref_wignerD_batch(twoj, _dp(re), _dp(im), _dp(a), _dp(b), _dp(g), len(b),
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), None, nthreads)
    return out.transpose(0, 2, 1)


def wignerdjmn_batch(twoj, twom, twon, betas, nthreads: int = 0) -> np.ndarray:
    twoj = np.ascontiguousarray(twoj, dtype=np.int32)
    twom = np.ascontiguousarray(twom, dtype=np.int32)
    twon = np.ascontiguousarray(twon, dtype=np.int32)
    betas = np.ascontiguousarray(betas, dtype=np.float64)
    tjmax = int(twoj.max())
    dp = ctypes.POINTER(ctypes.c_double)
    res = (dp * (tjmax + 1))()
    ims = (dp * (tjmax + 1))()
    keep = []
    for tj in np.unique(twoj):
        re, im = _soa(int(tj))
        keep.append((re, im))
        res[int(tj)] = _dp(re)
        ims[int(tj)] = _dp(im)
    out = np.empty(len(twoj))
    ip = ctypes.POINTER(ctypes.c_int32)
    lib().wdref_wignerdjmn_batch(twoj.ctypes.data_as(ip), twom.ctypes.data_as(ip), twon.ctypes.data_as(ip),
                                 _dp(betas), len(twoj), res, ims, _dp(out), nthreads)
    return out
