"""Host-side mirror of the reference's Julia API over the C ABI (ctypes).

The reference (jishnub/WignerD.jl v0.1.4) is a Julia module; Julia is not in this
image, so the host layer above ``libwignerd_b200.so`` is written in Python with the
reference's names, argument meaning and error behaviour (src/WignerD.jl:9-12, :214-245,
:306-357), so that the parity tests read like test/runtests.jl.  The same binding for
Julia (``ccall``) is in ``julia/WignerD.jl`` / INTEGRATION.md.

There is no CPU path here: if the shared library is missing or no B200 is visible,
calls raise.  Nothing in this module imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import math
import os
import threading
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WIGNERD_B200_LIB") or os.path.join(_HERE, "libwignerd_b200.so")   # env override: kernel experiments

WD_OK, WD_ERR_ARGUMENT, WD_ERR_ASSERT, WD_ERR_CUDA, WD_ERR_NOMEM, WD_ERR_UNSUPPORTED = range(6)
GENERIC, NORTH, SOUTH, EQUATOR = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1

# every symbol include/wignerd_b200.h declares (tests check the .so exports each one)
ABI_SYMBOLS = [
    "wd_create", "wd_destroy", "wd_last_error", "wd_abi_version", "wd_set_stream", "wd_synchronize",
    "wd_launch_count", "wd_cache_clear", "wd_prepare", "wd_eigvecs", "wd_d_batch", "wd_D_batch",
    "wd_d_matrix", "wd_D_matrix", "wd_d_multi", "wd_djmn_batch", "wd_Djmn_batch", "wd_set_variant",
    "wd_fp64_peak", "wd_k2_shared_bytes", "wd_create_multi", "wd_multi_size", "wd_multi_child", "wd_shard_bounds",
    "wd_gather", "wd_tables_save", "wd_tables_load", "wd_set_slab_bytes", "wd_rotate_sh",
]


class ArgumentError(ValueError):
    """reference: ArgumentError (src/WignerD.jl:98, :191-194)"""


class InexactError(ValueError):
    """reference: InexactError from Int(2j + 1) (src/WignerD.jl:130)"""


class CudaError(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen the in-tree library; fails loudly (no fallback) when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m wignerd.jl_b200.build` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
        L.wd_last_error.restype = ctypes.c_char_p
        L.wd_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
        L.wd_destroy.argtypes = [vp]
        L.wd_set_stream.argtypes = [vp, vp]
        L.wd_synchronize.argtypes = [vp]
        L.wd_launch_count.argtypes = [vp]
        L.wd_launch_count.restype = i64
        L.wd_cache_clear.argtypes = [vp]
        L.wd_prepare.argtypes = [vp, vp, i32]
        L.wd_eigvecs.argtypes = [vp, i32, vp]
        L.wd_d_batch.argtypes = [vp, i32, vp, i64, vp, ctypes.c_int]
        L.wd_D_batch.argtypes = [vp, i32, vp, vp, vp, i64, vp, ctypes.c_int]
        L.wd_d_matrix.argtypes = [vp, i32, dbl, ctypes.c_int, vp]
        L.wd_D_matrix.argtypes = [vp, i32, dbl, dbl, ctypes.c_int, dbl, vp]
        L.wd_d_multi.argtypes = [vp, vp, i32, vp, i64, vp, vp, ctypes.c_int]
        L.wd_djmn_batch.argtypes = [vp, vp, vp, vp, vp, i64, ctypes.c_int, vp, ctypes.c_int]
        L.wd_Djmn_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, ctypes.c_int, vp, ctypes.c_int]
        L.wd_set_variant.argtypes = [vp, ctypes.c_int]
        L.wd_k2_shared_bytes.argtypes = [i32, ctypes.c_int]
        L.wd_k2_shared_bytes.restype = i64
        L.wd_fp64_peak.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(dbl)]
        L.wd_create_multi.argtypes = [ctypes.c_int, vp, ctypes.POINTER(vp)]
        L.wd_multi_size.argtypes = [vp]
        L.wd_multi_child.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp)]
        L.wd_shard_bounds.argtypes = [i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(i64), ctypes.POINTER(i64)]
        L.wd_gather.argtypes = [vp, vp, vp, ctypes.c_int, vp]
        L.wd_tables_save.argtypes = [vp, ctypes.c_char_p]
        L.wd_tables_load.argtypes = [vp, ctypes.c_char_p]
        L.wd_set_slab_bytes.argtypes = [vp, i64]
        L.wd_rotate_sh.argtypes = [vp, i32, vp, i64, vp, vp, vp, i64, vp, ctypes.c_int]
        _lib = L
        return L


def _check(rc: int):
    if rc == WD_OK:
        return
    msg = load_library().wd_last_error().decode("utf-8", "replace")
    if rc == WD_ERR_ARGUMENT:
        raise ArgumentError(msg)
    if rc == WD_ERR_ASSERT:
        raise AssertionError(msg)
    if rc == WD_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == WD_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise CudaError(msg)


# ----------------------------------------------------------------------------------------
# special co-latitudes (src/WignerD.jl:14-46): Real subtypes usable wherever beta is
# ----------------------------------------------------------------------------------------
class SpecialCoLatitudes(float):
    kind = GENERIC

    def __new__(cls):
        return super().__new__(cls, cls._value)

    def __repr__(self):
        return f"{type(self).__name__}()"


class NorthPole(SpecialCoLatitudes):
    _value = 0.0
    kind = NORTH


class SouthPole(SpecialCoLatitudes):
    _value = math.pi
    kind = SOUTH


class Equator(SpecialCoLatitudes):
    _value = math.pi / 2
    kind = EQUATOR


def _beta_kind(beta) -> int:
    return getattr(beta, "kind", GENERIC)


def _sincos(alpha, _equator: Equator):
    """``WignerD._sincos(alpha, Equator())`` (:49-84): exact (sin, cos) of alpha*pi/2."""
    two_a = Fraction(alpha) * 2
    if two_a.denominator != 1:
        raise InexactError("alpha must be an integer or a half-integer")
    two_a = int(two_a)
    r = 1.0 / math.sqrt(2.0)
    if two_a % 2 == 0:
        c, s = [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)][(two_a // 2) % 4]
    else:
        c, s = [(1.0, 0.0), (r, r), (0.0, 1.0), (-r, r), (-1.0, 0.0), (-r, -r), (0.0, -1.0), (r, -r)][two_a % 8]
    return s, c


# ----------------------------------------------------------------------------------------
# j / m plumbing
# ----------------------------------------------------------------------------------------
def twoj_of(j) -> int:
    """Int(2j); InexactError when 2j is not an integer (:130), AssertionError when j < 0 (:96)."""
    try:
        tj = Fraction(j) * 2
    except TypeError:
        tj = Fraction(float(j)) * 2
    if tj.denominator != 1:
        raise InexactError(f"Int(2j + 1): 2j + 1 is not an integer for j = {j}")
    tj = int(tj)
    if tj < 0:
        raise AssertionError("j must be >= 0")
    return tj


def _two_of(x, what) -> int:
    try:
        t = Fraction(x) * 2
    except TypeError:
        t = Fraction(float(x)) * 2
    if t.denominator != 1:
        raise ArgumentError(f"{what} = {x} is not an integer or half-integer")
    return int(t)


def _check_mn(twoj, twom, twon, m, n):
    """the index checks of _wignerdjmn (:190-193)."""
    for t, v, name in ((twom, m, "m"), (twon, n, "n")):
        if (t + twoj) % 2 or abs(t) > twoj:
            raise ArgumentError(f"{name} = {v} is inconsistent with j = {Fraction(twoj, 2)}")


class WignerMatrix:
    """``WignerD.WignerMatrix`` / ``_offsetmatrix`` (:91-120, :158-159): index a (2j+1)x(2j+1)
    matrix by (m, n) in -j..j, integer or half-integer."""

    def __init__(self, j, D):
        self.twoj = twoj_of(j)
        D = np.asarray(D) if not isinstance(D, np.ndarray) else D
        if D.ndim != 2 or D.shape[0] != self.twoj + 1 or D.shape[1] != self.twoj + 1:
            raise AssertionError(f"size of matrix must be {self.twoj + 1}x{self.twoj + 1}")
        self.D = D

    def _index(self, m):
        t = _two_of(m, "index")
        if (t + self.twoj) % 2 or abs(t) > self.twoj:
            raise IndexError(f"index {m} out of range for j = {Fraction(self.twoj, 2)}")
        return (t + self.twoj) // 2

    def __getitem__(self, mn):
        m, n = mn
        return self.D[self._index(m), self._index(n)]

    def __setitem__(self, mn, val):
        m, n = mn
        self.D[self._index(m), self._index(n)] = val

    @property
    def shape(self):
        return self.D.shape

    def parent(self):
        return self.D


def _offsetmatrix(j, D):
    return WignerMatrix(j, D)


# ----------------------------------------------------------------------------------------
# engine = one wd_handle (GPU + stream + per-j table cache); one default engine per thread,
# like the reference's per-Task JyEigenDict (src/WignerD.jl:137-138)
# ----------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return ctypes.c_void_p(x.data_ptr())
    return ctypes.c_void_p(x.ctypes.data)


class Engine:
    """one wd_handle.  ``Engine(device)``: one GPU.  ``Engine(devices=[0, 1, ...])``: a multi-GPU handle
    (wd_create_multi) -- the batch methods then take HOST arrays and shard them over the devices."""

    def __init__(self, device: int = 0, devices=None):
        self._lib = load_library()
        h = ctypes.c_void_p()
        if devices is not None:
            devs = np.ascontiguousarray(list(devices), dtype=np.int32)
            _check(self._lib.wd_create_multi(len(devs), _ptr(devs), ctypes.byref(h)))
            self.devices = [int(d) for d in devs]
            device = self.devices[0]
        else:
            _check(self._lib.wd_create(int(device), ctypes.byref(h)))
            self.devices = [int(device)]
        self._h = h
        self.device = int(device)
        self._prepared: set[int] = set()

    @property
    def is_multi(self) -> bool:
        return int(self._lib.wd_multi_size(self._h)) > 1 or len(self.devices) > 1

    def tables_save(self, path: str):
        """persist every cached eigenvector table (wd_tables_save)"""
        _check(self._lib.wd_tables_save(self._h, os.fsencode(path)))

    def tables_load(self, path: str):
        """load persisted tables into the cache: a cold process skips the eigensolver (wd_tables_load)"""
        _check(self._lib.wd_tables_load(self._h, os.fsencode(path)))

    def set_slab_bytes(self, nbytes: int):
        _check(self._lib.wd_set_slab_bytes(self._h, int(nbytes)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.wd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        _check(self._lib.wd_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))

    def _use_torch_stream(self, device):
        """enqueue on torch's current stream; the legacy default stream is passed as cudaStreamLegacy (0x1)
        because 0 means "the handle's own stream" in wd_set_stream."""
        import torch
        self.set_stream(torch.cuda.current_stream(device).cuda_stream or 1)

    def synchronize(self):
        _check(self._lib.wd_synchronize(self._h))

    def launch_count(self) -> int:
        return int(self._lib.wd_launch_count(self._h))

    def cache_clear(self):
        _check(self._lib.wd_cache_clear(self._h))
        self._prepared.clear()

    def set_variant(self, variant: int):
        _check(self._lib.wd_set_variant(self._h, int(variant)))

    def fp64_peak(self, kind: int = 0, iters: int = 20000) -> float:
        out = ctypes.c_double()
        _check(self._lib.wd_fp64_peak(self._h, kind, iters, ctypes.byref(out)))
        return out.value

    def prepare(self, twoj_list):
        arr = np.ascontiguousarray(twoj_list, dtype=np.int32)
        _check(self._lib.wd_prepare(self._h, _ptr(arr), len(arr)))
        self._prepared.update(int(t) for t in arr)

    def eigvecs(self, j) -> np.ndarray:
        twoj = twoj_of(j)
        N = twoj + 1
        u = np.zeros((N, N), order="F")
        _check(self._lib.wd_eigvecs(self._h, twoj, _ptr(u)))
        self._prepared.add(twoj)
        return u

    def _check_Jy(self, twoj, Jy):
        """preallocated-Jy asserts of coeffi (:148-149); like the reference they are only
        evaluated on a cache miss.  The scratch itself is not needed by this engine."""
        if Jy is None or twoj in self._prepared:
            return
        Jy = np.asarray(Jy)
        if Jy.dtype != np.complex128:
            raise AssertionError("preallocated Jy matrix must have an element type of ComplexF64")
        if Jy.shape != (twoj + 1, twoj + 1):
            raise AssertionError(f"preallocated Jy matrix must be of size ({twoj + 1}, {twoj + 1})")

    # -- batched matrices ---------------------------------------------------------------
    def _check_device(self, t, what):
        if len(self.devices) > 1:
            raise AssertionError("a multi-GPU engine takes host (numpy) arrays; device tensors go to per-device engines")
        if not t.is_cuda or t.device.index != self.device:
            raise AssertionError(f"{what} must live on cuda:{self.device} (this engine's device), got {t.device}")

    def _angles(self, x, n=None, like=None):
        if _is_torch(x):
            if str(x.dtype) != "torch.float64" or not x.is_contiguous():
                raise AssertionError("device angle arrays must be contiguous float64 tensors")
            self._check_device(x, "angle arrays")
            return x, MEM_DEVICE
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
        return a, MEM_HOST

    def _check_out(self, out, mem, shape, cplx):
        """a caller-supplied out= must have exactly the layout the library writes: dtype, element count, contiguity,
        and the memory space / device of the inputs (a mismatch would be an out-of-bounds or cross-device write)."""
        n = int(np.prod(shape))
        if mem == MEM_DEVICE:
            if not _is_torch(out):
                raise AssertionError("out= must be a torch tensor when the inputs are device tensors")
            self._check_device(out, "out=")
            want = "torch.complex128" if cplx else "torch.float64"
            if str(out.dtype) != want or not out.is_contiguous() or out.numel() != n:
                raise AssertionError(f"out= must be a contiguous {want} tensor with {n} elements")
        else:
            if not isinstance(out, np.ndarray):
                raise AssertionError("out= must be a numpy array when the inputs are host arrays")
            want = np.complex128 if cplx else np.float64
            if out.dtype != want or not out.flags.c_contiguous or out.size != n or not out.flags.writeable:
                raise AssertionError(f"out= must be a writable C-contiguous {np.dtype(want).name} array with {n} elements")
        return out

    def wignerd_batch(self, j, betas, out=None):
        """d^j(beta_b) for all b: array ``res[b, j+m, j+n]`` (memory: nb column-major matrices back to
        back).  numpy in -> numpy out (host path, copies staged by the library); torch cuda tensor in ->
        torch tensor out on the same device, enqueued on the current torch stream, not synchronised."""
        twoj = twoj_of(j)
        N = twoj + 1
        b, mem = self._angles(betas)
        nb = b.numel() if mem == MEM_DEVICE else b.size
        if mem == MEM_DEVICE:
            import torch
            if out is None:
                out = torch.empty((nb, N, N), dtype=torch.float64, device=b.device)
            self._use_torch_stream(b.device)
        elif out is None:
            out = np.empty((nb, N, N), dtype=np.float64)
        out = self._check_out(out, mem, (nb, N, N), False).reshape(nb, N, N)
        _check(self._lib.wd_d_batch(self._h, twoj, _ptr(b), nb, _ptr(out), mem))
        self._prepared.add(twoj)
        return out.transpose(1, 2) if mem == MEM_DEVICE else out.transpose(0, 2, 1)

    def wignerD_batch(self, j, alphas, betas, gammas, out=None):
        twoj = twoj_of(j)
        N = twoj + 1
        a, mem = self._angles(alphas)
        b, mem_b = self._angles(betas)
        g, mem_g = self._angles(gammas)
        if not (mem == mem_b == mem_g):
            raise AssertionError("alpha, beta, gamma must live in the same memory space")
        nb = b.numel() if mem == MEM_DEVICE else b.size
        na = a.numel() if mem == MEM_DEVICE else a.size
        ng = g.numel() if mem == MEM_DEVICE else g.size
        if not (na == nb == ng):
            raise AssertionError("alpha, beta, gamma must have the same length")
        if mem == MEM_DEVICE:
            import torch
            if out is None:
                out = torch.empty((nb, N, N), dtype=torch.complex128, device=b.device)
            self._use_torch_stream(b.device)
        elif out is None:
            out = np.empty((nb, N, N), dtype=np.complex128)
        out = self._check_out(out, mem, (nb, N, N), True).reshape(nb, N, N)
        _check(self._lib.wd_D_batch(self._h, twoj, _ptr(a), _ptr(b), _ptr(g), nb, _ptr(out), mem))
        self._prepared.add(twoj)
        return out.transpose(1, 2) if mem == MEM_DEVICE else out.transpose(0, 2, 1)

    def wignerd_multi(self, js, betas, out=None, offsets=None):
        """all d^j for j in js on one angle grid.  Returns (flat buffer, offsets[k], views[k])."""
        twojs = np.ascontiguousarray([twoj_of(j) for j in js], dtype=np.int32)
        b, mem = self._angles(betas)
        nb = b.numel() if mem == MEM_DEVICE else b.size
        sizes = (twojs.astype(np.int64) + 1) ** 2 * nb
        if offsets is None:
            offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        total = int((offsets + sizes).max()) if len(sizes) else 0
        if mem == MEM_DEVICE:
            import torch
            if out is None:
                out = torch.empty(total, dtype=torch.float64, device=b.device)
            self._use_torch_stream(b.device)
        elif out is None:
            out = np.empty(total, dtype=np.float64)
        if len(offsets) != len(twojs) or (len(offsets) and int(offsets.min()) < 0):
            raise AssertionError("offsets must hold one non-negative entry per j")
        n_out = out.numel() if mem == MEM_DEVICE else out.size
        if n_out < total:
            raise AssertionError(f"out= has {n_out} elements, the offsets need {total}")
        out = self._check_out(out, mem, (n_out,), False)
        _check(self._lib.wd_d_multi(self._h, _ptr(twojs), len(twojs), _ptr(b), nb, _ptr(out), _ptr(offsets), mem))
        self._prepared.update(int(t) for t in twojs)
        views = []
        for k, tj in enumerate(twojs):
            N = int(tj) + 1
            v = out[int(offsets[k]):int(offsets[k]) + nb * N * N].reshape(nb, N, N)
            views.append(v.transpose(1, 2) if mem == MEM_DEVICE else v.transpose(0, 2, 1))
        return out, offsets, views

    def rotate_sh(self, lmax: int, alm, alphas, betas, gammas, out=None):
        """rotated spherical-harmonic coefficients ``res[b, l^2 + l + m'] = sum_m D^l_{m m'}(alpha_b, beta_b, gamma_b)
        alm[(b,) l^2 + l + m]`` for l <= lmax (docs/src/index.md:76-84), without forming a D matrix.  ``alm``: complex128,
        shape ((lmax+1)^2,) -- one set rotated by every rotation -- or (nb, (lmax+1)^2)."""
        lmax = int(lmax)
        L2 = (lmax + 1) ** 2
        a, mem = self._angles(alphas)
        b, mem_b = self._angles(betas)
        g, mem_g = self._angles(gammas)
        if not (mem == mem_b == mem_g):
            raise AssertionError("alpha, beta, gamma must live in the same memory space")
        nb = b.numel() if mem == MEM_DEVICE else b.size
        if not ((a.numel() if mem == MEM_DEVICE else a.size) == nb == (g.numel() if mem == MEM_DEVICE else g.size)):
            raise AssertionError("alpha, beta, gamma must have the same length")
        if mem == MEM_DEVICE:
            import torch
            if not _is_torch(alm) or str(alm.dtype) != "torch.complex128" or not alm.is_contiguous():
                raise AssertionError("alm must be a contiguous complex128 tensor when the angles are device tensors")
            self._check_device(alm, "alm")
            n_alm = alm.numel()
        else:
            alm = np.ascontiguousarray(alm, dtype=np.complex128)
            n_alm = alm.size
        if n_alm == L2:
            stride = 0
        elif n_alm == nb * L2:
            stride = L2
        else:
            raise AssertionError(f"alm must hold (lmax+1)^2 = {L2} or nb x (lmax+1)^2 = {nb * L2} coefficients, got {n_alm}")
        if mem == MEM_DEVICE:
            if out is None:
                out = torch.empty((nb, L2), dtype=torch.complex128, device=b.device)
            self._use_torch_stream(b.device)
        elif out is None:
            out = np.empty((nb, L2), dtype=np.complex128)
        out = self._check_out(out, mem, (nb, L2), True).reshape(nb, L2)
        _check(self._lib.wd_rotate_sh(self._h, lmax, _ptr(alm), stride, _ptr(a), _ptr(b), _ptr(g), nb, _ptr(out), mem))
        self._prepared.update(range(0, 2 * lmax + 1, 2))
        return out

    # -- batched single elements ----------------------------------------------------------
    def _tuples(self, js, ms, ns):
        if _is_torch(js) or _is_torch(ms) or _is_torch(ns):
            import torch
            if not (_is_torch(js) and _is_torch(ms) and _is_torch(ns)):
                raise AssertionError("2j, 2m, 2n must all be device tensors or all host arrays")
            res = []
            for t, name in ((js, "2j"), (ms, "2m"), (ns, "2n")):
                self._check_device(t, name)
                if t.dtype.is_floating_point or t.dtype == torch.bool:
                    raise AssertionError(f"{name} must be an integer tensor")
                # torch integer tensors default to int64: the ABI reads int32
                res.append(t.to(torch.int32).contiguous().reshape(-1))
            if not (res[0].numel() == res[1].numel() == res[2].numel()):
                raise AssertionError("2j, 2m, 2n must have the same length")
            return res[0], res[1], res[2], MEM_DEVICE
        tj = np.ascontiguousarray(js, dtype=np.int32).ravel()
        tm = np.ascontiguousarray(ms, dtype=np.int32).ravel()
        tn = np.ascontiguousarray(ns, dtype=np.int32).ravel()
        if not (tj.size == tm.size == tn.size):
            raise AssertionError("2j, 2m, 2n must have the same length")
        return tj, tm, tn, MEM_HOST

    def _tuple_angles(self, x, n, mem, what):
        a, mem_a = self._angles(x)
        if mem_a != mem:
            raise AssertionError(f"tuples and {what} must live in the same memory space")
        if (a.numel() if mem == MEM_DEVICE else a.size) != n:
            raise AssertionError(f"{what} must have one entry per tuple ({n})")
        return a

    def wignerdjmn_batch(self, twoj, twom, twon, betas, beta_kind=GENERIC, out=None):
        """tuples given as 2j, 2m, 2n int32 arrays (so half-integers need no special type)."""
        tj, tm, tn, mem = self._tuples(twoj, twom, twon)
        n = tj.numel() if mem == MEM_DEVICE else tj.size
        b = None
        if betas is not None:
            b = self._tuple_angles(betas, n, mem, "beta")
        elif int(beta_kind) == GENERIC:
            raise AssertionError("beta is required for generic co-latitudes")
        if mem == MEM_DEVICE:
            import torch
            if out is None:
                out = torch.empty(n, dtype=torch.float64, device=tj.device)
            self._use_torch_stream(tj.device)
        elif out is None:
            out = np.empty(n, dtype=np.float64)
        out = self._check_out(out, mem, (n,), False)
        _check(self._lib.wd_djmn_batch(self._h, _ptr(tj), _ptr(tm), _ptr(tn), _ptr(b), n, int(beta_kind), _ptr(out), mem))
        return out

    def wignerDjmn_batch(self, twoj, twom, twon, alphas, betas, gammas, beta_kind=GENERIC, out=None):
        tj, tm, tn, mem = self._tuples(twoj, twom, twon)
        n = tj.numel() if mem == MEM_DEVICE else tj.size
        a = self._tuple_angles(alphas, n, mem, "alpha")
        g = self._tuple_angles(gammas, n, mem, "gamma")
        b = None
        if betas is not None:
            b = self._tuple_angles(betas, n, mem, "beta")
        elif int(beta_kind) == GENERIC:
            raise AssertionError("beta is required for generic co-latitudes")
        if mem == MEM_DEVICE:
            import torch
            if out is None:
                out = torch.empty(n, dtype=torch.complex128, device=tj.device)
            self._use_torch_stream(tj.device)
        elif out is None:
            out = np.empty(n, dtype=np.complex128)
        out = self._check_out(out, mem, (n,), True)
        _check(self._lib.wd_Djmn_batch(self._h, _ptr(tj), _ptr(tm), _ptr(tn), _ptr(a), _ptr(b), _ptr(g), n,
                                       int(beta_kind), _ptr(out), mem))
        return out

    # -- the reference's scalar API ---------------------------------------------------------
    def wignerd_(self, d, j, beta, Jy=None):
        """``wignerd!(d, j, beta, [Jy])`` (:306-310): fills ``d`` in place and returns it."""
        twoj = twoj_of(j)
        N = twoj + 1
        self._check_Jy(twoj, Jy)
        if not isinstance(d, np.ndarray) or d.dtype != np.float64:
            raise AssertionError("d must be a float64 numpy array")
        if d.shape != (N, N):
            raise AssertionError(f"size of dj must be {N}x{N}")        # :267
        buf = d if d.flags.f_contiguous else np.asfortranarray(d)
        _check(self._lib.wd_d_matrix(self._h, twoj, float(beta), _beta_kind(beta), _ptr(buf)))
        self._prepared.add(twoj)
        if buf is not d:
            d[...] = buf
        return d

    def wignerd(self, j, beta, Jy=None):
        """``wignerd(j, beta, [Jy])`` (:320-324)."""
        twoj = twoj_of(j)
        d = np.zeros((twoj + 1, twoj + 1), dtype=np.float64, order="F")      # zeros(_matrixsize(j)) :321
        return self.wignerd_(d, j, beta, Jy)

    def wignerD_(self, D, j, alpha, beta, gamma, Jy=None):
        """``wignerD!(D, j, alpha, beta, gamma, [Jy])`` (:335-342)."""
        twoj = twoj_of(j)
        N = twoj + 1
        self._check_Jy(twoj, Jy)
        if not isinstance(D, np.ndarray) or D.dtype != np.complex128:
            raise AssertionError("D must be a complex128 numpy array")
        if D.shape != (N, N):
            raise AssertionError(f"size of dj must be {N}x{N}")
        buf = D if D.flags.f_contiguous else np.asfortranarray(D)
        _check(self._lib.wd_D_matrix(self._h, twoj, float(alpha), float(beta), _beta_kind(beta), float(gamma), _ptr(buf)))
        self._prepared.add(twoj)
        if buf is not D:
            D[...] = buf
        return D

    def wignerD(self, j, alpha, beta, gamma, Jy=None):
        """``wignerD(j, alpha, beta, gamma, [Jy])`` (:353-357)."""
        twoj = twoj_of(j)
        D = np.zeros((twoj + 1, twoj + 1), dtype=np.complex128, order="F")
        return self.wignerD_(D, j, alpha, beta, gamma, Jy)

    def wignerdjmn(self, j, m, n, beta, Jy=None) -> float:
        """``WignerD.wignerdjmn(j, m, n, beta, [Jy])`` (:214-237)."""
        twoj = twoj_of(j)
        twom, twon = _two_of(m, "m"), _two_of(n, "n")
        kind = _beta_kind(beta)
        if kind in (GENERIC, EQUATOR):
            self._check_Jy(twoj, Jy)
        out = np.empty(1)
        tj, tm, tn = (np.array([v], np.int32) for v in (twoj, twom, twon))     # keep alive across the call
        b = np.array([float(beta)])
        _check(self._lib.wd_djmn_batch(self._h, _ptr(tj), _ptr(tm), _ptr(tn), _ptr(b), 1, kind, _ptr(out), MEM_HOST))
        return float(out[0])

    def wignerDjmn(self, j, m, n, alpha, beta, gamma, Jy=None) -> complex:
        """``WignerD.wignerDjmn(j, m, n, alpha, beta, gamma, [Jy])`` (:245)."""
        twoj = twoj_of(j)
        twom, twon = _two_of(m, "m"), _two_of(n, "n")
        kind = _beta_kind(beta)
        if kind in (GENERIC, EQUATOR):
            self._check_Jy(twoj, Jy)
        out = np.empty(1, dtype=np.complex128)
        tj, tm, tn = (np.array([v], np.int32) for v in (twoj, twom, twon))     # keep alive across the call
        a, b, g = (np.array([float(v)]) for v in (alpha, beta, gamma))
        _check(self._lib.wd_Djmn_batch(self._h, _ptr(tj), _ptr(tm), _ptr(tn), _ptr(a), _ptr(b), _ptr(g), 1, kind,
                                       _ptr(out), MEM_HOST))
        return complex(out[0])


_tls = threading.local()


def default_engine() -> Engine:
    eng = getattr(_tls, "engine", None)
    if eng is None:
        eng = Engine(int(os.environ.get("WIGNERD_B200_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
        _tls.engine = eng
    return eng


# module-level functions with the reference's names
def wignerd(j, beta, Jy=None):
    return default_engine().wignerd(j, beta, Jy)


def wignerd_(d, j, beta, Jy=None):
    return default_engine().wignerd_(d, j, beta, Jy)


def wignerD(j, alpha, beta, gamma, Jy=None):
    return default_engine().wignerD(j, alpha, beta, gamma, Jy)


def wignerD_(D, j, alpha, beta, gamma, Jy=None):
    return default_engine().wignerD_(D, j, alpha, beta, gamma, Jy)


def wignerdjmn(j, m, n, beta, Jy=None):
    return default_engine().wignerdjmn(j, m, n, beta, Jy)


def wignerDjmn(j, m, n, alpha, beta, gamma, Jy=None):
    return default_engine().wignerDjmn(j, m, n, alpha, beta, gamma, Jy)
