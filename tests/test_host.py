"""Host logic and the C-ABI surface, without a GPU: the library loads, exports every symbol the header
declares, maps errors, and the product refuses to compute without CUDA (no CPU fallback)."""
import ctypes
import os
import re
from fractions import Fraction as F
from math import pi

import numpy as np
import pytest

import wignerd.jl_b200 as w

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_builds_and_loads():
    from wignerd.jl_b200 import build
    path = build.build()
    assert os.path.exists(path)
    L = w.load_library()
    assert L.wd_abi_version() == 1


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wignerd_b200.h")).read()
    declared = set(re.findall(r"\b(wd_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert declared == set(w.ABI_SYMBOLS), declared ^ set(w.ABI_SYMBOLS)
    L = ctypes.CDLL(w.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), sym


def test_resident_kernel_range():
    """shared-memory budget of the DMMA kernels (pure host arithmetic of the library): the benchmark's j = 100 and
    j = 49.5 must stay in the table-resident kernels; bench.py names the kernel of a config by the same limit"""
    L = w.load_library()
    limit = 232448                                      # sm_100a opt-in maximum per CTA (227 KB)
    fits = lambda twoj, k: 0 < L.wd_k2_shared_bytes(twoj, k) <= limit     # noqa: E731
    assert fits(200, 0) and fits(99, 1)
    assert max(t for t in range(0, 400) if fits(t, 0)) == 223
    assert max(t for t in range(0, 400) if fits(t, 1)) == 223
    assert fits(1024, 2) and fits(1023, 2)
    assert L.wd_k2_shared_bytes(-1, 0) == -1 and L.wd_k2_shared_bytes(10, 7) == -1


def test_package_path_literal():
    # the package is reachable as ./wignerd.jl_b200/ too (same directory)
    assert os.path.samefile(os.path.join(ROOT, "wignerd.jl_b200"), os.path.join(ROOT, "wignerd", "jl_b200"))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wignerd")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "libwignerd_ref" not in src, f


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(w.CudaError):
        w.Engine(0)
    with pytest.raises(w.CudaError):
        w.wignerd(1, 0.3)


def test_null_and_bad_arguments_return_status_codes():
    L = w.load_library()
    assert L.wd_create(0, None) == 2                       # WD_ERR_ASSERT, never a crash
    assert b"null" in L.wd_last_error()
    assert L.wd_d_batch(None, 2, None, 1, None, 0) == 2
    assert L.wd_prepare(None, None, 0) == 2
    assert L.wd_launch_count(None) == 0
    assert L.wd_destroy(None) == 0


def test_twoj_of():
    assert w.twoj_of(0) == 0 and w.twoj_of(1) == 2 and w.twoj_of(0.5) == 1 and w.twoj_of(F(7, 2)) == 7
    assert w.twoj_of(1.5) == 3 and w.twoj_of(100) == 200
    with pytest.raises(w.InexactError):
        w.twoj_of(0.3)
    with pytest.raises(AssertionError):
        w.twoj_of(-1)


def test_special_colatitudes():
    # test/runtests.jl:23-43
    assert float(w.Equator()) == pi / 2 and np.cos(w.NorthPole()) == 1 and np.sin(w.NorthPole()) == 0
    assert float(w.SouthPole()) == pi and float(w.NorthPole()) == 0.0
    assert w.Equator().kind == w.EQUATOR and w.NorthPole().kind == w.NORTH and w.SouthPole().kind == w.SOUTH
    assert isinstance(w.Equator() * 2, float)              # promote_rule -> Float64
    for a in range(-10, 11):                               # test/runtests.jl:17-22
        for x in (a, F(a, 2)):
            s, c = w._sincos(x, w.Equator())
            assert abs(s - np.sin(float(x) * pi / 2)) < 1e-14 and abs(c - np.cos(float(x) * pi / 2)) < 1e-14


def test_wigner_matrix_wrapper():
    # test/runtests.jl:361-365
    d = np.arange(4.0).reshape(2, 2)
    wm = w._offsetmatrix(0.5, d)
    assert wm.shape == (2, 2)
    assert wm[-0.5, -0.5] == 0.0 and wm[F(1, 2), -0.5] == 2.0 and wm[0.5, 0.5] == 3.0
    wm[0.5, -0.5] = 7.0
    assert d[1, 0] == 7.0
    with pytest.raises(AssertionError):
        w._offsetmatrix(1, d)
    with pytest.raises(IndexError):
        wm[1, 0]


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (100000, 8), (100000, 3), (4096, 8), (17, 2), (10**7, 8)])
def test_shard_bounds_partition(n, world):
    edges = [w.shard_bounds(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
        assert hi == lo2 and lo <= hi
    sizes = [hi - lo for lo, hi in edges]
    assert max(sizes) - min(sizes) <= 16 or n < 16 * world
    for lo, hi in edges[:-1]:
        assert lo % 16 == 0 or lo == n


def test_shard_arithmetic_native_equals_python():
    """wd_shard_bounds (what a non-Python host and the multi-GPU handle use) == shard.shard_bounds; slabs partition [0, n)"""
    rng = np.random.default_rng(5)
    cases = [(0, 1, 16), (1, 8, 16), (15, 2, 16), (16, 2, 16), (17, 3, 16), (100000, 8, 16), (4096, 8, 16), (10_000_000, 7, 1)]
    cases += [(int(rng.integers(0, 500000)), int(rng.integers(1, 17)), int(rng.choice([1, 8, 16]))) for _ in range(40)]
    for n, world, align in cases:
        prev = 0
        for r in range(world):
            lo, hi = w.shard_bounds_native(n, r, world, align)
            assert (lo, hi) == w.shard_bounds(n, r, world, align)
            assert lo == prev and lo <= hi and (lo % align == 0 or lo == n)
            prev = hi
        assert prev == n
    L = w.load_library()
    assert L.wd_shard_bounds(10, 3, 2, 16, None, None) == 2
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    assert L.wd_shard_bounds(10, 3, 2, 16, ctypes.byref(lo), ctypes.byref(hi)) == 2      # rank >= world


def test_symmetric_shard_is_closed_under_mirroring():
    """every rank's angle set is mapped onto itself by b -> n-1-b (the pairing of the contraction kernel needs it) and the
    ranks partition [0, n)"""
    for n in (100000, 100001, 4096, 800000, 37, 16, 1):
        for world in (1, 2, 4, 8):
            seen = np.zeros(n, dtype=int)
            for r in range(world):
                (lo, hi), (blo, bhi) = w.symmetric_shard(n, r, world)
                idx = np.concatenate([np.arange(lo, hi), np.arange(blo, bhi)])
                seen[idx] += 1
                assert np.array_equal(np.sort(n - 1 - idx), np.sort(idx))
                # the rank's own list (front slab then back slab) is mirror-symmetric position by position
                assert np.array_equal(idx + idx[::-1], np.full(len(idx), n - 1))
            assert (seen == 1).all()


def test_column_kernel_range_and_new_entry_points_reject_nulls():
    L = w.load_library()
    limit = 232448
    assert 0 < L.wd_k2_shared_bytes(200, 3) <= limit          # the benchmark's j = 100 runs in the column-staged kernel
    assert 0 < L.wd_k2_shared_bytes(99, 3) <= limit
    assert L.wd_k2_shared_bytes(4, 3) > limit                 # N < 6: plain kernel
    assert L.wd_k2_shared_bytes(301, 3) > limit
    # tensor-store complex-D kernel: cfg-3 (j = 49.5) inside, range 5 <= 2j <= 127 without gaps
    fits4 = [t for t in range(0, 400) if 0 < L.wd_k2_shared_bytes(t, 4) <= limit]
    assert fits4 == list(range(5, 128)), (fits4[:3], fits4[-3:])
    assert L.wd_create_multi(0, None, None) == 2
    h = ctypes.c_void_p()
    assert L.wd_create_multi(0, None, ctypes.byref(h)) == 2   # ndev must be >= 1
    assert L.wd_multi_size(None) == 0
    assert L.wd_tables_save(None, b"/tmp/x") == 2 and L.wd_tables_load(None, None) == 2
    assert L.wd_gather(None, None, None, 0, None) == 2
    assert L.wd_set_slab_bytes(None, 1 << 20) == 2


def test_c_program_links_against_every_entry_point(tmp_path):
    """a C (not C++, not Python) consumer: the header compiles as C and every declared symbol links and is callable"""
    import subprocess
    exe = str(tmp_path / "c_abi_link")
    libdir = os.path.dirname(w.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_link.c"), "-o", exe, "-L", libdir, "-lwignerd_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "c abi ok: 29 entry points" in out.stdout
