#!/usr/bin/env python
"""bench.py -- benchmark of the B200 Wigner-d/D engine.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (C port of its hot loop)

Headline (BASELINE.json's metric): Wigner-d matrix elements per second, FP64, at j = 100 over 10^5 angles per GPU
(cfg-2 of SURVEY.md section 8(d)); weak scaling: every rank owns 10^5 angles of an N*10^5-point linspace(0, pi) grid
(a slab of the first half of the grid plus its mirror image, so that a rank's own grid is symmetric about pi/2), no
data-path collective.  One step = one pass of the hot path (wd_d_batch) over the rank's angles.

The same line carries every other configuration of BASELINE.json under "configs" (cfg1, cfg2 on a generic random
grid, cfg3, cfg4 on its 4096/N split, cfg5), each with roofline / cpu_baseline / e2e; the eigensolver set-up times,
a cold end-to-end call, strong scaling of the fixed 10^5-angle workload and the host-copy ceiling of the box.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg1": dict(twoj=20, nb=1000, kind="d", desc="wignerd(10, beta), 1000 beta in [0, pi]"),
    "cfg2": dict(twoj=200, nb=100000, kind="d", desc="wignerd(100, beta), 10^5 beta in [0, pi] per GPU"),
    "cfg3": dict(twoj=99, nb=100000, kind="D", desc="wignerD(49.5, alpha, beta, gamma), 10^5 Euler triples per GPU"),
    "cfg4": dict(twoj=1024, nb=4096, kind="multi", desc="all d^j(beta), j=0..512, on the 4096-point beta grid sharded over the GPUs (4096/N angles per GPU), output ring-buffered"),
    "cfg5": dict(twoj=400, nb=10_000_000, kind="jmn", desc="wignerdjmn, 10^7 random (j<=200, m, n, beta) tuples per GPU"),
}
METRIC = {"d": ("wigner_d_elements_per_sec_fp64", "elements/s"), "D": ("wigner_d_elements_per_sec_fp64", "elements/s"),
          "multi": ("wigner_d_elements_per_sec_fp64", "elements/s"), "jmn": ("wignerdjmn_tuples_per_sec_fp64", "tuples/s")}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ======================================================================================================
# CPU side: the C port of the reference's hot loop (oracle/wignerd_ref.c) -- the only use of oracle/ in this file
# ======================================================================================================
def cpu_matrix_rate(twoj: int, nang: int, nthreads: int = 0, kind="d"):
    """elements/s of the reference loop on the host cores; the eigenvectors are cached as the reference does (:163-176)"""
    from oracle import cref
    from oracle import wignerd_oracle as o
    o.jy_eigen(twoj)
    N = twoj + 1
    rng = np.random.default_rng(1)
    betas = np.linspace(0, np.pi, nang)
    threads = nthreads or cref.num_threads()
    if kind == "D":
        a, g = rng.uniform(0, 2 * np.pi, nang), rng.uniform(0, 2 * np.pi, nang)
        cref.wignerD_batch(twoj, a[:threads], betas[:threads], g[:threads], threads)
        t0 = time.perf_counter()
        cref.wignerD_batch(twoj, a, betas, g, threads)
    else:
        cref.wignerd_batch(twoj, betas[:threads], threads)
        t0 = time.perf_counter()
        cref.wignerd_batch(twoj, betas, threads)
    dt = time.perf_counter() - t0
    return nang * N * N / dt, threads, dt


def cpu_baseline_for(name: str, cfg: dict):
    """bounded sample of the config's workload on all host threads"""
    from oracle import cref
    threads = cref.num_threads()
    unit = METRIC[cfg["kind"]][1]
    try:
        if cfg["kind"] in ("d", "D"):
            twoj = cfg["twoj"]
            nang = {"cfg1": 1000, "cfg3": 1024}.get(name, 512)
            rate, threads, dt = cpu_matrix_rate(twoj, min(nang, cfg["nb"]), 0, cfg["kind"])
            sample = f"{min(nang, cfg['nb'])} of {cfg['nb']} angles, {dt:.2f} s wall on {threads} host threads"
        elif cfg["kind"] == "multi":
            # cost of the reference per angle and j is c * (j+1)^2 (2j+1) sincos terms: c is fitted on a few j, the whole
            # j = 0..512 workload extrapolated (the reference's cost is linear in the number of terms)
            js, nang = [32, 96, 192, 320], threads
            terms = el = 0.0
            t_all = 0.0
            for j in js:
                r, _, dt = cpu_matrix_rate(2 * j, nang, 0, "d")
                t_all += dt
                terms += nang * (j + 1) ** 2 * (2 * j + 1)
            c = t_all / terms
            jmax = cfg["twoj"] // 2
            tot_terms = sum((j + 1) ** 2 * (2 * j + 1) for j in range(jmax + 1))
            tot_el = sum((2 * j + 1) ** 2 for j in range(jmax + 1))
            rate = tot_el / (c * tot_terms)
            sample = (f"j in {js} x {nang} angles, {t_all:.2f} s wall on {threads} host threads; seconds per sincos term fitted "
                      f"({c * threads * 1e9:.1f} ns per term and core) and applied to all j <= {jmax}")
        else:
            rng = np.random.default_rng(3)
            n = 40000
            tj = rng.integers(0, cfg["twoj"] + 1, n).astype(np.int32)
            tm = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
            tn = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
            b = rng.uniform(0, np.pi, n)
            cref.wignerdjmn_batch(tj[:2000], tm[:2000], tn[:2000], b[:2000], threads)
            t0 = time.perf_counter()
            cref.wignerdjmn_batch(tj, tm, tn, b, threads)
            dt = time.perf_counter() - t0
            rate = n / dt
            sample = f"{n} of {cfg['nb']} tuples, {dt:.2f} s wall on {threads} host threads"
        return {"value": rate, "unit": unit, "cores": threads, "kind": "port",
                "sample": sample + " (C port of the reference loop, oracle/wignerd_ref.c; Julia is not in the image)"}
    except Exception as e:      # noqa: BLE001
        return {"value": None, "unit": unit, "cores": 0, "kind": "port", "sample": repr(e)}


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU algorithm (C port; Julia is not in the image), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref
    threads = cref.num_threads()
    twoj = cfg["twoj"]
    N = twoj + 1
    metric, unit = METRIC[cfg["kind"]]
    if cfg["kind"] == "jmn":
        rng = np.random.default_rng(3)
        n = 20000
        tj = rng.integers(0, twoj + 1, n).astype(np.int32)
        tm = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
        tn = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
        b = rng.uniform(0, np.pi, n)
        fn = lambda: cref.wignerdjmn_batch(tj, tm, tn, b, threads)      # noqa: E731
        units = n
        sample = f"{n} of {cfg['nb']} tuples per step"
    else:
        if cfg["kind"] == "multi":
            twoj = 512                                     # a mid-size j of the list stands for the whole list in this arm
            N = twoj + 1
        from oracle import wignerd_oracle as o
        o.jy_eigen(twoj)                                   # cached once per j by the reference (:163-176): not timed
        nang = max(threads * 2, 16)                        # bounded sample per step (~80 ms of wall time at j=100)
        betas = np.linspace(0, np.pi, nang)
        if cfg["kind"] == "D":
            rng = np.random.default_rng(2)
            a, g = rng.uniform(0, 2 * np.pi, nang), rng.uniform(0, 2 * np.pi, nang)
            fn = lambda: cref.wignerD_batch(twoj, a, betas, g, threads)      # noqa: E731
        else:
            fn = lambda: cref.wignerd_batch(twoj, betas, threads)            # noqa: E731
        units = nang * N * N
        sample = (f"{nang} of {cfg['nb']} angles per step, scaled linearly to the full batch: an EXTRAPOLATION (the "
                  "reference's cost is linear in the number of angles once the eigenvectors are cached)")
    for _ in range(max(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    value = units / dt
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "name": args.config, "twoj": twoj, "sample": sample},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C port of src/WignerD.jl:187-206,:266-295 (oracle/wignerd_ref.c); Julia is not installed in this image",
    }
    print(json.dumps(line), flush=True)


# ======================================================================================================
# GPU side
# ======================================================================================================
class Ctx:
    pass


def gpu_time(ctx, fn, steps, warmup):
    """CUDA events on torch's current stream (the stream the library launches on); max over ranks; returns ms per step"""
    torch = ctx.torch
    for _ in range(warmup):
        fn()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1) / steps
    return ctx.allmax(ms), ms


def wall_time(ctx, fn, steps, warmup):
    """host wall clock around synchronous (host-buffer) API calls; max over ranks; returns seconds per step"""
    for _ in range(warmup):
        fn()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    ctx.torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return ctx.allmax(dt), dt


def f_alg(twoj: int, nb: int) -> float:
    """2 W K nb: the minimal multiply-add count of the Feng contraction"""
    N = twoj + 1
    K = (N + 1) // 2
    W = ((twoj // 2 + 1) ** 2) if twoj % 2 == 0 else ((twoj + 1) // 2) * ((twoj + 3) // 2)
    return 2.0 * W * K * nb


def symmetric_betas(ctx, nb_rank: int):
    """this rank's nb_rank angles of linspace(0, pi, world * nb_rank): a slab of the first half of the grid followed by its
    mirror image (wignerd.jl_b200.symmetric_shard), so that the rank's own grid is symmetric about pi/2"""
    import wignerd.jl_b200 as w
    torch = ctx.torch
    total = nb_rank * ctx.world
    (lo, hi), (blo, bhi) = w.symmetric_shard(total, ctx.rank, ctx.world)
    idx = torch.cat([torch.arange(lo, hi, device=ctx.dev), torch.arange(blo, bhi, device=ctx.dev)]).to(torch.float64)
    return (idx * (np.pi / max(total - 1, 1))).contiguous()


def pinned_empty(ctx, shape, dtype):
    torch = ctx.torch
    try:
        import psutil
        need = int(np.prod(shape)) * (16 if dtype == torch.complex128 else 8) * int(os.environ.get("LOCAL_WORLD_SIZE", ctx.world))
        if psutil.virtual_memory().available < 1.5 * need:
            return None
    except Exception:
        pass
    try:
        return torch.empty(shape, dtype=dtype, pin_memory=True)
    except RuntimeError:
        torch.cuda.empty_cache()
        return None


def release_pinned(ctx):
    """give pinned host blocks back to the OS (torch's caching host allocator keeps them): 8 ranks x (32 + 16 + ...) GB otherwise"""
    try:
        ctx.torch.cuda.synchronize()
        ctx.torch._C._host_emptyCache()
    except Exception:
        pass


def fp64_peaks(ctx):
    """FP64 denominators measured in THIS run (MEASURED_PEAKS.json has none): DMMA / DFMA register loops, cuBLAS DGEMM"""
    torch, eng = ctx.torch, ctx.eng
    out = {}
    try:
        out["dmma_loop_tflops"] = eng.fp64_peak(0, 20000)
        out["dfma_loop_tflops"] = eng.fp64_peak(1, 20000)
        out["sin_evals_per_s"] = eng.fp64_peak(2, 2000) * 1e12
        out["rotation_steps_per_s"] = eng.fp64_peak(3, 20000) * 1e12
        a = torch.randn(6144, 6144, dtype=torch.float64, device=ctx.dev)
        torch.matmul(a, a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            torch.matmul(a, a)
        e1.record()
        torch.cuda.synchronize()
        out["cublas_dgemm_tflops"] = 3 * 2 * 6144 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a
    except Exception as e:      # noqa: BLE001
        out["error"] = repr(e)
    out["peak_tflops"] = max(out.get("cublas_dgemm_tflops", 0.0), out.get("dmma_loop_tflops", 0.0)) or None
    out["peak_source"] = ("max(cuBLAS DGEMM 6144^3 via torch.matmul float64, DMMA.8x8x4 register loop) measured in this run; "
                          "MEASURED_PEAKS.json carries no FP64 figure")
    return out


def rooflines(ctx, twoj, nb, ms_rank, cplx, kernel, traffic):
    """both roofs of a matrix config for THIS rank's kernel; the binding one (the slower of the two) is 'roofline'"""
    N = twoj + 1
    peak, peak_src = ctx.hbm_peak
    bytes_alg = nb * N * N * (16 if cplx else 8)
    t = ms_rank * 1e-3
    hbm = {"bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": peak, "unit": "GB/s", "frac": bytes_alg / t / 1e9 / peak,
           "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu capture of a shorter launch, scaled)" if traffic else None,
           "peak_source": peak_src, "kernel": kernel, "algorithmic_bytes_per_launch": bytes_alg}
    res = {"roofline_hbm": hbm, "roofline": hbm}
    p64 = ctx.fp64.get("peak_tflops")
    if p64:
        fa = f_alg(twoj, nb)
        ten = {"bound": "tensor", "achieved": fa / t / 1e12, "peak": p64, "unit": "TFLOP/s", "frac": fa / t / 1e12 / p64,
               "traffic": traffic, "kernel": kernel, "algorithmic_flops_per_launch": fa, "peak_source": ctx.fp64["peak_source"]}
        res["roofline_tensor"] = ten
        if fa / (p64 * 1e12) > bytes_alg / (peak * 1e9):
            res["roofline"] = ten
    return res


def traffic_of(name):
    """DRAM bytes (read + write) of the config's dominant kernel per launch, FROM profiles/traffic.json: an ncu --set full
    capture of a shorter launch of the same kernel scaled to the benched batch (see its _note), not a live measurement"""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tp)).get(name)
        return t
    except Exception:
        return None


def bench_matrices(ctx, name, cfg, steps, warmup, grid="symmetric", e2e=True, pageable=False, clocks=False):
    """cfg-1/2/3: one j, nb angles per rank; device-resident timing + the same call with HOST buffers"""
    torch, eng = ctx.torch, ctx.eng
    twoj, nb = cfg["twoj"], cfg["nb"]
    N = twoj + 1
    cplx = cfg["kind"] == "D"
    jv = twoj // 2 if twoj % 2 == 0 else twoj / 2
    if cplx:
        gen = torch.Generator(device=ctx.dev).manual_seed(2 + ctx.rank)
        alpha = torch.rand(nb, dtype=torch.float64, device=ctx.dev, generator=gen) * (2 * np.pi)
        gamma = torch.rand(nb, dtype=torch.float64, device=ctx.dev, generator=gen) * (2 * np.pi)
        beta = torch.rand(nb, dtype=torch.float64, device=ctx.dev, generator=gen) * np.pi
        out = torch.empty((nb, N, N), dtype=torch.complex128, device=ctx.dev)
        step = lambda: eng.wignerD_batch(jv, alpha, beta, gamma, out=out)      # noqa: E731
        inputs = "alpha, gamma ~ U[0, 2pi), beta ~ U[0, pi], torch.Generator(cuda).manual_seed(2 + rank)"
    else:
        if grid == "symmetric":
            beta = symmetric_betas(ctx, nb)
            nb_nominal, nb = nb, beta.numel()             # the 16-aligned symmetric shards differ by up to 32 angles between ranks
            inputs = "beta = the rank's symmetric shard (slab of the first half + its mirror image) of linspace(0, pi, n_gpus * angles_per_gpu)"
        else:
            gen = torch.Generator(device=ctx.dev).manual_seed(1 + ctx.rank)
            beta = torch.rand(nb, dtype=torch.float64, device=ctx.dev, generator=gen) * np.pi
            inputs = "beta ~ U[0, pi), torch.Generator(cuda).manual_seed(1 + rank) (the generic-grid variant of SURVEY 8(d))"
        out = torch.empty((nb, N, N), dtype=torch.float64, device=ctx.dev)
        step = lambda: eng.wignerd_batch(jv, beta, out=out)                     # noqa: E731
    elems = nb * N * N                                    # this rank's
    elems_all = (nb_nominal if (not cplx and grid == "symmetric") else nb) * N * N * ctx.world
    bytes_out = elems * (16 if cplx else 8)
    sampler = ClockSampler(ctx.local) if (clocks and ctx.rank == 0) else None
    for _ in range(warmup):
        step()
    ctx.barrier()
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    ms_max, ms_rank = gpu_time(ctx, step, steps, 0)
    launches = eng.launch_count() - l0
    clk = sampler.stop() if sampler else None
    if cplx:
        variant = getattr(ctx, "variant", 0)
        kernel = ("k3_tma_kernel" if (5 <= twoj <= 127 and variant in (0, 7)) else
                  "k2_cplx_kernel" if twoj <= 223 else "k2_stream_kernel + phase_kernel")
    elif twoj > 223:
        kernel = "k2_stream_kernel"
    else:
        kernel = "k2_col_kernel (paired)" if (grid == "symmetric" and 5 <= twoj <= 200 and nb >= 32) else "k2_real_kernel"
    rec = {"ms_per_step": ms_max, "value": elems_all / (ms_max * 1e-3), "unit": "elements/s", "steps": steps,
           "gpu_launches": int(launches), "angles_per_gpu": nb, "twoj": twoj, "inputs": inputs,
           "l2": ("outputs (>= 16 GB per step) far exceed the 126 MB L2; no flush needed" if bytes_out > 4 * 126e6 else
                  "L2-RESIDENT timing: the step's output fits in the 126 MB L2 and is rewritten every step without a flush "
                  "(CPU-runnable sanity config, launch/latency bound; not a bandwidth claim)")}
    rec.update(rooflines(ctx, twoj, nb, ms_rank, cplx, kernel, traffic_of(name if (cplx or grid == "symmetric") else name + "_random_grid")))
    if clk is not None:
        rec["clocks"] = clk
    del out
    torch.cuda.empty_cache()
    if e2e:
        ne = nb
        dt_ = torch.complex128 if cplx else torch.float64
        h_out = pinned_empty(ctx, (ne, N, N), dt_)
        if h_out is None:
            ne = min(nb, 16384)
            h_out = torch.empty((ne, N, N), dtype=dt_, pin_memory=True)
        hb = beta[:ne].cpu().pin_memory()
        if cplx:
            ha, hg = alpha[:ne].cpu().pin_memory(), gamma[:ne].cpu().pin_memory()
            estep = lambda o_: eng.wignerD_batch(jv, ha.numpy(), hb.numpy(), hg.numpy(), out=o_)      # noqa: E731
        else:
            estep = lambda o_: eng.wignerd_batch(jv, hb.numpy(), out=o_)                                # noqa: E731
        esteps = 3 if bytes_out > 1e9 else 20
        dt_max, dt_rank = wall_time(ctx, lambda: estep(h_out.numpy()), esteps, 2)
        d2h = ne * N * N * (16 if cplx else 8)
        rec["e2e"] = {"value": ne * N * N * ctx.world / dt_max, "unit": "elements/s", "h2d_bytes_per_step": ne * 8 * (3 if cplx else 1),
                      "d2h_bytes_per_step": d2h, "steps": esteps, "angles_per_gpu": ne,
                      "d2h_gbs_this_rank": d2h / dt_rank / 1e9,
                      "kernel_share_of_e2e": (ms_rank * 1e-3 * ne / nb) / dt_rank,
                      "note": "wd_d_batch / wd_D_batch with pinned HOST buffers; PCIe copies inside the timed region; the kernels of "
                              "slab k+1 run under the copy of slab k, so the call is bound by the device-to-host copy"}
        if pageable:
            # the same call into PAGEABLE host memory (what a Julia Array is): pinned bounce ring + host copy threads
            npg = min(ne, 16384)
            pg = np.empty((npg, N, N), dtype=np.complex128 if cplx else np.float64)
            pg[...] = 0                                   # touch the pages: first-touch faults are the OS's cost, not the library's
            if cplx:
                pstep = lambda: eng.wignerD_batch(jv, ha.numpy()[:npg], hb.numpy()[:npg], hg.numpy()[:npg], out=pg)      # noqa: E731
            else:
                pstep = lambda: eng.wignerd_batch(jv, hb.numpy()[:npg], out=pg)                                           # noqa: E731
            dtp, _ = wall_time(ctx, pstep, 3, 1)
            rec["e2e_pageable"] = {"value": npg * N * N * ctx.world / dtp, "unit": "elements/s", "angles_per_gpu": npg,
                                   "d2h_bytes_per_step": npg * N * N * (16 if cplx else 8),
                                   "note": "destination = pageable numpy memory (pages touched beforehand)"}
            del pg
        del h_out
        release_pinned(ctx)
    return rec


def bench_multi(ctx, name, cfg, steps, warmup, e2e=True):
    """cfg-4: every integer j <= 512 on this rank's symmetric shard of the 4096-point grid (4096 / N angles); the output
    (0.74 TB per 512 angles) is consumed in place: each j overwrites one reusable buffer -- the ring buffer of SURVEY 7.7."""
    torch, eng = ctx.torch, ctx.eng
    jmax = cfg["twoj"] // 2
    nb_nominal = max(16, cfg["nb"] // ctx.world)
    beta = symmetric_betas(ctx, nb_nominal)
    nb = beta.numel()
    js = list(range(0, jmax + 1))
    buf = torch.empty(nb * (2 * jmax + 1) ** 2, dtype=torch.float64, device=ctx.dev)
    offsets = np.zeros(len(js), dtype=np.int64)
    elems = nb * sum((2 * j + 1) ** 2 for j in js)
    t0 = time.perf_counter()
    eng.prepare([2 * j for j in js])
    ctx.torch.cuda.synchronize()
    prep_s = time.perf_counter() - t0
    step = lambda: eng.wignerd_multi(js, beta, out=buf, offsets=offsets)      # noqa: E731
    for _ in range(warmup):
        step()
    ctx.barrier()
    l0 = eng.launch_count()
    ms_max, ms_rank = gpu_time(ctx, step, steps, 0)
    launches = eng.launch_count() - l0
    fa = 2.0 * nb * sum((j + 1) ** 3 for j in js)
    t = ms_rank * 1e-3
    peak, peak_src = ctx.hbm_peak
    elems_all = nb_nominal * ctx.world * sum((2 * j + 1) ** 2 for j in js)
    rec = {"ms_per_step": ms_max, "value": elems_all / (ms_max * 1e-3), "unit": "elements/s", "steps": steps,
           "gpu_launches": int(launches), "angles_per_gpu": nb, "j_list": f"0..{jmax}", "prepare_all_j_s": prep_s,
           "inputs": "beta = the rank's symmetric shard of linspace(0, pi, 4096)",
           "l2": "every j writes nb*(2j+1)^2*8 bytes into one reused buffer (4.3 GB at j=512, 512 angles): outputs exceed the L2 "
                 "for j >= 90, which is where ~all of the time goes"}
    p64 = ctx.fp64.get("peak_tflops")
    if p64:
        rec["roofline"] = {"bound": "tensor", "achieved": fa / t / 1e12, "peak": p64, "unit": "TFLOP/s", "frac": fa / t / 1e12 / p64,
                           "traffic": None, "kernel": "k2_stream_kernel (2j > 223), k2_real_kernel / k2_col_kernel below",
                           "algorithmic_flops_per_launch": fa, "peak_source": ctx.fp64["peak_source"]}
    rec["roofline_hbm"] = {"bound": "hbm", "achieved": elems * 8 / t / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": elems * 8 / t / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": elems * 8}
    del buf
    torch.cuda.empty_cache()
    if e2e:
        # host buffers: all j for a 32-angle slice, every j copied to one reused pinned host buffer (the d matrices are
        # consumed on the host as they arrive: 46 GB cross PCIe per step)
        ne = 32
        hbuf = torch.empty(ne * (2 * jmax + 1) ** 2, dtype=torch.float64, pin_memory=True)
        hb = beta[:ne].cpu().numpy()
        dt_max, dt_rank = wall_time(ctx, lambda: eng.wignerd_multi(js, hb, out=hbuf.numpy(), offsets=offsets), 1, 1)
        el = ne * sum((2 * j + 1) ** 2 for j in js)
        rec["e2e"] = {"value": el * ctx.world / dt_max, "unit": "elements/s", "h2d_bytes_per_step": ne * 8 * len(js),
                      "d2h_bytes_per_step": el * 8, "steps": 1, "angles_per_gpu": ne, "d2h_gbs_this_rank": el * 8 / dt_rank / 1e9,
                      "note": "wd_d_multi with HOST buffers on a 32-angle slice of the rank's grid, every j delivered to a reused pinned "
                              "buffer: PCIe-bound"}
        del hbuf
        release_pinned(ctx)
    return rec


def bench_rotate(ctx, steps, warmup):
    """the consumer of cfg-4: rotate one set of spherical-harmonic coefficients to lmax = 512 by the rank's
    4096 / N rotations -- wd_rotate_sh works from the eigenvectors (two tensor-core GEMMs per l), no d matrix is formed"""
    torch, eng = ctx.torch, ctx.eng
    lmax = 512
    nb_nominal = max(16, 4096 // ctx.world)
    L2 = (lmax + 1) ** 2
    gen = torch.Generator(device=ctx.dev).manual_seed(4 + ctx.rank)
    alm = torch.complex(torch.randn(L2, dtype=torch.float64, device=ctx.dev, generator=gen), torch.randn(L2, dtype=torch.float64, device=ctx.dev, generator=gen))
    beta = symmetric_betas(ctx, nb_nominal)
    nb = beta.numel()
    alpha = torch.rand(nb, dtype=torch.float64, device=ctx.dev, generator=gen) * (2 * np.pi)
    gamma = torch.rand(nb, dtype=torch.float64, device=ctx.dev, generator=gen) * (2 * np.pi)
    out = torch.empty((nb, L2), dtype=torch.complex128, device=ctx.dev)
    step = lambda: eng.rotate_sh(lmax, alm, alpha, beta, gamma, out=out)      # noqa: E731
    for _ in range(warmup):
        step()
    ctx.barrier()
    l0 = eng.launch_count()
    ms_max, ms_rank = gpu_time(ctx, step, steps, 0)
    launches = eng.launch_count() - l0
    flops = 8.0 * nb * sum((2 * l + 1) ** 2 for l in range(lmax + 1))        # two N x N real x (nb complex) products per l
    d_elems = nb * sum((2 * l + 1) ** 2 for l in range(lmax + 1))            # the D-matrix elements this call stands for
    t = ms_rank * 1e-3
    rec = {"ms_per_step": ms_max, "value": nb_nominal * L2 * ctx.world / (ms_max * 1e-3), "unit": "rotated coefficients/s", "steps": steps,
           "gpu_launches": int(launches), "rotations_per_gpu": nb, "lmax": lmax,
           "equivalent_d_elements_per_s": d_elems / nb * nb_nominal * ctx.world / (ms_max * 1e-3),
           "note": "the same rotations through the matrices would need every d^l(beta) of cfg4 (5.9 TB at 4096 angles); "
                   "equivalent_d_elements_per_s counts the D elements that were never formed"}
    p64 = ctx.fp64.get("peak_tflops")
    if p64:
        rec["roofline"] = {"bound": "tensor", "achieved": flops / t / 1e12, "peak": p64, "unit": "TFLOP/s", "frac": flops / t / 1e12 / p64,
                           "traffic": None, "kernel": "k5_gemm_kernel", "algorithmic_flops_per_launch": flops, "peak_source": ctx.fp64["peak_source"]}
    del out
    torch.cuda.empty_cache()
    ne = min(nb, 1024)
    ho = pinned_empty(ctx, (ne, L2), torch.complex128)
    if ho is not None:
        ha, hb, hg = alpha[:ne].cpu().numpy(), beta[:ne].cpu().numpy(), gamma[:ne].cpu().numpy()
        halm = alm.cpu().numpy()
        dt_max, dt_rank = wall_time(ctx, lambda: eng.rotate_sh(lmax, halm, ha, hb, hg, out=ho.numpy()), 2, 1)
        rec["e2e"] = {"value": ne * L2 * ctx.world / dt_max, "unit": "rotated coefficients/s", "h2d_bytes_per_step": L2 * 16 + ne * 24,
                      "d2h_bytes_per_step": ne * L2 * 16, "steps": 2, "rotations_per_gpu": ne,
                      "kernel_share_of_e2e": (ms_rank * 1e-3 * ne / nb) / dt_rank,
                      "note": "wd_rotate_sh with HOST buffers (coefficients and angles in, rotated coefficients out)"}
        del ho
        release_pinned(ctx)
    return rec


def bench_jmn(ctx, name, cfg, steps, warmup, e2e=True):
    torch, eng = ctx.torch, ctx.eng
    n = cfg["nb"]
    gen = torch.Generator(device=ctx.dev).manual_seed(3 + ctx.rank)
    tj = torch.randint(0, cfg["twoj"] + 1, (n,), device=ctx.dev, generator=gen, dtype=torch.int32)
    u1 = torch.rand(n, device=ctx.dev, generator=gen, dtype=torch.float64)
    u2 = torch.rand(n, device=ctx.dev, generator=gen, dtype=torch.float64)
    tm = (-tj + 2 * torch.floor(u1 * (tj + 1).double()).to(torch.int32)).contiguous()
    tn = (-tj + 2 * torch.floor(u2 * (tj + 1).double()).to(torch.int32)).contiguous()
    beta = torch.rand(n, device=ctx.dev, generator=gen, dtype=torch.float64) * np.pi
    out = torch.empty(n, dtype=torch.float64, device=ctx.dev)
    step = lambda: eng.wignerdjmn_batch(tj, tm, tn, beta, out=out)     # noqa: E731
    for _ in range(warmup):
        step()
    ctx.barrier()
    l0 = eng.launch_count()
    ms_max, ms_rank = gpu_time(ctx, step, steps, 0)
    launches = eng.launch_count() - l0
    terms = float(((tj.double() + 2) // 2).sum().item())          # K = ceil((2j+1)/2) terms of the half-sum per tuple
    t = ms_rank * 1e-3
    rec = {"ms_per_step": ms_max, "value": n * ctx.world / (ms_max * 1e-3), "unit": "tuples/s", "steps": steps,
           "gpu_launches": int(launches), "tuples_per_gpu": n,
           "inputs": "2j ~ U{0..400}, 2m, 2n ~ U{-2j, -2j+2, .., 2j}, beta ~ U[0, pi), torch.Generator(cuda).manual_seed(3 + rank)",
           "l2": "tuples in + results out = 280 MB per step (> 126 MB L2); the eigenvector tables (43 MB for 2j <= 400) are "
                 "L2-resident by design"}
    sin_peak = ctx.fp64.get("sin_evals_per_s")
    if sin_peak:
        rec["roofline"] = {"bound": "fp64_sin", "achieved": terms / t, "peak": sin_peak, "unit": "evaluations/s", "frac": terms / t / sin_peak,
                           "traffic": None, "kernel": "k4_jmn_kernel", "algorithmic_terms_per_launch": terms,
                           "peak_source": "full-range FP64 sin() issue rate of the whole chip, register loop measured in this run "
                                          "(wd_fp64_peak kind 2): one sin or cos per term of the half-sum is the algorithmic minimum "
                                          "of the reference's per-term form (src/WignerD.jl:199)"}
    peak, peak_src = ctx.hbm_peak
    rec["roofline_hbm"] = {"bound": "hbm", "achieved": n * 28 / t / 1e9, "peak": peak, "unit": "GB/s", "frac": n * 28 / t / 1e9 / peak,
                           "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": n * 28}
    if e2e:
        h = [x.cpu().pin_memory() for x in (tj, tm, tn, beta)]
        ho = torch.empty(n, dtype=torch.float64, pin_memory=True)
        dt_max, _ = wall_time(ctx, lambda: eng.wignerdjmn_batch(h[0].numpy(), h[1].numpy(), h[2].numpy(), h[3].numpy(), out=ho.numpy()), 3, 1)
        rec["e2e"] = {"value": n * ctx.world / dt_max, "unit": "tuples/s", "h2d_bytes_per_step": n * 20, "d2h_bytes_per_step": n * 8,
                      "steps": 3, "note": "wd_djmn_batch with pinned HOST buffers"}
        del h, ho
        release_pinned(ctx)
    return rec


def eigen_setup(ctx):
    """K1 (the batched eigensolver that replaces eigen!, :161) alone: 2j = 200; every 2j <= 1024; and the cold first call"""
    import wignerd.jl_b200 as w
    torch = ctx.torch
    res = {}
    eng = w.Engine(ctx.local)
    try:
        eng.prepare([3])                                  # module load / first-launch costs are not the eigensolver's
        for label, lst in (("2j=200", [200]), ("all 2j<=1024", list(range(0, 1025)))):
            reps = []
            for _ in range(4):
                eng.cache_clear()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                eng.prepare(lst)
                reps.append((time.perf_counter() - t0) * 1e3)
            res[label + " ms"] = float(np.median(reps))
            res[label + " ms (all repetitions, in order)"] = [round(x, 3) for x in reps]
        eng.cache_clear()
        b = np.linspace(0, np.pi, 64)
        t0 = time.perf_counter()
        eng.wignerd_batch(100, b)                         # cache miss: K1 + contraction + copies, host buffers
        res["cold wignerd(100, 64 angles) host call ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        eng.wignerd_batch(100, b)
        res["warm wignerd(100, 64 angles) host call ms"] = (time.perf_counter() - t0) * 1e3
        # scalar calls of the reference's API (test/runtests.jl calls them in tight loops)
        eng.wignerdjmn(10, 1, -2, 0.3); eng.wignerd(1, 0.4)
        for label, fn, reps in (("wignerdjmn(10,1,-2,b) us per call", lambda: eng.wignerdjmn(10, 1, -2, 0.3), 300),
                                ("wignerd(1,b) us per call", lambda: eng.wignerd(1, 0.4), 300),
                                ("wignerd(10,NorthPole) us per call", lambda: eng.wignerd(10, w.NorthPole()), 200)):
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            res[label] = (time.perf_counter() - t0) / reps * 1e6
        try:
            from oracle import wignerd_oracle as o
            o.wignerdjmn(10, 1, -2, 0.3)
            t0 = time.perf_counter()
            for _ in range(200):
                o.wignerdjmn(10, 1, -2, 0.3)
            res["numpy oracle wignerdjmn(10,1,-2,b) us per call"] = (time.perf_counter() - t0) / 200 * 1e6
        except Exception:
            pass
    finally:
        eng.close()
    return res


def strong_scaling(ctx, steps):
    """the metric's FIXED workload -- 10^5 angles at j = 100 -- split over the ranks with the product's shard arithmetic"""
    import wignerd.jl_b200 as w
    torch, eng = ctx.torch, ctx.eng
    total, twoj = 100000, 200
    N = twoj + 1
    (lo, hi), (blo, bhi) = w.symmetric_shard(total, ctx.rank, ctx.world)
    idx = torch.cat([torch.arange(lo, hi, device=ctx.dev), torch.arange(blo, bhi, device=ctx.dev)]).to(torch.float64)
    beta = (idx * (np.pi / (total - 1))).contiguous()
    nb = beta.numel()
    out = torch.empty((nb, N, N), dtype=torch.float64, device=ctx.dev)
    ms_max, _ = gpu_time(ctx, lambda: eng.wignerd_batch(100, beta, out=out), steps, 3)
    del out
    res = {"angles_total": total, "angles_this_rank": nb, "ms": ms_max, "value": total * N * N / (ms_max * 1e-3), "unit": "elements/s",
           "sharding": "wignerd.jl_b200.symmetric_shard (16-aligned slabs of the first half + mirror images)"}
    if ctx.world > 1:
        # the N = 1 time of the same workload, measured in this run on rank 0 alone (the other ranks wait)
        ms1 = 0.0
        if ctx.rank == 0:
            b1 = (torch.arange(total, dtype=torch.float64, device=ctx.dev) * (np.pi / (total - 1))).contiguous()
            o1 = torch.empty((total, N, N), dtype=torch.float64, device=ctx.dev)
            for _ in range(3):
                eng.wignerd_batch(100, b1, out=o1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                eng.wignerd_batch(100, b1, out=o1)
            e1.record()
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / steps
            del o1
        ms1 = ctx.allmax(ms1)
        res["ms_n1_same_run"] = ms1
        res["efficiency_vs_n1"] = ms1 / (ctx.world * ms_max)
    torch.cuda.empty_cache()
    return res


def host_copy_ceiling(ctx):
    """plain cudaMemcpyAsync device -> pinned host of 2 GiB per rank, all ranks at once: the ceiling of every e2e number"""
    torch = ctx.torch
    n = 1 << 28
    try:
        d = torch.empty(n, dtype=torch.float64, device=ctx.dev)
        h = torch.empty(n, dtype=torch.float64, pin_memory=True)
        h.copy_(d, non_blocking=True); torch.cuda.synchronize()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 2
        dt_max = ctx.allmax(dt)
        del d, h
        torch.cuda.empty_cache()
        release_pinned(ctx)
        cpus = sorted(os.sched_getaffinity(0))
        return {"d2h_gbs_per_rank": n * 8 / dt / 1e9, "d2h_gbs_aggregate": n * 8 * ctx.world / dt_max / 1e9,
                "host_cpus_visible": len(cpus),
                "note": "2 GiB torch copy_ into pinned memory per rank, concurrently on all ranks; the box exposes one NUMA node to the "
                        "container (nvidia-smi topo: same CPU affinity for every GPU), so NUMA-local pinning is not available"}
    except Exception as e:      # noqa: BLE001
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--nb", type=int, default=0, help="override angles per GPU (debug)")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (wd_set_variant), experiments only")
    ap.add_argument("--grid", default="symmetric", choices=["symmetric", "random"], help="headline beta grid of real-d configs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline config only (kernel experiments)")
    ap.add_argument("--e2e-angles", type=int, default=0, help="(ignored; kept for old command lines)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                   # timing rule: W >= 3
    cfg = dict(CONFIGS[args.config])
    if args.nb:
        cfg["nb"] = args.nb
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    import wignerd.jl_b200 as w

    ctx = Ctx()
    ctx.torch = torch
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(ctx.local)
    ctx.dev = torch.device("cuda", ctx.local)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)

    def barrier():
        if ctx.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=ctx.dev)
        if ctx.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx.barrier, ctx.allmax = barrier, allmax
    ctx.eng = w.Engine(ctx.local)
    ctx.eng.set_variant(args.variant)
    ctx.variant = args.variant
    ctx.hbm_peak = measured_peaks()
    ctx.fp64 = fp64_peaks(ctx)
    extras = args.config == "cfg2" and not args.no_extras and not args.nb and args.variant == 0

    kind = cfg["kind"]
    if kind in ("d", "D"):
        head = bench_matrices(ctx, args.config, cfg, args.steps, args.warmup, grid=args.grid, e2e=True,
                              pageable=extras, clocks=True)
    elif kind == "multi":
        head = bench_multi(ctx, args.config, cfg, max(1, min(args.steps, 5)), 3)
    else:
        head = bench_jmn(ctx, args.config, cfg, args.steps, args.warmup)
    metric, unit = METRIC[kind]
    line = {
        "metric": metric, "value": head["value"], "unit": unit,
        "n_gpus": ctx.world, "steps": head["steps"], "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "name": args.config, "twoj": cfg["twoj"], "variant": args.variant,
                   "inputs": head.get("inputs"), "l2": head.get("l2"), "angles_per_gpu": head.get("angles_per_gpu")},
        "clocks": head.get("clocks"),
        "e2e": head.get("e2e"),
        "gpu_launches": head["gpu_launches"],
        "fp64": dict(ctx.fp64),
    }
    for k in ("roofline", "roofline_hbm", "roofline_tensor", "e2e_pageable"):
        if k in head:
            line[k] = head[k]
    if "roofline_tensor" in head:
        line["fp64"]["alg_tflops_achieved"] = head["roofline_tensor"]["achieved"]
    if line["clocks"] is None:
        line["clocks"] = {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled for this config"]}

    if extras:
        sub = {}
        sub_steps = max(3, min(args.steps, 10))
        sub["cfg1"] = bench_matrices(ctx, "cfg1", CONFIGS["cfg1"], 200, 5)
        sub["cfg2_random_grid"] = bench_matrices(ctx, "cfg2", CONFIGS["cfg2"], sub_steps, 3, grid="random", e2e=False)
        sub["cfg3"] = bench_matrices(ctx, "cfg3", CONFIGS["cfg3"], sub_steps, 3)
        sub["cfg4"] = bench_multi(ctx, "cfg4", CONFIGS["cfg4"], 2 if ctx.world < 4 else 3, 1)
        sub["cfg4_rotate_sh"] = bench_rotate(ctx, 3, 1)
        sub["cfg5"] = bench_jmn(ctx, "cfg5", CONFIGS["cfg5"], sub_steps, 3)
        line["strong"] = strong_scaling(ctx, max(5, min(args.steps, 20)))
        line["host_copy_ceiling"] = host_copy_ceiling(ctx)
        if ctx.rank == 0:
            line["eigen_setup"] = eigen_setup(ctx)
            if not args.no_cpu_baseline:
                for name, rec in sub.items():
                    if name != "cfg4_rotate_sh":
                        rec["cpu_baseline"] = cpu_baseline_for(name.split("_")[0], CONFIGS[name.split("_")[0]])
        line["configs"] = sub
    if ctx.rank == 0 and not args.no_cpu_baseline:
        cb = cpu_baseline_for(args.config, cfg)
        if kind in ("d", "D"):
            try:
                r1, _, d1 = cpu_matrix_rate(cfg["twoj"], 16, 1, kind)
                cb["value_1thread"] = r1
                cb["sample_1thread"] = f"16 angles, {d1:.2f} s on 1 thread"
            except Exception:
                pass
        line["cpu_baseline"] = cb
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.eng.close()
    if ctx.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
