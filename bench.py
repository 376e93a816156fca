#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Wigner-d/D engine.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (C port of its hot loop)

Metric (BASELINE.json): Wigner-d matrix elements per second, FP64, at j=100 over 10^5 angles per GPU
(cfg-2 of SURVEY.md section 8(d)); weak scaling: every rank owns its own 10^5-angle slab of an N*10^5 grid,
no data-path collective.  One step = one pass of the hot path (wd_d_batch) over the rank's slab.
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
    # name: (twoj, angles per GPU, complex?)
    "cfg1": dict(twoj=20, nb=1000, kind="d", desc="wignerd(10, beta), 1000 beta in [0, pi]"),
    "cfg2": dict(twoj=200, nb=100000, kind="d", desc="wignerd(100, beta), 10^5 beta in [0, pi] per GPU"),
    "cfg3": dict(twoj=99, nb=100000, kind="D", desc="wignerD(49.5, alpha, beta, gamma), 10^5 Euler triples per GPU"),
    "cfg4": dict(twoj=1024, nb=512, kind="multi", desc="all d^j(beta), j=0..512, on a 512-angle slab per GPU (4096-point grid over 8 GPUs), output ring-buffered"),
    "cfg5": dict(twoj=400, nb=10_000_000, kind="jmn", desc="wignerdjmn, 10^7 random (j<=200, m, n, beta) tuples per GPU"),
}


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


def cpu_port_rate(twoj: int, nang: int, nthreads: int = 0, reps: int = 1, kind="d"):
    """elements/s of the C port of the reference's hot loop (oracle/wignerd_ref.c) on the host cores."""
    from oracle import cref
    from oracle import wignerd_oracle as o
    o.jy_eigen(twoj)                                  # the reference caches the eigenvectors (:163-176): not timed
    N = twoj + 1
    rng = np.random.default_rng(1)
    betas = np.linspace(0, np.pi, nang)
    threads = nthreads or cref.num_threads()
    if kind == "D":
        a, g = rng.uniform(0, 2 * np.pi, nang), rng.uniform(0, 2 * np.pi, nang)
        cref.wignerD_batch(twoj, a[:threads], betas[:threads], g[:threads], threads)
        t0 = time.perf_counter()
        for _ in range(reps):
            cref.wignerD_batch(twoj, a, betas, g, threads)
    else:
        cref.wignerd_batch(twoj, betas[:threads], threads)  # warm-up
        t0 = time.perf_counter()
        for _ in range(reps):
            cref.wignerd_batch(twoj, betas, threads)
    dt = (time.perf_counter() - t0) / reps
    return nang * N * N / dt, threads, dt


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU algorithm (C port; Julia is not in the image), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref
    threads = cref.num_threads()
    twoj = cfg["twoj"]
    N = twoj + 1
    nang = max(threads * 2, 16)                        # bounded sample per step (~80 ms of wall time at j=100)
    if cfg["kind"] == "jmn":
        return run_reference_jmn(args, cfg, threads)
    from oracle import wignerd_oracle as o
    o.jy_eigen(twoj)                                   # cached once per j by the reference (:163-176): not timed
    betas = np.linspace(0, np.pi, nang)
    if cfg["kind"] == "D":
        rng = np.random.default_rng(2)
        a, g = rng.uniform(0, 2 * np.pi, nang), rng.uniform(0, 2 * np.pi, nang)
        fn = lambda: cref.wignerD_batch(twoj, a, betas, g, threads)      # noqa: E731
    else:
        fn = lambda: cref.wignerd_batch(twoj, betas, threads)            # noqa: E731
    for _ in range(max(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    value = nang * N * N / dt
    sample = f"{nang} of {cfg['nb']} angles per step (cost is linear in #angles once the eigenvectors are cached)"
    line = {
        "impl": "reference", "metric": "wigner_d_elements_per_sec_fp64", "value": value, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "name": args.config, "twoj": twoj, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C port of src/WignerD.jl:187-206,:266-295 (oracle/wignerd_ref.c); Julia is not installed in this image",
    }
    print(json.dumps(line), flush=True)


def run_reference_jmn(args, cfg, threads):
    from oracle import cref
    rng = np.random.default_rng(3)
    n = 20000
    tj = rng.integers(0, cfg["twoj"] + 1, n).astype(np.int32)
    tm = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
    tn = (-tj + 2 * rng.integers(0, tj + 1)).astype(np.int32)
    b = rng.uniform(0, np.pi, n)
    cref.wignerdjmn_batch(tj, tm, tn, b, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cref.wignerdjmn_batch(tj, tm, tn, b, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt
    sample = f"{n} of {cfg['nb']} tuples per step"
    print(json.dumps({
        "impl": "reference", "metric": "wignerdjmn_tuples_per_sec_fp64", "value": value, "unit": "tuples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "tuples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tuples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--nb", type=int, default=0, help="override angles per GPU (debug)")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 plain FP64-ALU kernel, 2 DMMA kernel")
    ap.add_argument("--e2e-angles", type=int, default=0, help="angles per e2e step (0 = the whole workload, nb)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = w.Engine(local)
    eng.set_variant(args.variant)
    twoj, nb = cfg["twoj"], cfg["nb"]
    N = twoj + 1
    # this rank's slab of the global synthetic grid (weak scaling: nb angles per rank)
    gtotal = nb * world
    if cfg["kind"] == "jmn":
        return bench_jmn(args, cfg, eng, torch, dist, rank, world, local, barrier)
    if cfg["kind"] == "multi":
        return bench_multi(args, cfg, eng, torch, dist, rank, world, local, barrier)
    lo = rank * nb
    beta = (torch.arange(lo, lo + nb, dtype=torch.float64, device=dev) * (np.pi / max(gtotal - 1, 1))).contiguous()
    cplx = cfg["kind"] == "D"
    if cplx:
        gen = torch.Generator(device=dev).manual_seed(2 + rank)
        alpha = torch.rand(nb, dtype=torch.float64, device=dev, generator=gen) * (2 * np.pi)
        gamma = torch.rand(nb, dtype=torch.float64, device=dev, generator=gen) * (2 * np.pi)
        beta = torch.rand(nb, dtype=torch.float64, device=dev, generator=gen) * np.pi
        out = torch.empty((nb, N, N), dtype=torch.complex128, device=dev)
        step = lambda: eng.wignerD_batch(twoj / 2 if twoj % 2 else twoj // 2, alpha, beta, gamma, out=out)   # noqa: E731
    else:
        out = torch.empty((nb, N, N), dtype=torch.float64, device=dev)
        step = lambda: eng.wignerd_batch(twoj // 2 if twoj % 2 == 0 else twoj / 2, beta, out=out)           # noqa: E731
    elems = nb * N * N
    bytes_per_launch = elems * (16 if cplx else 8)

    # FP64 denominators measured on this box, this run: DMMA/DFMA register loops and cuBLAS DGEMM
    fp64 = {}
    if rank == 0:
        try:
            fp64["dmma_loop_tflops"] = eng.fp64_peak(0, 20000)
            fp64["dfma_loop_tflops"] = eng.fp64_peak(1, 20000)
            a = torch.randn(6144, 6144, dtype=torch.float64, device=dev)
            torch.matmul(a, a)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                torch.matmul(a, a)
            e1.record()
            torch.cuda.synchronize()
            fp64["cublas_dgemm_tflops"] = 3 * 2 * 6144 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
            del a
        except Exception as e:      # noqa: BLE001
            fp64["error"] = repr(e)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = elems * world / (ms_per_step * 1e-3)

    # ---- e2e: the same call through the public API with HOST buffers (pinned), copies inside the timed region
    # (whole workload per step: 32 GB of pinned host memory at cfg-2; if the host cannot pin that much, a 16384-angle
    # slice is used instead and the line says so in config.e2e_angles_per_gpu)
    ne = min(args.e2e_angles, nb) if args.e2e_angles > 0 else nb
    try:                                   # all ranks of the box pin their buffers at once: keep well inside the host's RAM
        import psutil
        need = ne * N * N * (16 if cplx else 8) * int(os.environ.get("LOCAL_WORLD_SIZE", world))
        if ne > 16384 and psutil.virtual_memory().available < 2 * need:
            ne = 16384
    except Exception:
        pass
    h_out = None
    while h_out is None:
        try:
            h_out = torch.empty((ne, N, N), dtype=torch.complex128 if cplx else torch.float64, pin_memory=True)
        except RuntimeError:
            if ne <= 16384:
                raise
            ne = 16384
            torch.cuda.empty_cache()
    h_beta = torch.empty(ne, dtype=torch.float64).pin_memory()
    h_beta.copy_(beta[:ne].cpu())
    jv = twoj // 2 if twoj % 2 == 0 else twoj / 2
    if cplx:
        h_a = alpha[:ne].cpu().pin_memory(); h_g = gamma[:ne].cpu().pin_memory()
        e2e_step = lambda: eng.wignerD_batch(jv, h_a.numpy(), h_beta.numpy(), h_g.numpy(), out=h_out.numpy())   # noqa: E731
    else:
        e2e_step = lambda: eng.wignerd_batch(jv, h_beta.numpy(), out=h_out.numpy())                              # noqa: E731
    e2e_steps = max(3, min(args.steps, 10 if ne <= 16384 else 4))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()                                    # synchronous: returns when the result is in host memory
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_rate = ne * N * N * world / (float(t.item()) / e2e_steps)
    h2d = ne * 8 * (3 if cplx else 1)
    d2h = ne * N * N * (16 if cplx else 8)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    achieved = bytes_per_launch / (ms / args.steps * 1e-3) / 1e9          # this rank's kernel, GB/s
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.config)
        except Exception:
            traffic = None
    K = (N + 1) // 2
    W = ((twoj // 2 + 1) ** 2) if twoj % 2 == 0 else ((twoj + 1) // 2) * ((twoj + 3) // 2)
    f_alg = 2.0 * W * K * nb
    fp64["alg_tflops_achieved"] = f_alg / (ms / args.steps * 1e-3) / 1e12
    line = {
        "metric": "wigner_d_elements_per_sec_fp64", "value": value, "unit": "elements/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "name": args.config, "twoj": twoj, "angles_per_gpu": nb,
                   "output_bytes_per_gpu": bytes_per_launch,
                   "l2": ("outputs (>=16 GB per step) far exceed the 126 MB L2; no flush needed" if bytes_per_launch > 4 * 126e6 else
                          "L2-RESIDENT timing: the step's output fits in the 126 MB L2 and is rewritten every step without a flush "
                          "(CPU-runnable sanity config, launch/latency bound; not a bandwidth claim)"),
                   "variant": args.variant, "e2e_angles_per_gpu": ne,
                   "inputs": ("alpha, gamma ~ U[0, 2pi), beta ~ U[0, pi], torch.Generator(cuda).manual_seed(2 + rank)" if cplx
                              else "beta = rank's slab of linspace(0, pi, n_gpus * angles_per_gpu)")},
        "clocks": clocks,
        "e2e": {"value": e2e_rate, "unit": "elements/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "note": "wd_d_batch with pinned HOST buffers; PCIe copies inside the timed region"},
        "gpu_launches": int(launches),
        "fp64": fp64,
    }
    # Binding roof = the SLOWER of the two: HBM write of the algorithmic bytes at the measured copy
    # bandwidth vs the algorithmic flops F_alg = 2 W K nb at the FP64 tensor peak measured in THIS run (cuBLAS DGEMM;
    # MEASURED_PEAKS.json carries no FP64 figure).  Both are reported; "roofline" is the binding one.
    t_step = ms / args.steps * 1e-3
    kname = "k2_cplx_kernel" if cplx else ("k2_real_kernel" if twoj <= 223 else "k2_stream_kernel")
    rl_hbm = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
              "traffic": traffic, "peak_source": peak_src, "kernel": kname,
              "algorithmic_bytes_per_launch": bytes_per_launch}
    p64 = fp64.get("cublas_dgemm_tflops") or fp64.get("dmma_loop_tflops")
    line["roofline_hbm"] = rl_hbm
    line["roofline"] = rl_hbm
    if p64:
        ach64 = f_alg / t_step / 1e12
        rl_t = {"bound": "tensor", "achieved": ach64, "peak": p64, "unit": "TFLOP/s", "frac": ach64 / p64,
                "traffic": traffic, "kernel": kname, "algorithmic_flops_per_launch": f_alg,
                "peak_source": "cuBLAS DGEMM 6144^3 (torch.matmul float64) measured in this run; DMMA.8x8x4 register loop: "
                               f"{fp64.get('dmma_loop_tflops', float('nan')):.1f} TFLOP/s"}
        line["roofline_tensor"] = rl_t
        if f_alg / (p64 * 1e12) > bytes_per_launch / (peak * 1e9):
            line["roofline"] = rl_t
    if not args.no_cpu_baseline:
        try:
            nang = 512 if twoj >= 150 else 4096
            rate, threads, dt_cpu = cpu_port_rate(twoj, nang, 0, kind=cfg["kind"])
            n1 = max(8, nang // 32)
            rate1, _, dt1 = cpu_port_rate(twoj, n1, 1, kind=cfg["kind"])
            line["cpu_baseline"] = {"value": rate, "unit": "elements/s", "cores": threads, "kind": "port",
                                    "sample": f"{nang} of {nb} angles, {dt_cpu:.2f} s wall on {threads} host threads "
                                              "(C port of the reference loop, oracle/wignerd_ref.c; Julia absent)",
                                    "value_1thread": rate1, "sample_1thread": f"{n1} angles, {dt1:.2f} s on 1 thread"}
        except Exception as e:      # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "elements/s", "cores": 0, "kind": "port", "sample": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_multi(args, cfg, eng, torch, dist, rank, world, local, barrier):
    """cfg-4: every integer j <= 512 on this rank's 512-angle slab of the 4096-point grid; the 0.74 TB of output per GPU
    is consumed in place (each j overwrites one reusable 4.3 GB buffer -- the ring buffer of SURVEY 7.7)."""
    dev = torch.device("cuda", local)
    nb, jmax = cfg["nb"], cfg["twoj"] // 2
    gtotal = nb * world
    beta = (torch.arange(rank * nb, rank * nb + nb, dtype=torch.float64, device=dev) * (np.pi / max(gtotal - 1, 1))).contiguous()
    js = list(range(0, jmax + 1))
    buf = torch.empty(nb * (2 * jmax + 1) ** 2, dtype=torch.float64, device=dev)
    offsets = np.zeros(len(js), dtype=np.int64)
    elems = nb * sum((2 * j + 1) ** 2 for j in js)
    eng.prepare([2 * j for j in js])
    step = lambda: eng.wignerd_multi(js, beta, out=buf, offsets=offsets)      # noqa: E731
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        msps = float(t.item()) / args.steps
        f_alg = 2.0 * nb * sum((j + 1) ** 2 * (j + 1) for j in js)
        print(json.dumps({"metric": "wigner_d_elements_per_sec_fp64", "value": elems * world / (msps * 1e-3), "unit": "elements/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": msps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic",
                          "config": {"workload": cfg["desc"], "name": args.config,
                                     "l2": "every j writes nb*(2j+1)^2*8 bytes into one reused buffer (4.3 GB at j=512): "
                                           "outputs exceed the L2 for j >= 90, which is where ~all of the time goes"},
                          "gpu_launches": int(eng.launch_count() - l0),
                          "fp64": {"alg_tflops_achieved": f_alg / (msps * 1e-3) / 1e12}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_jmn(args, cfg, eng, torch, dist, rank, world, local, barrier):
    dev = torch.device("cuda", local)
    n = cfg["nb"]
    gen = torch.Generator(device=dev).manual_seed(3 + rank)
    tj = torch.randint(0, cfg["twoj"] + 1, (n,), device=dev, generator=gen, dtype=torch.int32)
    u1 = torch.rand(n, device=dev, generator=gen, dtype=torch.float64)
    u2 = torch.rand(n, device=dev, generator=gen, dtype=torch.float64)
    tm = (-tj + 2 * torch.floor(u1 * (tj + 1).double()).to(torch.int32)).contiguous()
    tn = (-tj + 2 * torch.floor(u2 * (tj + 1).double()).to(torch.int32)).contiguous()
    beta = torch.rand(n, device=dev, generator=gen, dtype=torch.float64) * np.pi
    out = torch.empty(n, dtype=torch.float64, device=dev)
    step = lambda: eng.wignerdjmn_batch(tj, tm, tn, beta, out=out)     # noqa: E731
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        msps = float(t.item()) / args.steps
        print(json.dumps({"metric": "wignerdjmn_tuples_per_sec_fp64", "value": n * world / (msps * 1e-3), "unit": "tuples/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": msps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic",
                          "config": {"workload": cfg["desc"], "name": args.config,
                                     "l2": "tuples in + results out = 280 MB per step (> 126 MB L2); the eigenvector tables "
                                           "(43 MB for 2j <= 400) are L2-resident by design"},
                          "gpu_launches": int(eng.launch_count() - l0)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
