"""where does the time of Engine.prepare(all 2j <= 1024) go when the process already holds a lot of device / pinned memory
(bench.py's eigen_setup saw 0.16 - 2.5 s against 12 ms in a fresh process)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wignerd.jl_b200 as w

def t_prepare(eng, tag):
    for rep in range(3):
        eng.cache_clear(); torch.cuda.synchronize()
        t0 = time.perf_counter(); eng.prepare(list(range(1025))); dt = (time.perf_counter() - t0) * 1e3
        print(f"{tag}: rep {rep}: {dt:.1f} ms", flush=True)

eng = w.Engine(0); eng.prepare([3])
t_prepare(eng, "fresh")
big = torch.empty(int(100e9), dtype=torch.uint8, device="cuda"); big.zero_(); torch.cuda.synchronize()
t_prepare(eng, "100 GB device tensor held")
pin = torch.empty(int(8e9), dtype=torch.uint8).pin_memory()
t_prepare(eng, "+ 8 GB pinned")
del big; torch.cuda.empty_cache(); torch.cuda.synchronize()
t_prepare(eng, "device tensor freed (empty_cache)")
del pin
import gc; gc.collect()
from torch._C import _host_emptyCache
_host_emptyCache()
t_prepare(eng, "pinned freed")
eng2 = w.Engine(0); eng2.prepare([3])
t_prepare(eng2, "second engine")
