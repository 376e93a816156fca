"""very large j: eigensolver residual/orthogonality and d orthogonality/trace (no oracle needed)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wignerd.jl_b200 as w
e = w.Engine(0)
for twoj in (1024, 1100, 1200, 1300, 1401, 2000, 3000):
    N = twoj + 1
    u = e.eigvecs(twoj / 2 if twoj % 2 else twoj // 2)
    k = np.arange(1, N); b = 0.5 * np.sqrt(k * (N - k))
    Tu = np.zeros_like(u); Tu[:-1] += b[:, None] * u[1:]; Tu[1:] += b[:, None] * u[:-1]
    lam = np.arange(N) - twoj / 2
    res = np.abs(Tu - u * lam).max(); orth = np.abs(u.T @ u - np.eye(N)).max()
    bad = np.argmax(np.abs(u.T @ u - np.eye(N)).max(axis=0))
    d = e.wignerd_batch(twoj / 2 if twoj % 2 else twoj // 2, np.array([0.7]))[0]
    m = lam
    print(twoj, "eig res", res, "orth", orth, "worst col", bad, "| d orth", np.abs(d.T @ d - np.eye(N)).max(), "trace err", abs(np.trace(d) - np.cos(m * 0.7).sum()), flush=True)
