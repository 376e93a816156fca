"""Regenerates tests/golden/*.json.  Run from the repo root:
    python tests/golden/make_golden.py

* readme_vectors.json   -- the known answers printed in the reference's
  README.md:13-34, typed in by hand (Julia is not available to re-run them).
* mp_exact_samples.json -- exact multiprecision d^j_mn(beta) samples from
  oracle/mp_exact.py (explicit Wigner sum, >= 60+0.7*2j digits) at the BASELINE
  j values, seeded; they triangulate the oracle where the reference has no
  tests (j > 5).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.mp_exact import wigner_d_exact  # noqa: E402

readme = {
    "source": "/root/reference/README.md:13-34",
    "wignerd_0.5_0": [[1.0, -0.0], [0.0, 1.0]],                       # :13-16 (printed, 6 sig. digits)
    "wignerd_1_pi_3": [[0.75, 0.612372, 0.25],                        # :18-22
                       [-0.612372, 0.5, 0.612372],
                       [0.25, -0.612372, 0.75]],
    "wignerdjmn_1_1_1_pi_3": 0.7500000000000004,                      # :24-25 (full precision)
    "wignerD_1_0_pi_3_pi_2": [                                        # :27-31 (re, im) pairs, 6 sig. digits
        [[4.59243e-17, 0.75], [0.612372, -0.0], [1.53081e-17, -0.25]],
        [[-3.7497e-17, -0.612372], [0.5, 0.0], [3.7497e-17, -0.612372]],
        [[1.53081e-17, 0.25], [-0.612372, 0.0], [4.59243e-17, -0.75]]],
    "wignerDjmn_1_1_1_0_pi_3_pi_2": [4.592425496802577e-17, -0.7500000000000004],   # :33-34
}
with open(os.path.join(HERE, "readme_vectors.json"), "w") as f:
    json.dump(readme, f, indent=1)

rng = np.random.default_rng(20261019)
samples = []
for twoj, count in [(1, 8), (4, 16), (20, 60), (99, 60), (200, 60), (400, 32), (1024, 16)]:
    for _ in range(count):
        mi, ni = rng.integers(0, twoj + 1, 2)
        beta = float(rng.uniform(0.0, np.pi)) if rng.random() < 0.8 else float(rng.uniform(-31.0, 31.0))
        twom, twon = int(2 * mi - twoj), int(2 * ni - twoj)
        samples.append({"twoj": twoj, "twom": twom, "twon": twon, "beta": beta.hex(),
                        "d": wigner_d_exact(twoj, twom, twon, beta).hex()})
with open(os.path.join(HERE, "mp_exact_samples.json"), "w") as f:
    json.dump({"seed": 20261019, "generator": "tests/golden/make_golden.py", "samples": samples}, f, indent=0)
print(len(samples), "samples written")
