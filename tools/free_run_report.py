"""Print |eps_gpu - eps_reference| next to the reference's own roundoff envelope (P2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity_suite as ps  # noqa: E402
from helpers import golden  # noqa: E402
from perceptron_dqa_b200._lib import Plan  # noqa: E402

fac = lambda N, n_xi, P, dt, chi, b: Plan(N, n_xi, P, dt, chi, batch=b)  # noqa: E731
for name, steps in (("c1_n12_chi10", 44), ("c2_n21_chi10", 24)):
    g = golden(name)
    pl = ps.make(fac, g)
    pl.reset_plus()
    pl.step(steps)
    eps = pl.eps_trace()[0, : steps + 1]
    diff = np.abs(eps - g["eps"][: steps + 1])
    env = ps.reference_envelope(name, steps)
    tr = golden(name + "_truth")["eps_truth"][: steps + 1]
    print(name)
    print(" step |gpu-ref|  |gpu-truth| |ref-truth| env(ref)")
    for s in range(steps + 1):
        print(f" {s:3d}  {diff[s]:.2e}  {abs(eps[s]-tr[s]):.2e}   {abs(g['eps'][s]-tr[s]):.2e}   {env[s]:.2e}")
