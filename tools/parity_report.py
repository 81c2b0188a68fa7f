"""Teacher-forced parity numbers per fixture: worst fidelity defect, Schmidt-spectrum and
eps differences between the device sub-step and the unmodified reference's stored post-state."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import dqa_oracle as O  # noqa: E402
import parity_suite as ps  # noqa: E402
from helpers import golden, snapshots  # noqa: E402
from perceptron_dqa_b200._lib import Plan  # noqa: E402

fac = lambda N, n_xi, P, dt, chi, b: Plan(N, n_xi, P, dt, chi, batch=b)  # noqa: E731
print("fixture                 N  chi substeps  max|1-F|   max|dSchmidt|  max|d eps|   bond dims equal")
for name in ("n7_chi4", "n7_chi8_untruncated", "c1_n12_chi10", "n15_chi10", "c2_n21_chi10", "c3_n21_chi32",
             "n12_chi48_saturated", "n12_chi64_untruncated", "n14_chi64_saturated", "n32_chi8_synthetic"):
    g = golden(name)
    pl = ps.make(fac, g)
    wf = ws = we = 0.0
    same = True
    snaps = snapshots(g)
    for p, mu, pre, post in snaps:
        pl.set_mps(pre, 0)
        pl.apply_pattern(p, mu)
        got = pl.get_mps(0)
        same &= [t.shape for t in got] == [t.shape for t in post]
        wf = max(wf, abs(O.fidelity_defect(got, post)))
        for a, b in zip(O.schmidt_spectra(got), O.schmidt_spectra(post)):
            n = min(len(a), len(b))
            ws = max(ws, float(np.abs(a[:n] - b[:n]).max()))
        we = max(we, abs(O.compute_loss_fast(got, g["xi"], g["fxft"]) - O.compute_loss_fast(post, g["xi"], g["fxft"])))
    print(f"{name:22s} {g['xi'].shape[1]:3d} {int(g['chi']):4d} {len(snaps):6d}    {wf:.2e}   {ws:.2e}      {we:.2e}    {same}")
