"""N=21 runs in the regime of the paper's Fig. 7a / the reference's numpy_17-21.csv: eps(1) for the three shipped
pattern realisations at chi = 10, 20, next to the reference's published numbers (complex64 jax, complex128 quimb,
digitised paper curve).  Each (P, dt, chi) is one batched plan of 3 realisations."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from perceptron_dqa_b200.dQA_mps import mydQA  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "reference_curves.npz"))
xi = np.stack([g[f"patterns_17_21_{i}"] for i in (1, 2, 3)])
tab = g["table_numpy_17_21"]
qt = g["table_quimb_17_21"]
paper = g["paper_7a_mps1000"]


def lookup(t, P, dt):
    m = (t[:, 0] == P) & (np.abs(t[:, 1] - dt) < 1e-9)
    return float(t[m, 3][0]) if m.any() else None


runs = [(100, 1.0, 10), (100, 1.0, 20), (1000, 0.5, 10), (1000, 1.0, 10), (1000, 1.0, 20)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    runs = runs[:2]
for P, dt, chi in runs:
    t0 = time.time()
    obj = mydQA(xi, P=P, dt=dt, max_bond=chi)
    obj.init_fourier()
    obj.run(print_info=False)
    e1 = obj.eps[:, -1].real
    rec = {"P": P, "dt": dt, "chi": chi, "eps1_realisations_1_2_3": e1.tolist(), "mean": float(e1.mean()),
           "ref_jax_c64": lookup(tab, P, dt) if chi == 10 else None, "ref_quimb_c128": lookup(qt, P, dt) if chi == 10 else None,
           "paper_mps": float(paper[np.argmin(np.abs(paper[:, 0] - dt)), 1]) if P == 1000 else None,
           "status": obj.plan.status().tolist(), "seconds": round(time.time() - t0, 1)}
    if P == 100 and dt == 1.0:
        stored = g[f"numpy_17_21_P100_dt1_bd{chi}"]
        rec["max_abs_diff_to_stored_curve_realisation_1"] = float(np.abs(obj.eps[0] - stored).max())
        rec["stored_curve_eps1"] = float(stored[-1].real)
    print(json.dumps(rec), flush=True)
