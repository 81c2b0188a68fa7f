"""TEST INFRASTRUCTURE ONLY -- copy the reference's own stored result vectors into one small fixture.

The reference keeps no test-suite; its validation is "stored curves agree across implementations"
(SURVEY section 4).  These are those stored vectors for the hot path (complex64 jax runs and one complex128
quimb run), plus the three N=21 pattern files and the published eps(1) tables -- inputs and envelope targets
for tests/test_gpu_parity.py::test_p4_*.  Run in the build container:  python oracle/make_reference_curves.py
"""
import csv
import os

import numpy as np

REF = os.environ.get("DQA_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_curves.npz")
B = os.path.join(REF, "data", "benchmark")
out = {}
for i in (1, 2, 3):
    out[f"patterns_17_21_{i}"] = np.load(os.path.join(REF, "data", f"patterns_17-21.{i}.npy")).astype(np.int8)
out["patterns_12_15"] = np.load(os.path.join(REF, "data", "patterns_12-15.npy")).astype(np.int8)
for bd in (10, 20):
    out[f"numpy_17_21_P100_dt1_bd{bd}"] = np.load(os.path.join(B, "numpy_17-21_loss", f"17-21_full-loss_bd{bd}.npy"))
out["numpy_9_12_P100_dt05_bd10"] = np.load(os.path.join(B, "numpy_9-12_loss", "dt0.5_bd10.npy"))
out["quimb_9_12_P100_dt05_bd10"] = np.load(os.path.join(B, "quimb_9-12_loss_dt0.5_bd10.npy"))
out["numpy_12_15_P100_dt12_bd10"] = np.load(os.path.join(B, "numpy_12-15_loss_dt1.2_bd10.npy"))


def table(path):
    rows = []
    with open(path, encoding="utf-8-sig") as f:
        for r in csv.DictReader(f):
            try:
                rows.append((int(r["P"]), float(r["dt"]), int(r["max_bond_dim"]), float(r["output"])))
            except (ValueError, KeyError):
                pass
    return np.array(rows, dtype=np.float64)


out["table_numpy_17_21"] = table(os.path.join(B, "numpy_17-21.csv"))   # columns: P, dt, chi, eps(1)  (jax, complex64)
out["table_quimb_17_21"] = table(os.path.join(B, "quimb_17-21.csv"))   # quimb, complex128
pm = np.loadtxt(os.path.join(REF, "data", "paper_7a_mps1000.csv"), delimiter=",", skiprows=1)
out["paper_7a_mps1000"] = pm                                            # dt, eps(1) digitised from the paper
np.savez_compressed(OUT, **out)
print("wrote", OUT, {k: np.asarray(v).shape for k, v in out.items()})
