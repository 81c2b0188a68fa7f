"""TEST INFRASTRUCTURE ONLY -- exact-arithmetic (80-bit) eps(s) for the P2 envelope.
    OMP_NUM_THREADS=1 python oracle/make_truth.py c1 c2
Writes tests/golden/<config>_truth.npz with `eps_truth` (see oracle/extended_precision.py)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import extended_precision as X  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
JOBS = {"c1": ("c1_n12_chi10", 44), "c2": ("c2_n21_chi10", 24), "n7": ("n7_chi4", 20)}

for key in sys.argv[1:]:
    name, steps = JOBS[key]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    t = time.time()
    tr = X.run_truth(g["xi"], int(g["P"]), float(g["dt"]), int(g["chi"]), steps)
    np.savez_compressed(os.path.join(GOLD, name + "_truth.npz"), eps_truth=tr, steps=steps,
                        note="80-bit long double run of the same map; oracle/extended_precision.py")
    print(name, steps, "steps", round(time.time() - t, 1), "s; |ref-truth| =", np.abs(tr - g["eps"][: steps + 1])[-1], flush=True)
