"""Config C2 as the reference is actually driven (benchmarker-mps.py:32-43): ONE realisation, the full P=1000 run.
usage: python tools/c2_full_run.py [chi ...]   -> one JSON line per bond dimension (wall time, steps/s, eps(1))
Patterns: data/patterns_17-21.1.npy (stored in tests/golden/c2_n21_chi10.npz), dt = 1.0 (paper Fig. 7a regime)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from perceptron_dqa_b200.dQA_mps import mydQA  # noqa: E402

xi = np.load(os.path.join(ROOT, "tests", "golden", "c2_n21_chi10.npz"))["xi"]
P, dt = 1000, 1.0
for chi in [int(c) for c in sys.argv[1:]] or [10, 20]:
    obj = mydQA(xi, P=P, dt=dt, max_bond=chi)
    obj.init_fourier()
    obj.plan.reset_plus()
    obj.plan.step(2)  # graph capture / first-launch cost outside the timed run
    t0 = time.perf_counter()
    obj.run(print_info=False)
    wall = time.perf_counter() - t0
    eps = np.array([v for _s, v in obj.loss])
    print(json.dumps({"config": f"C2: patterns_17-21.1, N=21, N_xi=17, P={P}, dt={dt}, chi={chi}, batch 1, full run from |+>",
                      "graph": os.environ.get("DQA_GRAPH", "1") != "0", "wall_s": wall, "steps_per_s": P / wall,
                      "eps1": float(eps[-1].real), "eps_half": float(eps[P // 2].real), "max_abs_imag": float(np.abs(eps.imag).max()),
                      "bd_mid_max": int(max(obj.bd_monitor)), "status_nonzero": int(np.count_nonzero(obj.plan.status())),
                      "launches": obj.plan.launch_count()}), flush=True)
