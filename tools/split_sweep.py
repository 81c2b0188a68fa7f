"""Step time vs DQA_SPLIT (sub-batches on parallel branches of the step graph).  usage: split_sweep.py chi:batch[:splits,...] ..."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dqa_oracle as O
from perceptron_dqa_b200._lib import Plan
N, nxi, P, dt = 21, 17, 1000, 1.0
for arg in sys.argv[1:]:
    f = arg.split(":")
    chi, batch = int(f[0]), int(f[1])
    splits = [int(x) for x in f[2].split(",")] if len(f) > 2 else [1, 2, 3]
    uk, fx = O.fourier_tables(N, P, dt)
    rng = np.random.default_rng(20230317)
    xi = (2 * rng.integers(0, 2, size=(batch, nxi, N)) - 1).astype(np.int8)
    for ns in splits:
        os.environ["DQA_SPLIT"] = str(ns)
        pl = Plan(N, nxi, P, dt, chi, batch=batch)
        pl.set_patterns(xi); pl.set_fourier(uk, fx); pl.reset_plus(); pl.pp = P // 2
        pl.step(2)
        torch.cuda.synchronize()
        K = 3 if chi >= 32 else 6
        t0 = time.perf_counter(); pl.step(K); torch.cuda.synchronize(); t = (time.perf_counter() - t0) / K
        print(json.dumps({"chi": chi, "batch": batch, "split": ns, "ms_per_step": round(t * 1e3, 2), "rs_per_s": round(batch / t, 1)}), flush=True)
        pl.close()
