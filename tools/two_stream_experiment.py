"""Do two half-batches on two streams overlap better than one full batch? (experiment)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dqa_oracle as O
from perceptron_dqa_b200._lib import Plan
N, nxi, P, dt, chi = 21, 17, 1000, 1.0, 32
uk, fx = O.fourier_tables(N, P, dt)
rng = np.random.default_rng(20230317)
def mk(batch, stream):
    xi = (2 * rng.integers(0, 2, size=(batch, nxi, N)) - 1).astype(np.int8)
    pl = Plan(N, nxi, P, dt, chi, batch=batch, stream=stream)
    pl.set_patterns(xi); pl.set_fourier(uk, fx); pl.reset_plus(); pl.pp = P // 2
    return pl
for nsplit, total in ((1, 888), (2, 888), (3, 888), (2, 1776)):
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    plans = [mk(total // nsplit, s.cuda_stream) for s in streams]
    for pl in plans: pl.step(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 3
    for _ in range(K):
        for pl in plans: pl.step_async(1)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / K
    print(f"splits={nsplit} total={total}: {t*1e3:.1f} ms/step -> {total/t:.1f} realisation-steps/s", flush=True)
    del plans
