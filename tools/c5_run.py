"""C5 (BASELINE.json config 5): the synthetic large instance N=64, N_xi=51, chi=64, P=1000 schedule, steps at p ~ P/2.
Writes a driver-reproducible record (per-family ms, executed flops, status, bonds, eps) as one JSON line.

usage: python tools/c5_run.py [--batch 4] [--steps 3] [--warmup 1] [--chi 64] > gpurun_out/r02_c5.json
Patterns: default_rng(64), iid +-1.  Checks that can be made at this size (no CPU oracle exists for C5:
days per step): eps(0) closed form, unit norm after every step, status words clean."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from perceptron_dqa_b200.dQA_mps import mydQA  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--chi", type=int, default=64)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--nxi", type=int, default=51)
    a = ap.parse_args()
    rng = np.random.default_rng(64)
    xi = (2 * rng.integers(0, 2, size=(max(a.batch, 1), a.nxi, a.n)) - 1).astype(np.int8)
    P, dt = 1000, 1.0
    obj = mydQA(xi if a.batch > 1 else xi[0], P=P, dt=dt, max_bond=a.chi)
    obj.init_fourier()
    pl = obj.plan
    pl.reset_plus()
    eps0 = float(pl.eps_trace()[0, 0].real)
    pl.pp = P // 2
    t0 = time.time()
    pl.step(a.warmup)
    pl.sync()
    t_warm = time.time() - t0
    pl.counters(reset=True)
    pl.profile_begin()
    t0 = time.time()
    pl.step(a.steps)
    ms, n = pl.profile_end()
    wall = time.time() - t0
    cnt = pl.counters().sum(0).astype(np.float64)
    eps = pl.eps_trace()[:, P // 2 + a.warmup + 1: P // 2 + a.warmup + a.steps + 1].real
    # closed form of eps(0): (N_xi/N) sum_x C(N,x) 2^-N f(x)  
    from math import comb
    f = [max(0.0, 2 * x - a.n) / np.sqrt(a.n) for x in range(a.n + 1)]
    eps0_exact = a.nxi / a.n * sum(comb(a.n, x) * 2.0 ** -a.n * f[x] for x in range(a.n + 1))
    fl = {"lq": cnt[4], "jacobi": cnt[5], "sector_qr": cnt[6], "theta": cnt[7]}
    out = {
        "config": f"C5: N={a.n}, N_xi={a.nxi}, chi={a.chi}, P={P}, dt={dt}, batch {a.batch}, steps at p=P/2 after {a.warmup} warm-up step(s)",
        "s_per_step": wall / a.steps, "realisation_steps_per_s": a.batch * a.steps / wall, "warmup_s": t_warm,
        "ms_per_step": {k: v / a.steps for k, v in ms.items()}, "launches_per_step": {k: v // a.steps for k, v in n.items()},
        "executed_gflop_per_realisation_step": {k: v / a.steps / a.batch / 1e9 for k, v in fl.items()},
        "achieved_tflops": {k: (v / (ms[k] * 1e-3) / 1e12 if ms.get(k) else None) for k, v in fl.items()},
        "sweeps_per_svd": float(cnt[1]) / max(1.0, float(cnt[2])),
        "bonds_inst0": pl.bonds(0).tolist(), "status_nonzero": int(np.count_nonzero(pl.status())),
        "eps0": eps0, "eps0_closed_form": eps0_exact, "eps_after_steps_inst0": eps[0].tolist(),
        "lq_panel_rows": None, "workspace_gib": pl.workspace_bytes / 2 ** 30,
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
