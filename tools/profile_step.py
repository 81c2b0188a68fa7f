"""Per-kernel-family device time of a few dQA steps (CUDA events inside libdqa.so).
usage: python tools/profile_step.py --chi 32 --batch 1 --steps 2 [--n 21 --nxi 17]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from perceptron_dqa_b200._lib import Plan  # noqa: E402
from perceptron_dqa_b200.dQA_mps import mydQA  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=21)
    ap.add_argument("--nxi", type=int, default=17)
    ap.add_argument("--chi", type=int, default=32)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--P", type=int, default=1000)
    ap.add_argument("--dt", type=float, default=1.0)
    a = ap.parse_args()
    rng = np.random.default_rng(20230317)
    xi = (2 * rng.integers(0, 2, size=(a.batch, a.nxi, a.n)) - 1).astype(np.int8)
    obj = mydQA(xi if a.batch > 1 else xi[0], P=a.P, dt=a.dt, max_bond=a.chi)
    obj.init_fourier()
    pl = obj.plan
    pl.reset_plus()
    pl.pp = a.P // 2
    t0 = time.time()
    pl.step(a.warmup)
    pl.sync()
    t1 = time.time()
    pl.counters(reset=True)
    pl.profile_begin()
    pl.step(a.steps)
    ms, n = pl.profile_end()
    cnt = pl.counters().sum(0)
    t2 = time.time()
    pl.step(a.steps)
    pl.sync()
    t3 = time.time()
    out = dict(cfg=vars(a), warm_s=t1 - t0, ms_per_step={k: v / a.steps for k, v in ms.items()}, launches=n,
               wall_s_per_step_profiled=(t2 - t1) / a.steps, wall_s_per_step=(t3 - t2) / a.steps,
               rotations_per_svd=float(cnt[0]) / max(1, int(cnt[2])), sweeps_per_svd=float(cnt[1]) / max(1, int(cnt[2])),
               bonds=pl.bonds(0).tolist(), sqr_cycles_per_cta={k: float(cnt[i]) / max(1.0, float(cnt[11])) for k, i in (('gemm_load', 8), ('factor', 9), ('out_formq', 10))},
               lq_cycles_per_cta={k: float(cnt[i]) / max(1.0, float(cnt[15])) for k, i in (('panel', 12), ('tmat', 13), ('trailing', 14))},                flops_per_step_per_inst={k: float(cnt[i]) / a.steps / a.batch for k, i in (('lq', 4), ('jacobi', 5), ('sector_qr', 6), ('theta', 7))}, fp64_peak_tflops=pl.fp64_peak_tflops())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
