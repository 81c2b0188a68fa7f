"""Do sub-batches in DIFFERENT phases of the step share an SM better than one batch in lock step? (experiment)

S plans of batch total/S on S streams; plan i starts `i * stagger_ms` late (a device-side spin), so that while one
sub-batch is in its canonisation sweep (sector QR) the others are in their truncation sweeps (Theta / LQ / Jacobi).
DQA_JAC_SMEM_PAD (read by libdqa at plan creation) pads the Jacobi's shared memory so that only one of its CTAs fits an SM.
usage: tenant_experiment.py [total] ; prints one JSON line per configuration."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dqa_oracle as O
from perceptron_dqa_b200._lib import Plan
N, nxi, P, dt, chi = 21, 17, 1000, 1.0, int(os.environ.get("CHI", 32))
total = int(sys.argv[1]) if len(sys.argv) > 1 else 888
uk, fx = O.fourier_tables(N, P, dt)
MHZ = 1965.0


def mk(batch, stream, seed):
    rng = np.random.default_rng(seed)
    xi = (2 * rng.integers(0, 2, size=(batch, nxi, N)) - 1).astype(np.int8)
    pl = Plan(N, nxi, P, dt, chi, batch=batch, stream=stream)
    pl.set_patterns(xi); pl.set_fourier(uk, fx); pl.reset_plus(); pl.pp = P // 2
    return pl


def run(nsplit, stagger_ms, pad, K=3):
    if pad: os.environ["DQA_JAC_SMEM_PAD"] = str(pad)
    else: os.environ.pop("DQA_JAC_SMEM_PAD", None)
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    plans = [mk(total // nsplit, s.cuda_stream, 100 + i) for i, s in enumerate(streams)]
    for pl in plans: pl.step(2)   # saturate the bonds
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i, (pl, s) in enumerate(zip(plans, streams)):
        if stagger_ms and i:
            with torch.cuda.stream(s):
                torch.cuda._sleep(int(i * stagger_ms * 1e-3 * MHZ * 1e6))
        pl.step_async(K)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0 - (nsplit - 1) * stagger_ms * 1e-3 * 0.5) / K  # the stagger is idle time of the late plans only
    raw = (time.perf_counter() - t0) / K
    for pl in plans: pl.close()
    print(json.dumps({"splits": nsplit, "stagger_ms": stagger_ms, "jac_pad": pad, "total": total, "ms_per_step_raw": round(raw * 1e3, 1),
                      "rs_per_s_raw": round(total / raw, 1)}), flush=True)


if __name__ == "__main__":
    cfgs = [(1, 0, 0), (2, 0, 0), (2, 15, 0), (2, 30, 0), (3, 0, 0), (3, 20, 0), (3, 40, 0), (4, 15, 0),
            (2, 30, 51200), (3, 20, 51200), (3, 40, 51200), (4, 15, 51200), (1, 0, 51200)]
    for c in cfgs:
        run(*c, K=4)
