"""TEST INFRASTRUCTURE ONLY -- tests/golden/c4_subset16.npz: SURVEY App. C P4, "ensemble mean of C4 compared with the
CPU oracle on a 16-realisation subset".

Run in the build container (needs /root/reference; ~3 min on 8 cores):
    OMP_NUM_THREADS=1 python oracle/make_c4_golden.py
The first 16 realisations of the C4 table (default_rng(20230317), iid +-1, N=21, N_xi=17) are run by the
UNMODIFIED reference (dQA_mps.py through the import shim, complex128, sgn(0):=+1) with P=100, dt=1.0, chi=10 for
the first STEPS steps -- the pre-chaotic part of the trajectory, where two correct implementations still agree to
1e-10 (5.7e-13 at step 11).  Stored: the patterns, eps_r(s) of every realisation, their mean.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "c4_subset16.npz")
N, NXI, P, DT, CHI, STEPS, R = 21, 17, 100, 1.0, 10, 10, 16


def c4_table(rows=1024):
    rng = np.random.default_rng(20230317)
    return (2 * rng.integers(0, 2, size=(rows, NXI, N)) - 1).astype(np.int8)


def work(xi):
    import ref_runner

    mod = ref_runner.load(sgn0fix=True)
    obj = mod.mydQA(xi.astype(np.float64), P=P, dt=DT, max_bond=CHI)
    obj.init_fourier()
    with ref_runner.quiet():
        obj.psi = [np.array([[[2 ** -0.5], [2 ** -0.5]]], dtype=np.complex128)] * N
        obj.pp = 0
        obj.loss = [(0, obj.compute_loss(obj.psi))]
        obj.bd_monitor = []
        for _ in range(STEPS):
            obj.single_step()
    return np.array([v for _s, v in obj.loss])


def main():
    xi = c4_table()[:R]
    with mp.get_context("spawn").Pool(min(R, os.cpu_count() or 1)) as pool:
        eps = np.array(pool.map(work, list(xi)))
    np.savez_compressed(OUT, xi=xi, eps=eps, eps_mean=eps.real.mean(0), P=P, dt=DT, chi=CHI, steps=STEPS)
    print("wrote", OUT, "mean eps(s):", np.array2string(eps.real.mean(0), precision=6))


if __name__ == "__main__":
    main()
