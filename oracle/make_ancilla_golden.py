"""TEST INFRASTRUCTURE ONLY -- tests/golden/ancilla_loss.npz from the UNMODIFIED reference's lib/dQA_loss.py.

Run in the build container (needs /root/reference):
    OMP_NUM_THREADS=1 python oracle/make_ancilla_golden.py
`mydQA_ancilla(...).compute_loss(state)` (lib/dQA_loss.py:53-78) is called through the import shim
(oracle/ref_runner.load_ancilla) on external MPS with 0 / 3 trailing ancilla sites, with and without
flip_endian -- the way dQA_circuit.py:152-156,177-180,219 calls it.  The MPS and the reference's answers are
stored so the GPU test (tests/test_gpu_ops.py::test_mydqa_ancilla_vs_reference_fixture) can replay them where the
reference tree does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_runner  # noqa: E402
from make_golden import pack_mps  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "ancilla_loss.npz")


def rand_mps(rng, bonds):
    return [(rng.standard_normal((bonds[i], 2, bonds[i + 1])) + 1j * rng.standard_normal((bonds[i], 2, bonds[i + 1]))) / 3
            for i in range(len(bonds) - 1)]


def main():
    mod, dqa_loss = ref_runner.load_ancilla()
    rng = np.random.default_rng(11)
    n, n_xi, P, dt = 7, 5, 20, 0.7
    xi = (2 * rng.integers(0, 2, size=(n_xi, n)) - 1).astype(np.int64)
    out = {"xi": xi.astype(np.int8), "P": P, "dt": dt}
    cases = []
    for ci, (n_anc, flip) in enumerate([(0, False), (0, True), (3, False), (3, True), (1, True)]):
        bonds = [1, 2, 4, 6, 6, 4, 2] + ([1] if n_anc == 0 else [2] + [2] * (n_anc - 1) + [1])
        assert len(bonds) == n + n_anc + 1
        psi = rand_mps(rng, bonds)
        with ref_runner.quiet():
            obj = dqa_loss.mydQA_ancilla(xi, P=P, dt=dt, n_ancilla=n_anc, flip_endian=flip)
            eps = complex(obj.compute_loss(psi))
        b, flat = pack_mps(psi)
        out[f"case{ci}_bonds"], out[f"case{ci}_mps"] = b, flat
        out[f"case{ci}_n_anc"], out[f"case{ci}_flip"], out[f"case{ci}_eps"] = n_anc, flip, eps
        cases.append(ci)
        print(f"case {ci}: n_ancilla={n_anc} flip_endian={flip} eps={eps}")
    out["cases"] = np.array(cases)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
