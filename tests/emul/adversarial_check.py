"""TEST INFRASTRUCTURE ONLY -- run a few parity checks on the SIMT emulator under an adversarial warp schedule
(DQA_EMUL_SEED must be set before the emulator library is loaded, hence a separate process).
usage: DQA_EMUL_SEED=3 python tests/emul/adversarial_check.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)
os.environ.setdefault("OMP_NUM_THREADS", "1")

import emul_lib  # noqa: E402
import parity_suite as ps  # noqa: E402
from perceptron_dqa_b200._lib import Plan  # noqa: E402

assert os.environ.get("DQA_EMUL_SEED"), "set DQA_EMUL_SEED"
lib = emul_lib.load()


def factory(N, n_xi, P, dt, chi, batch):
    return Plan(N, n_xi, P, dt, chi, batch=batch, lib=lib, workspace=emul_lib.numpy_workspace)


# saturated bonds, so that the column pairs of one SVD span several warps (10 pairs = 3 warps at chi = 10)
ps.check_teacher_forced(factory, "n7_chi4")
ps.check_teacher_forced(factory, "c1_n12_chi10", max_snaps=2)
print("adversarial schedule ok, seed", os.environ["DQA_EMUL_SEED"])
