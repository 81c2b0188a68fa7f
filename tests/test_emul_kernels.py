"""CPU: the product's .cu sources compiled with g++ on the fiber SIMT emulator
(tests/emul), checked against the reference-derived fixtures at small sizes.  This tests
kernel *logic* (indexing, barriers, shuffles); the `-m gpu` tests are the parity tests
proper.  The emulator library is never loaded by the package."""
import pytest

import parity_suite as ps


@pytest.fixture(scope="module")
def factory():
    import emul_lib
    from perceptron_dqa_b200._lib import Plan

    lib = emul_lib.load()

    def make(N, n_xi, P, dt, chi, batch):
        return Plan(N, n_xi, P, dt, chi, batch=batch, lib=lib, workspace=emul_lib.numpy_workspace)

    return make


def test_index_tensors(factory):
    ps.check_index_tensors(factory, "n7_chi4")
    ps.check_index_tensors(factory, "c1_n12_chi10")


def test_eps0(factory):
    ps.check_eps0(factory, "n7_chi4")


def test_teacher_forced_n7(factory):
    ps.check_teacher_forced(factory, "n7_chi4")
    ps.check_teacher_forced(factory, "n7_chi8_untruncated")


def test_energy_and_mixer(factory):
    ps.check_energy_on_snapshots(factory, "n7_chi4")
    ps.check_mixer(factory, "n7_chi4")


def test_free_run_n7(factory):
    ps.check_free_run(factory, "n7_chi8_untruncated", steps=3)


def test_batch(factory):
    ps.check_batch_consistency(factory, "n7_chi4", steps=1, batch=3)


def test_batch_split_into_sub_batches(factory, monkeypatch):
    # DQA_SPLIT: the step runs the batch as independent sub-batches (views of the plan's buffers: 1 + 2 and 1 + 1 + 1 instances)
    for ns in ("2", "3"):
        monkeypatch.setenv("DQA_SPLIT", ns)
        ps.check_batch_consistency(factory, "n7_chi4", steps=2, batch=3)


def test_batch_split_with_per_instance_dt(monkeypatch):
    # the views of a split plan carry the per-instance table index and dt: two instances with different dt, one per sub-batch
    import numpy as np

    import dqa_oracle as O
    import emul_lib
    from helpers import golden
    from perceptron_dqa_b200._lib import Plan

    monkeypatch.setenv("DQA_SPLIT", "2")
    kw = dict(lib=emul_lib.load(), workspace=emul_lib.numpy_workspace)
    g = golden("n7_chi4")
    xi, P, dt, chi = g["xi"], int(g["P"]), float(g["dt"]), int(g["chi"])
    n, nx = xi.shape[1], xi.shape[0]
    dts = [dt, 0.4]
    uks = np.stack([O.fourier_tables(n, P, v)[0] for v in dts])
    both = Plan(n, nx, P, dt, chi, batch=2, n_tables=2, **kw)
    both.set_patterns(np.stack([xi, xi]))
    both.set_fourier_tables(uks, g["fxft"], [0, 1], dts)
    both.reset_plus()
    both.step(2)
    monkeypatch.setenv("DQA_SPLIT", "1")
    for i, v in enumerate(dts):
        one = Plan(n, nx, P, v, chi, **kw)
        one.set_patterns(xi[None])
        one.set_fourier(uks[i], g["fxft"])
        one.reset_plus()
        one.step(2)
        assert np.abs(both.eps_trace()[i, :3] - one.eps_trace()[0, :3]).max() < 1e-13
    assert np.abs(both.eps_trace()[0, :3] - g["eps"][:3]).max() < 1e-10


def test_small_edge_cases(factory):
    ps.check_small_edge_cases(factory)


def test_bond_sweep_non_power_of_two(factory):
    ps.check_bond_sweep(factory, chis=(6, 12), n=8, n_xi=3, P=4, steps=2)


def test_bond_above_16_uses_the_six_and_eight_slot_kernels(factory):
    # chi in (16,24]: k_sector_qr<6>, k_jacobi_sys<6>; chi in (24,32]: the <8> ones; both with the paired DMMA tiles
    # (bonds 20 and 28 are reached at N=10)
    ps.check_bond_sweep(factory, chis=(20, 28), n=10, n_xi=2, P=3, steps=1)


def test_teacher_forced_n12_one_substep(factory):
    ps.check_teacher_forced(factory, "c1_n12_chi10", max_snaps=1)


# ---- op-level API (lib/tenn.py functions) on the emulator -------------------------------------------
@pytest.fixture(scope="module")
def tenn_emul():
    import emul_lib
    from perceptron_dqa_b200 import tenn
    from perceptron_dqa_b200._lib import Ops

    o = Ops(scratch_bytes=64 << 20, lib=emul_lib.load(), workspace=emul_lib.numpy_workspace)
    tenn.set_ops(o)
    yield tenn, o
    tenn.set_ops(None)


def test_ops_apply_braket_qr(tenn_emul):
    import ops_suite as os_

    T, _ = tenn_emul
    os_.check_apply_and_braket(T)
    os_.check_qr(T)


def test_ops_canonize_svd_compress(tenn_emul):
    import ops_suite as os_

    T, _ = tenn_emul
    os_.check_right_canonize(T)
    os_.check_svd_truncated(T)
    os_.check_compress(T)


def test_ops_dense_substep_and_energy(tenn_emul):
    import ops_suite as os_

    T, o = tenn_emul
    os_.check_dense_substep_equals_reference(T)
    os_.check_energy_ext(T, o)
    os_.check_statevector(o)


def test_bond_64_kernels(factory):
    # chi = 64: k_sector_qr<16> and the chi > 56 Jacobi kernel (bonds up to 43 here: all its columns are in shared memory;
    # the saturated 128-column case, where 20 columns stay in global memory, is the GPU test on n14_chi64_saturated)
    ps.check_teacher_forced(factory, "n12_chi64_untruncated", max_snaps=1)


def test_bond_48_saturated(factory):
    # chi = 48 at saturation: k_sector_qr<12>, blocked LQ with 96 rows, the 12-row-slot systolic Jacobi (one CTA per SM on the GPU)
    ps.check_teacher_forced(factory, "n12_chi48_saturated")


def test_round2_abi_on_the_emulator():
    """The round-2 entry points through the same C ABI on the CPU build: truncation-policy plan fields and the discarded-weight
    trace against the numpy sector model, per-instance dt tables against single-dt plans, status words cleared by a new state."""
    import numpy as np

    import dqa_oracle as O
    import emul_lib
    import sector_model as S
    from helpers import golden, snapshots
    from perceptron_dqa_b200._lib import DqaError, Plan

    lib = emul_lib.load()
    kw = dict(lib=lib, workspace=emul_lib.numpy_workspace)
    g = golden("n7_chi4")
    xi, P, dt, chi = g["xi"], int(g["P"]), float(g["dt"]), int(g["chi"])
    n, nx = xi.shape[1], xi.shape[0]
    hbit, _, _ = O.hamming_tables(xi)
    p, mu, pre, _post = snapshots(g)[2]
    u = S.uz_diagonal(g["Uk_FT"][:, p], n)
    nbit = np.stack([hbit[mu], 1 - hbit[mu]], axis=1)
    for mode, renorm, cutoff in ((4, 0, 1e-10), (2, 0, 5e-2), (5, 1, 1e-2)):
        pl = Plan(n, nx, P, dt, chi, cutoff=cutoff, cutoff_mode=mode, renorm=renorm, absorb=1, **kw)
        pl.set_patterns(xi[None])
        pl.set_fourier(g["Uk_FT"], g["fxft"])
        pl.set_mps(pre)
        pl.apply_pattern(p, mu)
        got = pl.get_mps()
        stats = []
        want = S.apply_pattern_sector([t.copy() for t in pre], u, nbit, chi, cutoff=cutoff, cutoff_mode=mode, renorm=renorm, stats=stats)
        assert [t.shape for t in got] == [t.shape for t in want]
        assert abs(abs(O.overlap(got, want)) / np.sqrt(O.mps_norm2(got) * O.mps_norm2(want)) - 1.0) < 1e-12
        w = pl.truncation_error()
        for l, _shape, sv, n_chi in stats:
            assert abs(w[l] - float((sv[n_chi:] ** 2).sum()) / float((sv ** 2).sum())) < 1e-13
    with pytest.raises(DqaError):
        Plan(n, nx, P, dt, chi, cutoff_mode=4, renorm=2, absorb=-1, **kw)
    # two instances with different dt in one plan == two single-dt plans
    dts = [dt, 0.4]
    uks = np.stack([O.fourier_tables(n, P, v)[0] for v in dts])
    both = Plan(n, nx, P, dt, chi, batch=2, n_tables=2, **kw)
    both.set_patterns(np.stack([xi, xi]))
    both.set_fourier_tables(uks, g["fxft"], [0, 1], dts)
    both.reset_plus()
    both.step(2)
    for i, v in enumerate(dts):
        one = Plan(n, nx, P, v, chi, **kw)
        one.set_patterns(xi[None])
        one.set_fourier(uks[i], g["fxft"])
        one.reset_plus()
        one.step(2)
        assert np.abs(both.eps_trace()[i, :3] - one.eps_trace()[0, :3]).max() < 1e-13
    assert np.abs(both.eps_trace()[0, :3] - g["eps"][:3]).max() < 1e-10
    with pytest.raises(DqaError):
        both.set_fourier(uks[0], g["fxft"])  # a two-table plan wants set_fourier_tables
    # a non-finite state raises and sets the status word; loading a state clears it
    good = both.get_mps(0)
    bad = [t.copy() for t in good]
    bad[2][...] = np.nan
    both.set_mps(bad, 0)
    with pytest.raises(DqaError):
        both.step(1)
    assert both.status()[0] != 0 and both.status()[1] == 0
    both.set_mps(good, 0)
    assert both.status()[0] == 0


@pytest.mark.parametrize("seed", [1, 2])
def test_adversarial_warp_schedule(seed):
    """Same kernels with the emulator's warps scheduled in random order and randomly held back (DQA_EMUL_SEED): hand-offs
    between warps that rely on lock-step scheduling instead of a barrier show up here as wrong results or a deadlock."""
    import os
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, DQA_EMUL_SEED=str(seed), OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(here, "emul", "adversarial_check.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
