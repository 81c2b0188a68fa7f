"""CPU: pin the oracle (oracle/dqa_oracle.py, the numpy restatement that travels) against
outputs of the UNMODIFIED reference stored in tests/golden (made by oracle/make_golden.py),
against the reference's known answers, and -- when /root/reference is present -- against
the reference itself run live behind the import shim."""
import numpy as np
import pytest

import dqa_oracle as O
import sector_model as S
from helpers import compare_states, golden, snapshots

CONFIGS = ["n7_chi4", "n7_chi8_untruncated", "c1_n12_chi10", "n15_chi10", "c2_n21_chi10"]
BIG = ["c3_n21_chi32", "n12_chi64_untruncated", "n32_chi8_synthetic"]  # tables only on the CPU (a reference sub-step takes ~10 s here)


@pytest.mark.parametrize("name", CONFIGS + BIG)
def test_tables_match_reference(name):
    g = golden(name)
    xi = g["xi"]
    uk, fx = O.fourier_tables(xi.shape[1], int(g["P"]), float(g["dt"]))
    assert np.abs(uk - g["Uk_FT"]).max() == 0.0 and np.abs(fx - g["fxft"]).max() == 0.0
    hbit, e, q = O.hamming_tables(xi)
    assert np.array_equal(hbit, g["hbit"]) and np.array_equal(e, g["e"]) and np.array_equal(q, g["q"])
    assert set(np.unique(e)) <= {0, 2} and q.max() <= 2 * xi.shape[1]
    # the phase the reference exponentiates (lib/dQA_utils.py:106) is exp(i pi q/(N+1))
    n = xi.shape[1]
    k, mu, i, s = 3, 0, 1, 1
    ref = np.exp(1.0j * (np.pi / (n + 1)) * k * (1 - float(xi[mu, i]) * (-1) ** s))
    assert abs(ref - np.exp(1j * np.pi * q[mu, i, s, k] / (n + 1))) < 1e-15


def test_known_answers():
    # quimb-dqa/quimb-dqa-demo.ipynb cell 21 and the closed forms of SURVEY section 4
    assert abs(O.eps0_closed_form(12, 9) - 0.2930447289172962) < 1e-14
    assert abs(O.eps0_closed_form(21, 17) - 0.3268194661061751) < 1e-14
    assert abs(O.eps0_closed_form(64, 51) - 0.3166677775716445) < 1e-14
    g = golden("c1_n12_chi10")
    assert abs(g["eps"][0] - 0.2930447289172962) < 1e-14
    psi = O.plus_state(12)
    assert abs(O.compute_loss(psi, g["xi"], g["fxft"]) - 0.2930447289172962) < 1e-14
    assert abs(O.compute_loss_fast(psi, g["xi"], g["fxft"]) - 0.2930447289172962) < 1e-14


@pytest.mark.parametrize("name,steps", [("n7_chi4", 8), ("n7_chi8_untruncated", 20), ("c1_n12_chi10", 6)])
def test_oracle_reproduces_reference_trajectory(name, steps):
    g = golden(name)
    o = O.OracleDQA(g["xi"], int(g["P"]), float(g["dt"]), int(g["chi"]))
    o.init_fourier()
    o.run(steps=steps)
    got = np.array([l[1] for l in o.loss])
    assert np.abs(got - g["eps"][: steps + 1]).max() < 1e-11
    assert o.bd_monitor == g["bd_monitor"][: len(o.bd_monitor)].tolist()


@pytest.mark.parametrize("name", CONFIGS + ["n12_chi48_saturated", "n14_chi64_saturated"])
def test_oracle_substeps_teacher_forced(name):
    g = golden(name)
    o = O.OracleDQA(g["xi"], int(g["P"]), float(g["dt"]), int(g["chi"]), fast_loss=True)
    o.init_fourier()
    for p, mu, pre, post in snapshots(g)[:4]:
        o.pp = p
        got = o.apply_pattern([t.copy() for t in pre], mu)
        compare_states(got, post, g["xi"], g["fxft"])


@pytest.mark.parametrize("name", ["n7_chi4", "c1_n12_chi10", "c2_n21_chi10", "n32_chi8_synthetic"])
def test_sector_model_equals_reference_substep(name):
    """The product's algorithm (counter-bond sectors + Householder LQ + one-sided Jacobi),
    stated in numpy, reproduces the reference's sub-step on gauge-invariant quantities."""
    g = golden(name)
    xi = g["xi"]
    n = xi.shape[1]
    hbit, _, _ = O.hamming_tables(xi)
    for p, mu, pre, post in snapshots(g)[:2]:
        u = S.uz_diagonal(g["Uk_FT"][:, p], n)
        fx = O.f_perceptron(np.arange(n + 1), n)
        assert np.abs(u - np.exp(-1j * (p + 1) / int(g["P"]) * float(g["dt"]) * fx)).max() < 1e-13
        nbit = np.stack([hbit[mu] ^ 0, hbit[mu] ^ 1], axis=1)
        got = S.apply_pattern_sector(pre, u, nbit, int(g["chi"]))
        compare_states(got, post, xi, g["fxft"])
        if n <= 12:
            got = S.apply_pattern_sector(pre, u, nbit, int(g["chi"]), svd=lambda x, full_matrices=False: S.jacobi_svd_model(x))
            compare_states(got, post, xi, g["fxft"])


def test_statevector_oracle_agrees_untruncated():
    g = golden("n7_chi8_untruncated")
    sv = O.statevector_dqa(g["xi"], int(g["P"]), float(g["dt"]))
    assert np.abs(sv - g["eps"].real).max() < 1e-6  # floor set by the 1e-10 cutoff
    o = O.OracleDQA(g["xi"], int(g["P"]), float(g["dt"]), 64, fast_loss=True)
    o.init_fourier()
    o.run(steps=6)
    assert np.abs(sv[:7] - np.array([l[1] for l in o.loss]).real).max() < 1e-7


def test_loss_variants_agree():
    g = golden("n7_chi4")
    _p, _mu, _pre, post = snapshots(g)[0]
    a = O.compute_loss(post, g["xi"], g["fxft"])
    b = O.compute_loss_fast(post, g["xi"], g["fxft"])
    assert abs(a - b) < 1e-14


def test_trim_rule_edge_cases():
    s = np.array([1.0, 1e-3, 1e-6, 1e-9])
    assert O.trim_rule(s, cutoff=1e-10, cutoff_mode=4, max_bond=10, renorm=2)[0] == 2
    assert O.trim_rule(s, cutoff=1e-10, cutoff_mode=4, max_bond=1, renorm=2)[0] == 1
    n, sc = O.trim_rule(np.array([1.0]), cutoff=1e-10, cutoff_mode=4, max_bond=10, renorm=2)
    assert n == 1 and sc == 1.0
    n, sc = O.trim_rule(np.array([0.6, 0.8]), cutoff=1e-10, cutoff_mode=4, max_bond=1, renorm=2)
    assert n == 1 and abs(sc - 1.0 / 0.6) < 1e-15


def test_qr_stabilized_sign_zero_quirk():
    """with the reference's sign(0)=0 an exactly-zero pivot deletes a column."""
    x = np.array([[1.0, 2.0], [0.0, 0.0]], dtype=np.complex128).T.copy()  # rank 1, R[1,1] == 0
    x = np.array([[1.0, 1.0], [1.0, 1.0]], dtype=np.complex128)
    q0, r0 = O.qr_stabilized(x, sgn0fix=False)
    q1, r1 = O.qr_stabilized(x, sgn0fix=True)
    assert np.abs(q1 @ r1 - x).max() < 1e-15
    assert np.abs(q1.conj().T @ q1 - np.eye(2)).max() < 1e-15
    if np.diag(np.linalg.qr(x)[1])[1] == 0:
        assert np.abs(q0[:, 1]).max() == 0.0


def test_live_reference_if_present():
    import ref_runner

    if not ref_runner.available():
        pytest.skip("reference tree not present on this box")
    mod = ref_runner.load(sgn0fix=True)
    g = golden("n7_chi4")
    obj = mod.mydQA(g["xi"], P=int(g["P"]), dt=float(g["dt"]), max_bond=int(g["chi"]))
    obj.init_fourier()
    with ref_runner.quiet():
        obj.run(print_info=False)
    assert np.abs(np.array([l[1] for l in obj.loss]) - g["eps"]).max() < 1e-13
    o = O.OracleDQA(g["xi"], int(g["P"]), float(g["dt"]), int(g["chi"]))
    o.init_fourier()
    o.run(steps=5)
    assert np.abs(np.array([l[1] for l in o.loss]) - np.array([l[1] for l in obj.loss])[:6]).max() < 1e-12
    # op-level: the reference's own functions against the restatement
    rng = np.random.default_rng(0)
    x = rng.normal(size=(12, 5)) + 1j * rng.normal(size=(12, 5))
    qa, ra = mod.qr_stabilized(x)
    qb, rb = O.qr_stabilized(x)
    assert np.abs(qa - qb).max() < 1e-14 and np.abs(ra - rb).max() < 1e-14
    ua, va = mod.svd_truncated(x, cutoff=1e-10, absorb=1, max_bond=3, cutoff_mode=4, renorm=2)
    ub, vb = O.svd_truncated(x, cutoff=1e-10, absorb=1, max_bond=3, cutoff_mode=4, renorm=2)
    assert np.abs(ua - ub).max() < 1e-14 and np.abs(va - vb).max() < 1e-14


def test_ancilla_loss_port_matches_reference_fixture_and_live():
    """a11: the port's compute_loss(n_ancilla) against answers of the unmodified lib/dQA_loss.py
    (tests/golden/ancilla_loss.npz) and, where the reference tree exists, against a live call."""
    import ref_runner
    from helpers import unpack_mps

    g = golden("ancilla_loss")
    live = ref_runner.load_ancilla()[1] if ref_runner.available() else None
    fx = O.fourier_tables(g["xi"].shape[1], int(g["P"]), float(g["dt"]))[1]
    for ci in g["cases"]:
        psi = unpack_mps(g[f"case{ci}_bonds"], g[f"case{ci}_mps"])
        xi = g["xi"][:, ::-1] if bool(g[f"case{ci}_flip"]) else g["xi"]
        want = complex(g[f"case{ci}_eps"])
        got = O.compute_loss(psi, xi, fx, n_ancilla=int(g[f"case{ci}_n_anc"]))
        assert abs(got - want) < 1e-12 * max(1.0, abs(want))
        if live is not None:
            with ref_runner.quiet():
                obj = live.mydQA_ancilla(g["xi"].astype(np.int64), P=int(g["P"]), dt=float(g["dt"]),
                                         n_ancilla=int(g[f"case{ci}_n_anc"]), flip_endian=bool(g[f"case{ci}_flip"]))
                assert abs(complex(obj.compute_loss(psi)) - want) < 1e-13 * max(1.0, abs(want))
