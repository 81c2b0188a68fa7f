"""Op-level parity (the reference's lib/tenn.py functions) shared by the GPU tests and the CPU emulator
tests.  Modelled on the reference's own lib/tenn-tests.py (random complex MPO of bond 10 and MPS of bond 3
on N=5 sites; checks of apply_mpsmpo, right_canonize, compress_svd_normalized(max_bd=4)), but comparing
against the oracle (quimb is not installed) and, where a gauge is free, on gauge-invariant quantities."""
import numpy as np

import dqa_oracle as O
from helpers import golden, snapshots


def rand_mps(rng, n, bond):
    shapes = [(1, 2, bond)] + [(bond, 2, bond)] * (n - 2) + [(bond, 2, 1)]
    return [rng.uniform(-1, 1, s) + 1j * rng.uniform(-1, 1, s) for s in shapes]


def rand_mpo(rng, n, bond):
    shapes = [(1, bond, 2, 2)] + [(bond, bond, 2, 2)] * (n - 2) + [(bond, 1, 2, 2)]
    return [rng.uniform(-1, 1, s) + 1j * rng.uniform(-1, 1, s) for s in shapes]


def check_apply_and_braket(T):
    rng = np.random.default_rng(1)
    mps, mpo = rand_mps(rng, 5, 3), rand_mpo(rng, 5, 10)
    got = T.apply_mpsmpo(mps, mpo)
    want = O.apply_mpsmpo(mps, mpo)
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-14
    other = rand_mps(rng, 5, 30)
    e = T.braket(got, other)
    w = O.braket(want, other)
    assert e.shape == (1, 1) and abs(e[0, 0] - w[0, 0]) < 1e-9 * max(1.0, abs(w[0, 0]))
    # open right end: environment matrix
    e = T.braket(got[:3], other[:3])
    w = O.braket(want[:3], other[:3])
    assert e.shape == w.shape and np.abs(e - w).max() < 1e-10 * np.abs(w).max()


def check_qr(T):
    rng = np.random.default_rng(2)
    for m, n in ((12, 5), (6, 6), (4, 9), (1, 3), (40, 20)):
        x = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
        q, r = T.qr_stabilized(x)
        k = min(m, n)
        assert q.shape == (m, k) and r.shape == (k, n)
        assert np.abs(q @ r - x).max() < 1e-13
        assert np.abs(q.conj().T @ q - np.eye(k)).max() < 1e-13
        d = np.diag(r)
        assert np.abs(d.imag).max() < 1e-14 and d.real.min() >= 0
        assert np.abs(np.tril(r, -1)).max() == 0
        qo, ro = O.qr_stabilized(x)
        assert np.abs(q - qo).max() < 1e-12 and np.abs(r - ro).max() < 1e-12
    # exactly rank-deficient input: no column of Q may vanish (sgn(0) := +1)
    x = np.ones((4, 3), dtype=np.complex128)
    q, r = T.qr_stabilized(x)
    assert np.abs(q @ r - x).max() < 1e-14 and np.abs(q.conj().T @ q - np.eye(3)).max() < 1e-13


def check_right_canonize(T):
    rng = np.random.default_rng(3)
    mps = O.apply_mpsmpo(rand_mps(rng, 5, 3), rand_mpo(rng, 5, 10))
    got = T.right_canonize([t.copy() for t in mps], 1)
    want = O.right_canonize([t.copy() for t in mps], 1)
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-11
    for t in got[1:]:
        m = t.reshape(t.shape[0], -1)
        assert np.abs(m @ m.conj().T - np.eye(t.shape[0])).max() < 1e-12
    assert abs(O.mps_norm2(got) - 1.0) < 1e-12


def check_svd_truncated(T):
    rng = np.random.default_rng(4)
    x = (rng.normal(size=(12, 7)) + 1j * rng.normal(size=(12, 7))) @ np.diag([1, 0.5, 0.2, 1e-2, 1e-4, 1e-7, 1e-9])
    x = x @ (np.linalg.qr(rng.normal(size=(7, 7)) + 1j * rng.normal(size=(7, 7)))[0])
    cases = [dict(cutoff=1e-10, cutoff_mode=4, max_bond=3, absorb=1, renorm=2),
             dict(cutoff=1e-10, cutoff_mode=4, max_bond=-1, absorb=1, renorm=2),
             dict(cutoff=1e-3, cutoff_mode=1, max_bond=-1, absorb=-1, renorm=0),
             dict(cutoff=1e-3, cutoff_mode=2, max_bond=-1, absorb=0, renorm=0),
             dict(cutoff=1e-6, cutoff_mode=3, max_bond=-1, absorb=None, renorm=1),
             dict(cutoff=1e-6, cutoff_mode=5, max_bond=-1, absorb=1, renorm=1),
             dict(cutoff=1e-6, cutoff_mode=6, max_bond=5, absorb=1, renorm=0),
             dict(cutoff=-1.0, cutoff_mode=3, max_bond=4, absorb=0, renorm=0),
             dict(cutoff=-1.0, cutoff_mode=3, max_bond=-1, absorb=None, renorm=0)]
    for mat in (x, x.T.copy(), x[:5, :5].copy()):
        for kw in cases:
            got = T.svd_truncated(mat, **kw)
            want = O.svd_truncated(mat, **kw)
            if kw["absorb"] is None:
                assert got[1].shape == want[1].shape and np.abs(got[1] - want[1]).max() < 1e-12, kw
                pg = got[0] @ np.diag(got[1]) @ got[2]
                pw = want[0] @ np.diag(want[1]) @ want[2]
            else:
                assert got[0].shape == want[0].shape and got[1].shape == want[1].shape, kw
                pg, pw = got[0] @ got[1], want[0] @ want[1]
            assert np.abs(pg - pw).max() < 1e-12, kw


def check_compress(T):
    rng = np.random.default_rng(5)
    mps = O.right_canonize(O.apply_mpsmpo(rand_mps(rng, 5, 3), rand_mpo(rng, 5, 10)), 1)
    got = T.compress_svd_normalized([t.copy() for t in mps], 4)
    want = O.compress_svd_normalized([t.copy() for t in mps], 4)
    assert [t.shape for t in got] == [t.shape for t in want]
    assert abs(O.fidelity_defect(got, want)) < 1e-12 and abs(O.mps_norm2(got) - 1) < 1e-12
    for t in got[:-1]:
        m = t.reshape(-1, t.shape[2])
        assert np.abs(m.conj().T @ m - np.eye(t.shape[2])).max() < 1e-12


def check_dense_substep_equals_reference(T):
    """The reference's own sub-step spelled with the op-level API (dense Fourier MPO, dense QR sweep,
    dense SVD sweep: dQA_mps.py:94-103) reproduces the reference's stored post-state."""
    from perceptron_dqa_b200.dQA_utils import PerceptronHamiltonian

    g = golden("n7_chi4")
    p, mu, pre, post = snapshots(g)[2]
    uz = PerceptronHamiltonian.make_Uz(7, g["Uk_FT"][:, p], g["xi"][mu])
    for a, b in zip(uz, O.make_Uz(7, g["Uk_FT"][:, p], g["xi"][mu])):
        assert np.abs(a - b).max() == 0
    psi = T.apply_mpsmpo(pre, uz)
    psi = T.right_canonize(psi, 1)
    psi = T.compress_svd_normalized(psi, int(g["chi"]))
    assert [t.shape for t in psi] == [t.shape for t in post]
    assert abs(O.fidelity_defect(psi, post)) < 1e-12


def check_energy_ext(T, ops):
    g = golden("n7_chi4")
    _p, _mu, _pre, post = snapshots(g)[1]
    xi, fx = g["xi"], g["fxft"]
    e = ops.energy(post, xi, fx, 0)
    assert abs(e - O.compute_loss(post, xi, fx)) < 1e-13
    # three ancilla sites appended: a product state on the ancillas keeps the energy (lib/dQA_loss.py:40-44)
    rng = np.random.default_rng(6)
    anc = rand_mps(rng, 4, 2)[1:]  # bonds (2,2),(2,2),(2,1)
    ext = [t.copy() for t in post[:-1]] + [np.concatenate([post[-1], post[-1]], axis=2)] + anc
    want = O.compute_loss(ext, xi, fx, n_ancilla=3)
    got = ops.energy(ext, xi, fx, 3)
    assert abs(got - want) < 1e-12 * max(1.0, abs(want))


def check_statevector(ops):
    """GPU/emulated state-vector dQA against the numpy state-vector oracle."""
    g = golden("n7_chi8_untruncated")
    got = ops.statevector_dqa(g["xi"], int(g["P"]), float(g["dt"]))
    want = O.statevector_dqa(g["xi"], int(g["P"]), float(g["dt"]))
    assert np.abs(got - want).max() < 1e-13
    assert np.abs(got - g["eps"].real).max() < 1e-6  # the MPS reference run, up to its 1e-10 cutoff
