"""TEST INFRASTRUCTURE ONLY -- 80-bit (np.clongdouble) run of the dQA sub-step algebra.

Purpose: the truncated trajectory is roundoff-chaotic, so two correct
complex128 implementations drift apart exponentially.  To say how far a *correct* complex128
run may be from the reference's stored curve we need the exact-arithmetic trajectory.  This
file evaluates the same mathematical map (U_z, canonical form, optimal truncation per bond
with the reference's trim rule, U_x, energy) with 64-bit mantissas (eps = 1.1e-19) using
hand-rolled Householder and one-sided Jacobi (LAPACK has no long double), in the sector
representation of oracle/sector_model.py.  Its own roundoff is ~2000x smaller than any
complex128 run's, so for the first ~20 steps it serves as ground truth:
    |eps_reference(s) - eps_truth(s)|  = the reference's true roundoff envelope.
"""
import numpy as np

import dqa_oracle as O

LD = np.clongdouble
RD = np.longdouble


def householder_qr(m):
    """reduced QR, R diagonal real >= 0, in the dtype of m (unblocked Householder)."""
    a = np.array(m, dtype=LD)
    rows, cols = a.shape
    k = min(rows, cols)
    vs, taus = [], []
    for j in range(k):
        x = a[j:, j].copy()
        alpha = x[0]
        ss = (np.abs(x[1:]) ** 2).sum()
        if ss == 0 and alpha.imag == 0:
            vs.append(None)
            taus.append(LD(0))
            continue
        nrm = np.sqrt(np.abs(alpha) ** 2 + ss)
        beta = -nrm if alpha.real >= 0 else nrm
        tau = LD((beta - alpha.real) / beta - 1j * (alpha.imag / beta))
        v = x / (alpha - beta)
        v[0] = 1
        w = np.conj(v) @ a[j:, j:]
        a[j:, j:] -= np.conj(tau) * np.outer(v, w)
        vs.append(v)
        taus.append(tau)
    r = np.triu(a[:k, :])
    q = np.zeros((rows, k), dtype=LD)
    q[:k, :k] = np.eye(k, dtype=LD)
    for j in range(k - 1, -1, -1):
        if vs[j] is None:
            continue
        w = np.conj(vs[j]) @ q[j:, :]
        q[j:, :] -= taus[j] * np.outer(vs[j], w)
    d = np.diag(r).real
    sg = np.where(d < 0, RD(-1), RD(1))
    return q * sg[None, :], r * sg[:, None]


def jacobi_svd(x, max_sweeps=60):
    """(U, s, Vh) with own LQ + one-sided Jacobi, long double."""
    x = np.array(x, dtype=LD)
    m, n = x.shape
    qh, rr = householder_qr(x.conj().T)
    g = rr.conj().T.copy()
    k = g.shape[1]
    tol = RD(np.sqrt(m)) * np.finfo(RD).eps
    for _ in range(max_sweeps):
        rot = 0
        for p in range(k - 1):
            for q in range(p + 1, k):
                xp, xq = g[:, p].copy(), g[:, q].copy()
                al = (np.abs(xp) ** 2).sum()
                be = (np.abs(xq) ** 2).sum()
                gam = np.conj(xp) @ xq
                ag = np.abs(gam)
                if ag == 0 or ag <= tol * np.sqrt(al * be):
                    continue
                rot += 1
                zeta = (be - al) / (2 * ag)
                t = (1 if zeta >= 0 else -1) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
                c = 1 / np.sqrt(1 + t * t)
                s = c * t
                ph = np.conj(gam) / ag
                g[:, p] = c * xp - (s * ph) * xq
                g[:, q] = s * xp + (c * ph) * xq
        if rot == 0:
            break
    sig = np.sqrt((np.abs(g) ** 2).sum(0))
    order = np.argsort(-sig, kind="stable")
    sig = sig[order]
    safe = np.where(sig > 0, sig, 1)
    u = g[:, order] / safe
    vh = (u.conj().T @ x) / safe[:, None]
    return u, sig, vh


def run_truth(xi, P, dt, chi, steps):
    """eps(s), s = 0..steps, in 80-bit arithmetic (returned as complex128)."""
    import sector_model as S

    xi = np.asarray(xi)
    n = xi.shape[1]
    hbit, _, _ = O.hamming_tables(xi)
    fx = O.f_perceptron(np.arange(n + 1), n).astype(RD)
    fxft = np.fft.fft(np.asarray(fx, dtype=np.float64), norm="ortho")
    # exact f(x) weights instead of the double-precision Fourier route for the energy
    psi = [np.full((1, 2, 1), RD(0.5) ** RD(0.5), dtype=LD) for _ in range(n)]
    S.householder_qr_pos = householder_qr  # noqa: the sector sweep then runs in long double
    out = [energy_exact(psi, xi, fx)]
    for p in range(steps):
        gamma = RD(p + 1) / RD(P) * RD(dt)
        u = np.exp(-1j * (gamma * fx).astype(RD)).astype(LD)
        for mu in range(xi.shape[0]):
            nbit = np.stack([hbit[mu] ^ 0, hbit[mu] ^ 1], axis=1)
            psi = S.apply_pattern_sector(psi, u, nbit, chi, svd=lambda x, full_matrices=False: jacobi_svd(x))
        beta = (RD(1) - RD(p + 1) / RD(P)) * RD(dt)
        psi = S.mixer(psi, beta)
        out.append(energy_exact(psi, xi, fx))
    del fxft
    return np.array(out, dtype=np.complex128)


def energy_exact(psi, xi, fx):
    """eps = (1/N) sum_mu <psi| f(X_mu) |psi> via the Hamming-count environment (no Fourier sum)."""
    n = xi.shape[1]
    tot = LD(0)
    for mu in range(xi.shape[0]):
        env = {0: np.ones((1, 1), dtype=LD)}
        for i, a in enumerate(psi):
            new = {}
            for x, e in env.items():
                for s in range(2):
                    mism = int((1 - int(xi[mu, i]) * (1 if s == 0 else -1)) // 2)
                    t = a[:, s, :].T @ e @ np.conj(a[:, s, :])
                    new[x + mism] = new.get(x + mism, 0) + t
            env = new
        tot += sum(fx[x] * e[0, 0] for x, e in env.items())
    return tot / n
