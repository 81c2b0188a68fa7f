"""TEST INFRASTRUCTURE ONLY -- numpy model of the product's *sector* compression.

The CUDA hot path does not materialise the reference's bond-(chi*D) direct-sum MPS.
It uses the fact that U_z^mu = sum_k g_k Phi_k (lib/dQA_utils.py:46-59, Fourier form,
bond D=N+1) is exactly the diagonal unitary exp(-i gamma f(X_mu)), X_mu = Hamming
distance to pattern mu, and represents it with a *counter* bond: the bond index
between sites n-1 and n is (c, y) with y = number of mismatches on sites >= n.
Different y are orthogonal subspaces, so the QR right-canonisation
(lib/tenn.py:152-190) splits into independent small QRs, one per (site, y) sector,
and the left-to-right truncated SVD sweep (lib/tenn.py:229-262) sees a
(2c x sum_y q_y) matrix whose columns are grouped by sector.

This file states that algorithm in numpy, step for step as the kernels do it, so
tests can (i) check it against oracle/dqa_oracle.py on gauge-invariant quantities
and (ii) check the kernels' intermediate buffers.  It is not used by the product.
"""
from __future__ import annotations

import numpy as np

from dqa_oracle import CDT, trim_rule


def householder_qr_pos(m):
    """Reduced QR with real non-negative diagonal of R (what the sector-QR kernel
    produces); rank-deficient columns give zero rows of R, never a zero column of Q."""
    q, r = np.linalg.qr(m)
    d = np.diag(r)
    ph = np.where(np.abs(d) > 0, d / np.where(np.abs(d) > 0, np.abs(d), 1), 1)
    return q * ph.reshape(1, -1), r * np.conj(ph).reshape(-1, 1)


def right_sweep(psi, nbit):
    """Sector right-canonisation.  Returns per bond n=1..N dicts
    R[n][y] (q x b_n) and Qb[n][y] = {s: (q'_{y'} x q) block}, y' = y - nbit[n][s]."""
    n_sites = len(psi)
    R = {n_sites: {0: np.ones((1, 1), dtype=CDT)}}
    Qb = {}
    for n in range(n_sites - 1, 0, -1):
        R[n], Qb[n] = {}, {}
        a = psi[n]
        for y in range(0, n_sites - n + 1):
            blocks, keys = [], []
            for s in range(2):
                yp = y - nbit[n][s]
                if yp in R[n + 1]:
                    blocks.append(R[n + 1][yp] @ a[:, s, :].T)  # (q' x b_n)
                    keys.append(s)
            m = np.vstack(blocks)
            q, r = householder_qr_pos(m)
            R[n][y] = r
            Qb[n][y] = {}
            off = 0
            for s, blk in zip(keys, blocks):
                Qb[n][y][s] = q[off:off + blk.shape[0], :]
                off += blk.shape[0]
    return R, Qb


def sector_offsets(rdict):
    """Column offset of each sector y in the concatenated bond."""
    off, o = {}, 0
    for y in sorted(rdict):
        off[y] = o
        o += rdict[y].shape[0]
    return off, o


def theta0(psi, nbit, u, R):
    off, tot = sector_offsets(R[1])
    th = np.zeros((2, tot), dtype=CDT)
    for s in range(2):
        for yp, r in R[1].items():
            th[s, off[yp]:off[yp] + r.shape[0]] = u[yp + nbit[0][s]] * (r @ psi[0][0, s, :])
    return th


def next_theta(c, n, nbit, R, Qb):
    """Theta_n[(beta,s),(j',y')] = sum_j C[beta,(j,y)] Q_{n,y}[(s,j'),j], y = y'+nbit[n][s]."""
    off_in, _ = sector_offsets(R[n])
    off_out, tot_out = sector_offsets(R[n + 1])
    cl = c.shape[0]
    th = np.zeros((cl, 2, tot_out), dtype=CDT)
    for y, qb in Qb[n].items():
        cy = c[:, off_in[y]:off_in[y] + R[n][y].shape[0]]
        for s, blk in qb.items():
            yp = y - nbit[n][s]
            th[:, s, off_out[yp]:off_out[yp] + blk.shape[0]] = cy @ blk.T
    return th.reshape(2 * cl, tot_out)


def apply_pattern_sector(psi, u, nbit, chi, cutoff=1e-10, svd=None, stats=None, cutoff_mode=4, renorm=2):
    """One sub-step (U_z^mu, canonise, truncate) in the sector representation.

    psi   list of (b_n,2,b_{n+1}) complex128 site tensors
    u     (N+1,) complex: exp(-i gamma_p f(x))
    nbit  (N,2) int: mismatch indicator hbit[mu,n] XOR s
    Returns the truncated, left-canonical, unit-norm MPS."""
    n_sites = len(psi)
    R, Qb = right_sweep(psi, nbit)
    th = theta0(psi, nbit, u, R)
    out = []
    cl = 1
    for l in range(n_sites - 1):
        uu, s, vh = (svd or np.linalg.svd)(th, full_matrices=False)
        n_chi, scale = trim_rule(s, cutoff=cutoff, cutoff_mode=cutoff_mode, max_bond=chi, renorm=renorm)
        if stats is not None:
            stats.append((l, th.shape, s.copy(), n_chi))
        uk = uu[:, :n_chi]
        out.append(uk.reshape(cl, 2, n_chi))
        c = scale * (uk.conj().T @ th)
        th = next_theta(c, l + 1, nbit, R, Qb)
        cl = n_chi
    last = th.reshape(cl, 2, 1)
    out.append(last / np.linalg.norm(last.ravel()))
    return out


def mixer(psi, beta):
    """exp(+i beta sigma^x) on every physical leg (lib/dQA_utils.py:26-30)."""
    c, s = np.cos(beta), 1j * np.sin(beta)
    return [np.stack([c * a[:, 0, :] + s * a[:, 1, :], s * a[:, 0, :] + c * a[:, 1, :]], axis=1) for a in psi]


def uz_diagonal(uk_col, n):
    """u(x) = (1/sqrt(D)) sum_k Uk_FT[k,p] exp(2 pi i k x / D): the diagonal of U_z as the
    reference's Fourier MPO represents it (equals exp(-i gamma_p f(x)) to roundoff)."""
    d = n + 1
    k = np.arange(d)
    x = np.arange(d)
    return (uk_col[None, :] * np.exp(2j * np.pi * np.outer(x, k) / d)).sum(1) / np.sqrt(d)


# ----------------------------------------------------------------------------
# arithmetic model of the truncation kernel: Householder LQ + one-sided Jacobi
# ----------------------------------------------------------------------------


def round_robin_pairs(n):
    """The (n-1)-round tournament ordering of k_jacobi (bonds above 32; n even after padding): round r pairs position i
    with position n-1-i of the rotated ring.  The systolic kernel for chi <= 32 (k_jacobi_sys) visits the same pairs in
    the odd-even transposition order instead; both are cyclic orderings with the same fixed point, the SVD."""
    m = n + (n & 1)
    ring = list(range(m))
    rounds = []
    for _ in range(m - 1):
        pairs = []
        for i in range(m // 2):
            a, b = ring[i], ring[m - 1 - i]
            if a < n and b < n:
                pairs.append((min(a, b), max(a, b)))
        rounds.append(pairs)
        ring = [ring[0]] + [ring[-1]] + ring[1:-1]
    return rounds


def jacobi_svd_model(x, max_sweeps=30, counters=None):
    """SVD of x (m x n) the way the kernel does it: L from a Householder LQ
    (x = L Qh, L is m x k lower-trapezoidal, k = min(m,n)); one-sided Jacobi on the
    columns of L with right rotations until all pairs are orthogonal to
    sqrt(m)*eps; sigma = column norms, U = normalised columns, sorted descending;
    V^H = diag(1/sigma) U^H x.  Returns (U, s, Vh) like numpy.linalg.svd(full_matrices=False)."""
    m, n = x.shape
    k = min(m, n)
    qh, rr = np.linalg.qr(x.conj().T)  # x^H = Qh R  ->  x = R^H Qh^H
    g = rr.conj().T.copy()  # (m x k) = L
    tol = np.sqrt(m) * np.finfo(np.float64).eps
    rounds = round_robin_pairs(k)
    nrot = 0
    for sweep in range(max_sweeps):
        rotated = 0
        for pairs in rounds:
            for p, q in pairs:
                xp, xq = g[:, p], g[:, q]
                alpha = np.vdot(xp, xp).real
                beta = np.vdot(xq, xq).real
                gam = np.vdot(xp, xq)
                ag = abs(gam)
                if ag <= tol * np.sqrt(alpha * beta) or ag == 0.0:
                    continue
                rotated += 1
                zeta = (beta - alpha) / (2.0 * ag)
                t = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta))
                c = 1.0 / np.sqrt(1.0 + t * t)
                s = c * t
                ph = np.conj(gam) / ag  # e^{-i phi}
                g[:, p], g[:, q] = c * xp - (s * ph) * xq, s * xp + (c * ph) * xq
        nrot += rotated
        if rotated == 0:
            break
    if counters is not None:
        counters.append((sweep + 1, nrot))
    sig = np.linalg.norm(g, axis=0)
    order = np.argsort(-sig, kind="stable")
    sig = sig[order]
    u = g[:, order] / np.where(sig > 0, sig, 1.0)
    vh = (u.conj().T @ x) / np.where(sig > 0, sig, 1.0)[:, None]
    return u, sig, vh
