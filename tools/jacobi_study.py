"""CPU study (numpy): how many sweeps / rounds do one-sided Jacobi variants need on the L factors the
truncation sweep really sees (N=21, N_xi=17, chi=32, saturated bonds)?  Decides which SVD kernel design is
worth writing; results quoted in DESIGN.md.  Test infrastructure (imports oracle/).

usage: python tools/jacobi_study.py [--chi 32] [--collect 40]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dqa_oracle as O  # noqa: E402
import sector_model as S  # noqa: E402

EPS = np.finfo(np.float64).eps


def collect(chi, n=21, nxi=17, P=1000, dt=1.0, warm_patterns=6, want=40, seed=20230317):
    rng = np.random.default_rng(seed)
    xi = (2 * rng.integers(0, 2, size=(nxi, n)) - 1).astype(np.int8)
    hbit, _, _ = O.hamming_tables(xi)
    uk, fx = O.fourier_tables(n, P, dt)
    p = P // 2
    u = S.uz_diagonal(uk[:, p], n)
    psi = O.plus_state(n)
    mats = []
    mu = 0
    while len(mats) < want:
        nbit = np.stack([hbit[mu % nxi], 1 - hbit[mu % nxi]], axis=1)
        stats = []
        captured = []

        def svd(th, full_matrices=False):
            captured.append(th.copy())
            return np.linalg.svd(th, full_matrices=False)

        psi = S.apply_pattern_sector(psi, u, nbit, chi, svd=svd, stats=stats)
        if mu >= warm_patterns:
            for th in captured:
                if th.shape[0] == 2 * chi and th.shape[1] >= 2 * chi:
                    mats.append(th)
        mu += 1
        if mu % nxi == 0:
            psi = S.mixer(psi, (1 - (p + 1) / P) * dt)
    return mats[:want]


def lq_l(th):
    _, rr = np.linalg.qr(th.conj().T)
    return rr.conj().T.copy()


def oddeven_rounds(k):
    """pairs per round of the systolic (odd-even transposition) ordering on k columns"""
    pos = list(range(k))
    rounds = []
    for r in range(k):
        start = r & 1
        pairs = [(pos[i], pos[i + 1]) for i in range(start, k - 1, 2)]
        rounds.append(pairs)
        for i in range(start, k - 1, 2):
            pos[i], pos[i + 1] = pos[i + 1], pos[i]
    return rounds


def rot_round(g, pairs, tol, quad2, nrm2=None):
    """apply one round of disjoint rotations, vectorised. returns (#rotated, any above quad)"""
    p = np.array([a for a, _ in pairs])
    q = np.array([b for _, b in pairs])
    xp, xq = g[:, p], g[:, q]
    a = np.einsum('ij,ij->j', xp.conj(), xp).real
    b = np.einsum('ij,ij->j', xq.conj(), xq).real
    gam = np.einsum('ij,ij->j', xp.conj(), xq)
    ag2 = (gam * gam.conj()).real
    act = (ag2 > tol * tol * a * b) & (ag2 > 0)
    if not act.any():
        return 0, False
    ag = np.sqrt(np.where(act, ag2, 1.0))
    zeta = (b - a) / (2 * ag)
    t = np.where(zeta >= 0, 1.0, -1.0) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
    c = 1 / np.sqrt(1 + t * t)
    s = c * t
    ph = np.conj(gam) / ag
    c = np.where(act, c, 1.0)
    s = np.where(act, s, 0.0)
    ph = np.where(act, ph, 1.0)
    g[:, p] = c * xp - (s * ph) * xq
    g[:, q] = s * xp + (c * ph) * xq
    big = bool((ag2[act] > quad2 * a[act] * b[act]).any())
    return int(act.sum()), big


def scalar_jacobi(L, rounds, max_sweeps=30, quad=1e-10):
    g = L.copy()
    m = g.shape[0]
    tol = np.sqrt(m) * EPS
    nrot = 0
    active_rounds = 0
    for sweep in range(max_sweeps):
        rot = 0
        big = False
        for pairs in rounds:
            r, b = rot_round(g, pairs, tol, quad * quad)
            rot += r
            big |= b
            active_rounds += r > 0
        nrot += rot
        if rot == 0 or not big:
            break
    return g, sweep + 1, nrot, active_rounds


def block_jacobi(L, bs=8, max_sweeps=30, inner='eigh', quad=1e-10):
    """one-sided block Jacobi: for each block pair, diagonalise the 2bs x 2bs Gram exactly (eigh) and apply"""
    g = L.copy()
    m, k = g.shape
    nb = k // bs
    tol = np.sqrt(m) * EPS
    blocks = [np.arange(i * bs, (i + 1) * bs) for i in range(nb)]
    # round robin over blocks
    ring = list(range(nb))
    rounds = []
    for _ in range(nb - 1):
        rounds.append([(min(ring[i], ring[nb - 1 - i]), max(ring[i], ring[nb - 1 - i])) for i in range(nb // 2)])
        ring = [ring[0]] + [ring[-1]] + ring[1:-1]
    visits = 0
    for sweep in range(max_sweeps):
        worst = 0.0
        for pairs in rounds:
            for bi, bj in pairs:
                idx = np.concatenate([blocks[bi], blocks[bj]])
                x = g[:, idx]
                gram = x.conj().T @ x
                d = np.sqrt(np.abs(np.diag(gram)))
                rel = np.abs(gram) / np.maximum(np.outer(d, d), 1e-300)
                np.fill_diagonal(rel, 0)
                w = rel.max()
                worst = max(worst, w)
                if w <= tol:
                    continue
                visits += 1
                ev, W = np.linalg.eigh(gram)
                W = W[:, ::-1]
                g[:, idx] = x @ W
        if worst <= quad:
            break
    return g, sweep + 1, visits


def check(L, g):
    s_ref = np.linalg.svd(L, compute_uv=False)
    s = np.sort(np.linalg.norm(g, axis=0))[::-1]
    return np.max(np.abs(s - s_ref))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chi", type=int, default=32)
    ap.add_argument("--collect", type=int, default=24)
    a = ap.parse_args()
    t0 = time.time()
    mats = collect(a.chi, want=a.collect)
    print(f"collected {len(mats)} Theta matrices {mats[0].shape}..{mats[-1].shape} in {time.time()-t0:.1f}s")
    Ls = [lq_l(th) for th in mats]
    k = Ls[0].shape[1]
    rr = S.round_robin_pairs(k)
    oe = oddeven_rounds(k)
    res = {}

    def run(name, fn):
        out = [fn(L) for L in Ls]
        res[name] = out
        sw = np.mean([o[1] for o in out])
        extra = np.mean([o[2] for o in out])
        err = max(check(L, o[0]) for L, o in zip(Ls, out))
        more = f" active_rounds {np.mean([o[3] for o in out]):.0f}" if len(out[0]) > 3 else ""
        print(f"{name:34s} sweeps {sw:5.2f}  rot/visits {extra:8.1f}{more}  max|ds| {err:.1e}")

    run("scalar round-robin on L", lambda L: scalar_jacobi(L, rr))
    run("scalar odd-even on L", lambda L: scalar_jacobi(L, oe))
    run("scalar odd-even on L, no shortcut", lambda L: scalar_jacobi(L, oe, quad=0.0))
    run("scalar odd-even on L^H", lambda L: scalar_jacobi(L.conj().T.copy(), oe))

    def sorted_cols(L):
        o = np.argsort(-np.linalg.norm(L, axis=0))
        return scalar_jacobi(L[:, o], oe)
    run("odd-even on L, cols sorted by norm", sorted_cols)

    def rev_cols(L):
        return scalar_jacobi(L[:, ::-1].copy(), oe)
    run("odd-even on L, cols reversed", rev_cols)

    def second_lq(L):
        # one more LR step: L = Q2 R2 -> work on R2^H?  (L^H = Q R -> L = R^H Q^H; Jacobi on columns of R^H ... )
        q, r = np.linalg.qr(L)
        return scalar_jacobi(r.conj().T.copy(), oe)
    run("odd-even on (R of QR(L))^H", second_lq)
    run("block Jacobi bs=8 (eigh inner)", lambda L: block_jacobi(L, 8))
    run("block Jacobi bs=4 (eigh inner)", lambda L: block_jacobi(L, 4))
    run("block Jacobi bs=16 (eigh inner)", lambda L: block_jacobi(L, 16))
    # spectrum info
    s = np.linalg.svd(Ls[len(Ls) // 2], compute_uv=False)
    print("spectrum of a mid-chain L:", " ".join(f"{v:.1e}" for v in s[::4]))


if __name__ == "__main__":
    main()


def extra():
    import scipy.linalg as sl
    mats = collect(32, want=16)
    k = 64
    oe = oddeven_rounds(k)

    def rep(name, fn):
        out = [fn(th) for th in mats]
        print(f"{name:44s} sweeps {np.mean([o[1] for o in out]):5.2f} rot {np.mean([o[2] for o in out]):8.1f} active_rounds {np.mean([o[3] for o in out]):.0f}")

    def piv(th):
        q, r, p = sl.qr(th.conj().T, mode='economic', pivoting=True)
        return scalar_jacobi(r.conj().T.copy(), oe)
    rep("QRCP(Theta^H) -> Jacobi on R^H", piv)

    def piv_rows(th):
        q, r, p = sl.qr(th.conj().T, mode='economic', pivoting=True)
        return scalar_jacobi(r.copy(), oe)
    rep("QRCP(Theta^H) -> Jacobi on R", piv_rows)

    def piv2(th):
        q, r, p = sl.qr(th.conj().T, mode='economic', pivoting=True)
        q1, r1 = np.linalg.qr(r.conj().T)
        return scalar_jacobi(r1.conj().T.copy(), oe)
    rep("QRCP, then QR(R^H) -> Jacobi on R1^H", piv2)

    def nopiv_L_piv(th):
        L = lq_l(th)
        q, r, p = sl.qr(L, mode='economic', pivoting=True)   # L P = Q R -> Jacobi on R^H
        return scalar_jacobi(r.conj().T.copy(), oe)
    rep("LQ, then QRCP(L) -> Jacobi on R^H", nopiv_L_piv)

    def nopiv_LH_piv(th):
        L = lq_l(th)
        q, r, p = sl.qr(L.conj().T, mode='economic', pivoting=True)
        return scalar_jacobi(r.conj().T.copy(), oe)
    rep("LQ, then QRCP(L^H) -> Jacobi on R^H", nopiv_LH_piv)

    def gram_pre(th):
        # eigen-preconditioner: V0 from eigh(L^H L) in double (cheap model of a low-precision pre-solve), then Jacobi on L V0
        L = lq_l(th)
        w, v = np.linalg.eigh((L.conj().T @ L).astype(np.complex64).astype(np.complex128))
        v, _ = np.linalg.qr(v)
        return scalar_jacobi(L @ v[:, ::-1], oe)
    rep("L V0 (V0 = eigvecs of c64 Gram, re-orth)", gram_pre)


if __name__ == "__main__" and os.environ.get("EXTRA"):
    extra()


def bipartite_rounds(k):
    """Recursive bipartite ordering: level 0 pairs every column of the lower half with every column of the upper half
    (k/2 rounds, one half circulating past the other), then both halves do the same inside themselves in parallel, and
    so on: k/2 + k/4 + ... + 1 = k-1 rounds, k/2 disjoint pairs in every round.  (One column of every pair stays where it
    is for a whole level: the ordering of a systolic array that keeps one column of the pair in registers and lets the
    other circulate through shared memory by index arithmetic.)"""
    rounds = []
    size = k
    while size >= 2:
        h = size // 2
        for r in range(h):
            pairs = []
            for base in range(0, k, size):
                for i in range(h):
                    pairs.append((base + i, base + h + (i + r) % h))
            rounds.append(pairs)
        size = h
    return rounds


def ordering_study():
    mats = collect(32, want=16)
    Ls = [lq_l(th) for th in mats]
    k = 64
    for name, rounds in (("odd-even (systolic kernel)", oddeven_rounds(k)), ("round robin", S.round_robin_pairs(k)),
                         ("recursive bipartite", bipartite_rounds(k))):
        seen = set()
        for pr in rounds:
            flat = [x for p in pr for x in p]
            assert len(flat) == len(set(flat))
            seen |= {tuple(sorted(p)) for p in pr}
        out = [scalar_jacobi(L, rounds) for L in Ls]
        print(f"{name:28s} rounds/sweep {len(rounds):3d} pairs {len(seen):5d}  sweeps {np.mean([o[1] for o in out]):5.2f} "
              f"rot {np.mean([o[2] for o in out]):8.1f}  max|ds| {max(check(L, o[0]) for L, o in zip(Ls, out)):.1e}")


if __name__ == "__main__" and os.environ.get("ORDERING"):
    ordering_study()


def panel_pivot_lq(th, pb=8):
    """L of an LQ whose next panel is always the pb rows of largest residual norm (the k_lq variant measured in round 2):
    returns the row-permuted, lower-triangular factor."""
    m = th.shape[0]
    R = th.copy()  # residual rows: orthogonal complement of the directions chosen so far
    rem, order, qrows = list(range(m)), [], []
    while rem:
        nr = np.linalg.norm(R[rem], axis=1)
        for j in [rem[i] for i in np.argsort(-nr, kind="stable")[:pb]]:
            q = R[j] / max(np.linalg.norm(R[j]), 1e-300)
            qrows.append(q)
            rem.remove(j)
            order.append(j)
            for r in rem:
                R[r] = R[r] - (R[r] @ q.conj()) * q
    return th[order] @ np.array(qrows).conj().T


def pivot_study():
    """PIVOT=1: sweeps of the kernel's Jacobi on L from the natural-order LQ, from rows sorted once, from panel-level pivoting
    and from full row pivoting, on mid-chain factors of a warm chi=32 run."""
    import scipy.linalg as sl
    mats = collect(32, warm_patterns=40, want=24)
    oe = oddeven_rounds(64)

    def rep(name, fn):
        out = [scalar_jacobi(fn(th), oe) for th in mats]
        print(f"{name:52s} sweeps {np.mean([o[1] for o in out]):5.2f} rot {np.mean([o[2] for o in out]):8.1f} active_rounds {np.mean([o[3] for o in out]):.0f}")

    rep("LQ in the natural row order (k_lq)", lq_l)
    rep("rows sorted once by norm, then LQ", lambda th: lq_l(th[np.argsort(-np.linalg.norm(th, axis=1))]))
    rep("panel-level pivoting, panels of 8", lambda th: panel_pivot_lq(th, 8))
    rep("panel-level pivoting, panels of 16", lambda th: panel_pivot_lq(th, 16))
    rep("full row pivoting (QRCP of Theta^H)", lambda th: sl.qr(th.conj().T, mode='economic', pivoting=True)[1].conj().T.copy())


if __name__ == "__main__" and os.environ.get("PIVOT"):
    pivot_study()
