"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, complex128) of the dQA MPS step.

This file is the *oracle*: a plain-numpy restatement of the algorithm of
baronefr/perceptron-dqa's jax/numpy MPS path, written from the specification in
SURVEY.md App. A, each function citing the reference file:line it follows.  It is
not the product and no product module imports it; only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use it.

Parity status: PINNED.  tests/test_oracle_vs_reference.py runs the reference's own
unmodified files (oracle/ref_runner.py, import shim) next to this restatement in
the build container, and tests/golden/*.npz holds outputs of the unmodified
reference (made by oracle/make_golden.py) that the restatement must reproduce
on any box.  Known-answer pins: eps(0) closed form and the notebook value
0.2930447289172962 (quimb-dqa/quimb-dqa-demo.ipynb cell 21).

Conscious deviation: `sgn0fix=True` (default) uses sign(0):=+1
in qr_stabilized; `sgn0fix=False` reproduces the reference's jnp.sign verbatim.

Layouts are the reference's: MPS site (l, phys, r); MPO site (l, r, up, down).
"""
from __future__ import annotations

import math

import numpy as np

CDT = np.complex128

# ----------------------------------------------------------------------------
# a2 -- integer Hamming exponents (implicit in lib/dQA_utils.py:50-53 and :106)
# ----------------------------------------------------------------------------


def check_patterns(xi):
    """Cast a pattern array (int64/float64/float32 in the reference's data files,
    SURVEY App. B Q11) to int8 after checking every entry is +-1."""
    xi = np.asarray(xi)
    if xi.ndim != 2:
        raise ValueError("patterns must be (N_xi, N)")
    if not np.all((xi == 1) | (xi == -1)):
        raise ValueError("patterns must be +-1")
    return xi.astype(np.int8)


def hamming_tables(xi):
    """Integer index tensors of the pattern-to-MPO map.

    hbit[mu,i]    = (1 - xi[mu,i]) / 2                         in {0,1}
    e[mu,i,s]     = 1 - xi[mu,i]*(-1)**s = 2*(hbit XOR s)       in {0,2}
                    (lib/dQA_utils.py:51-52: exponents (1-xi), (1+xi);
                     lib/dQA_utils.py:106: 1-patterns[mu,i-1]*(-1)**s_i)
    q[mu,i,s,k]   = k * e[mu,i,s]                               in {0,2,..,2N}
                    phase = exp(i*pi*q/(N+1)) (lib/dQA_utils.py:47,106)
    All int64.  These must be reproduced bit-exactly by the device path.
    """
    xi = check_patterns(xi).astype(np.int64)
    n_xi, n = xi.shape
    sgn = np.array([1, -1], dtype=np.int64)  # (-1)**s
    e = np.rint(1 - xi[:, :, None] * sgn[None, None, :]).astype(np.int64)
    hbit = (1 - xi) // 2
    k = np.arange(n + 1, dtype=np.int64)
    q = e[:, :, :, None] * k[None, None, None, :]
    return hbit, e, q


# ----------------------------------------------------------------------------
# a3 -- Fourier tables (dQA_mps.py:62-66, lib/dQA_utils.py:78-95)
# ----------------------------------------------------------------------------


def h_perceptron(m, n):
    """h(m) = 0 for m >= 0 else -m/sqrt(N)  (lib/dQA_utils.py:78-85)."""
    m = np.asarray(m)
    return np.where(m >= 0, 0, -m / np.sqrt(n))


def f_perceptron(x, n):
    """f(x) = h(N - 2x), x = Hamming distance  (lib/dQA_utils.py:87-95)."""
    return h_perceptron(n - 2 * np.asarray(x), n)


def fourier_tables(n, p_steps, dt):
    """Uk_FT[:,p] = FFT_ortho(exp(-i (p+1)/P dt f(x))), fxft = FFT_ortho(f(x)), x=0..N
    (dQA_mps.py:62-66)."""
    fx = f_perceptron(np.arange(n + 1), n)
    uk = np.zeros((n + 1, p_steps), dtype=CDT)
    for p in range(p_steps):
        uk[:, p] = np.fft.fft(np.exp(-1.0j * ((p + 1) / p_steps) * dt * fx), norm="ortho")
    fxft = np.fft.fft(fx, norm="ortho")
    return uk, fxft


def eps0_closed_form(n, n_xi):
    """eps(0) on |+>^N: (N_xi/N) * sum_x C(N,x) 2^-N f(x)."""
    fx = f_perceptron(np.arange(n + 1), n)
    w = np.array([math.comb(n, x) for x in range(n + 1)], dtype=np.float64) / 2.0**n
    return n_xi / n * float(np.dot(w, fx))


# ----------------------------------------------------------------------------
# a4 / a8 -- operator builders (lib/dQA_utils.py:26-75, :98-109)
# ----------------------------------------------------------------------------


def make_Ux(n, beta_p):
    """Bond-1 MPO of exp(+i beta sigma^x) on every site (lib/dQA_utils.py:26-30)."""
    c, s = np.cos(beta_p), np.sin(beta_p)
    tb = np.array([[c, 1j * s], [1j * s, c]], dtype=CDT)
    return [tb[None, None, :, :].copy() for _ in range(n)]


def _wz_diag(n, uk, xi_i):
    """Diagonal entries W[k,s] of Eq. 17's site tensor (lib/dQA_utils.py:46-53):
    (Uk[k]/sqrt(N+1))**(1/N) * exp(i pi k (1 -+ xi)/(N+1))."""
    d = len(uk)
    coeff = np.power(uk / np.sqrt(n + 1), 1 / n)
    exx = 1j * np.arange(d) * np.pi / (n + 1)
    w = np.empty((d, 2), dtype=CDT)
    w[:, 0] = coeff * np.exp(exx * (1 - xi_i))
    w[:, 1] = coeff * np.exp(exx * (1 + xi_i))
    return w


def make_Uz(n, uk, xi_mu):
    """MPO of U_z^mu at the step whose Fourier column is `uk` (lib/dQA_utils.py:63-75):
    site 0 (1,D,2,2), bulk (D,D,2,2) diagonal in k and s, site N-1 (D,1,2,2)."""
    if len(xi_mu) != n:
        raise AssertionError("not matching dims")
    d = len(uk)
    out = []
    for i in range(n):
        w = _wz_diag(n, uk, xi_mu[i])
        if i == 0:
            t = np.zeros((1, d, 2, 2), dtype=CDT)
            for s in range(2):
                t[0, :, s, s] = w[:, s]
        elif i == n - 1:
            t = np.zeros((d, 1, 2, 2), dtype=CDT)
            for s in range(2):
                t[:, 0, s, s] = w[:, s]
        else:
            t = np.zeros((d, d, 2, 2), dtype=CDT)
            kk = np.arange(d)
            for s in range(2):
                t[kk, kk, s, s] = w[:, s]
        out.append(t)
    return out


def Hz_mu_singleK(n, mu, k, fxft, patterns, n_ancilla=0):
    """Bond-1 product operator H^{mu,k} (lib/dQA_utils.py:98-109); with trailing
    identity sites when n_ancilla>0 (lib/dQA_loss.py:28-46)."""
    out = []
    base = np.power(fxft[k] / np.sqrt(n + 1), 1 / n)
    for i in range(n):
        t = np.zeros((1, 1, 2, 2), dtype=CDT)
        for s in range(2):
            t[0, 0, s, s] = base * np.exp(1.0j * (np.pi / (n + 1)) * k * (1 - patterns[mu, i] * (-1) ** s))
        out.append(t)
    for _ in range(n_ancilla):
        out.append(np.eye(2, dtype=CDT)[None, None, :, :].copy())
    return out


# ----------------------------------------------------------------------------
# a5 / a9 -- contraction primitives (lib/tenn.py:67-109)
# ----------------------------------------------------------------------------


def contract_site(a, w):
    """One MPS site (l,phys,r) with one MPO site (l,r,up,down): contracts the MPS
    physical leg with MPO index 3 and keeps index 2; merged bonds are
    (mps index)*D + (mpo index)  (lib/tenn.py:67-81)."""
    t = np.einsum("abc,jklb->ajckl", a, w, optimize=True)
    t = t.reshape(a.shape[0] * w.shape[0], a.shape[2] * w.shape[1], w.shape[3])
    return np.transpose(t, (0, 2, 1))


def apply_mpsmpo(mps, mpo):
    """lib/tenn.py:83-90."""
    if len(mps) != len(mpo):
        raise AssertionError("must have same number of sites")
    return [contract_site(a, w) for a, w in zip(mps, mpo)]


def braket(bra, ket):
    """Transfer-matrix sweep; conjugates its SECOND argument (lib/tenn.py:93-109).
    Returns the final (r_bra, r_ket) environment, (1,1) for closed chains."""
    if len(bra) != len(ket):
        raise AssertionError("length mismatch")
    env = None
    for b, k in zip(bra, ket):
        site = np.einsum("axb,cxd->acbd", b, np.conj(k), optimize=True)
        env = site[0, 0] if env is None else np.einsum("xy,xyab->ab", env, site, optimize=True)
    return env


# ----------------------------------------------------------------------------
# a6 -- QR right-canonisation (lib/tenn.py:152-190, :294-304)
# ----------------------------------------------------------------------------


def qr_stabilized(x, sgn0fix=True):
    """Reduced QR with the diagonal of R made non-negative (lib/tenn.py:294-304).
    sgn0fix=False keeps the reference's sign(0)=0 (which deletes a column)."""
    q, r = np.linalg.qr(x)
    rd = np.diag(r)
    s = np.sign(rd)
    if sgn0fix:
        s = np.where(rd == 0, np.ones_like(s), s)
    return q * s.reshape(1, -1), r * s.reshape(-1, 1)


def right_canonize_twosites(a, b, sgn0fix=True):
    """lib/tenn.py:152-176: B -> (phys*r, l) matrix, QR, B <- Q, A <- A.R^T, A /= |A|."""
    m = b.transpose(1, 2, 0)
    shp = m.shape
    q, r = qr_stabilized(m.reshape(shp[0] * shp[1], shp[2]), sgn0fix)
    b_new = np.transpose(q.reshape(shp[0], shp[1], -1), (2, 0, 1))
    a_new = np.einsum("abc,dc->abd", a, r, optimize=True)
    a_new = a_new / np.linalg.norm(a_new)
    return a_new, b_new


def right_canonize(mps, site, sgn0fix=True):
    """lib/tenn.py:179-190 (mutates and returns the list, like the reference)."""
    assert 1 <= site < len(mps)
    for n in reversed(range(site, len(mps))):
        mps[n - 1], mps[n] = right_canonize_twosites(mps[n - 1], mps[n], sgn0fix)
    return mps


# ----------------------------------------------------------------------------
# a7 -- truncated-SVD compression (lib/tenn.py:229-262, :307-405)
# ----------------------------------------------------------------------------


def trim_rule(s, cutoff=1e-10, cutoff_mode=4, max_bond=-1, renorm=2):
    """The trimming decision of lib/tenn.py:317-354 on a singular-value vector.
    Returns (n_chi, scale) with `scale` the factor applied to the kept values."""
    tot = csp = None
    pw = 2
    if cutoff > 0.0 or renorm > 0:
        if cutoff_mode == 1:
            n_chi = int(np.count_nonzero(s > cutoff))
        elif cutoff_mode == 2:
            n_chi = int(np.count_nonzero(s > cutoff * s[0]))
        else:
            pw = 2 if cutoff_mode in (3, 4) else 1
            csp = np.cumsum(s**pw, 0)
            tot = csp[-1]
            if cutoff_mode in (4, 6):
                n_chi = int(np.count_nonzero(csp < (1 - cutoff) * tot)) + 1
            else:
                n_chi = int(np.count_nonzero((tot - csp) > cutoff)) + 1
        n_chi = max(n_chi, 1)
        if max_bond > 0:
            n_chi = min(n_chi, max_bond)
    elif max_bond > 0:
        n_chi = max_bond
    else:
        n_chi = s.shape[0]
    scale = 1.0
    if n_chi < s.shape[0] and renorm > 0:
        scale = (tot / csp[n_chi - 1]) ** (1 / pw)
    return n_chi, scale


def svd_truncated(x, cutoff=-1.0, cutoff_mode=3, max_bond=-1, absorb=0, renorm=0, svd=None):
    """lib/tenn.py:372-405 + :307-368.  `svd` lets tests swap the LAPACK driver."""
    u, s, vh = (svd or np.linalg.svd)(x, full_matrices=False)
    n_chi, scale = trim_rule(s, cutoff, cutoff_mode, max_bond, renorm)
    if n_chi < s.shape[0]:
        s = s[:n_chi] * scale
        u = u[:, :n_chi]
        vh = vh[:n_chi, :]
    if absorb is None:
        return u, s, vh
    if absorb == -1:
        u = u * s.reshape(1, -1)
    elif absorb == 1:
        vh = vh * s.reshape(-1, 1)
    else:
        rs = np.sqrt(s)
        u = u * rs.reshape(1, -1)
        vh = vh * rs.reshape(-1, 1)
    return u, vh


def compress_normalize_onesite(a, a_next, max_bd, svd=None):
    """lib/tenn.py:229-252 with the hard-coded policy cutoff=1e-10, mode 4, renorm 2,
    absorb right (lib/tenn.py:238)."""
    u, v = svd_truncated(a.reshape(-1, a.shape[2]), cutoff=1e-10, absorb=1, max_bond=max_bd,
                         cutoff_mode=4, renorm=2, svd=svd)
    return u.reshape(a.shape[0], a.shape[1], -1), np.tensordot(v, a_next, axes=(1, 0))


def compress_svd_normalized(mps, max_bd, svd=None):
    """lib/tenn.py:255-262 (mutates and returns the list)."""
    n = len(mps)
    for l in range(n):
        if l == n - 1:
            mps[l] = mps[l] / np.linalg.norm(mps[l].ravel())
        else:
            mps[l], mps[l + 1] = compress_normalize_onesite(mps[l], mps[l + 1], max_bd, svd)
    return mps


# ----------------------------------------------------------------------------
# a9 / a11 -- energy density (dQA_mps.py:68-81, lib/dQA_loss.py:67-78)
# ----------------------------------------------------------------------------


def compute_loss(psi, patterns, fxft, n_ancilla=0):
    """eps = (1/N) sum_mu sum_k <psi| H^{mu,k} |psi>; complex scalar."""
    n = patterns.shape[1]
    n_tens = len(psi) - n_ancilla
    eps = 0.0
    for mu in range(patterns.shape[0]):
        for k in range(n + 1):
            mpo = Hz_mu_singleK(n, mu, k, fxft, patterns, n_ancilla)
            eps = eps + braket(apply_mpsmpo(psi, mpo), psi) / n_tens
    return eps[0, 0]


def compute_loss_fast(psi, patterns, fxft, n_ancilla=0):
    """Same number as compute_loss, vectorised over k (the site operator is a pure
    phase on the physical leg, so no MPO needs materialising).  Used where the
    tests need many energies quickly; checked against compute_loss in the tests."""
    n = patterns.shape[1]
    n_tens = len(psi) - n_ancilla
    d = n + 1
    kk = np.arange(d)
    base = np.power(fxft / np.sqrt(d), 1 / n)
    total = 0.0 + 0.0j
    for mu in range(patterns.shape[0]):
        env = np.ones((d, 1, 1), dtype=CDT)
        for i, a in enumerate(psi):
            if i < n:
                ph = np.stack([base * np.exp(1j * np.pi / d * kk * (1 - patterns[mu, i] * (-1) ** s))
                               for s in range(2)], axis=1)  # (D,2)
            else:
                ph = np.ones((d, 2), dtype=CDT)
            # env[k,a,c] -> env'[k,b,d] = sum_{a,c,s} env[k,a,c] ph[k,s] A[a,s,b] conj(A[c,s,d])
            t = np.einsum("kac,asb->kcsb", env, a, optimize=True)
            t = t * ph[:, None, :, None]
            env = np.einsum("kcsb,csd->kbd", t, np.conj(a), optimize=True)
        total += env[:, 0, 0].sum()
    return total / n_tens


# ----------------------------------------------------------------------------
# a1 / a10 -- the driver (dQA_mps.py:38-174)
# ----------------------------------------------------------------------------


def plus_state(n):
    """|+>^N as N bond-1 sites (dQA_mps.py:139)."""
    return [np.full((1, 2, 1), 2**-0.5, dtype=CDT) for _ in range(n)]


class OracleDQA:
    """Same constructor/attributes as the reference's mydQA (dQA_mps.py:38-60)."""

    def __init__(self, dataset, P, dt, max_bond=10, sgn0fix=True, svd=None, fast_loss=False):
        self.dataset = np.load(dataset) if isinstance(dataset, str) else np.asarray(dataset)
        self.N_xi, self.N = self.dataset.shape
        self.P, self.dt, self.tau, self.max_bond = P, dt, dt * P, max_bond
        self.pp = 0
        self.sgn0fix, self.svd, self.fast_loss = sgn0fix, svd, fast_loss
        self.snapshots = None  # set to a list to record (pp, mu, stage, mps) tuples

    def init_fourier(self):
        self.Uk_FT, self.fxft = fourier_tables(self.N, self.P, self.dt)

    def compute_loss(self, psi):
        f = compute_loss_fast if self.fast_loss else compute_loss
        return f(psi, self.dataset, self.fxft)

    def reset(self):
        self.psi = plus_state(self.N)
        self.pp = 0
        self.loss = [(0, self.compute_loss(self.psi))]
        self.bd_monitor = []

    def _snap(self, mu, stage, psi):
        if self.snapshots is not None:
            self.snapshots.append((self.pp, mu, stage, [np.array(t) for t in psi]))

    def apply_pattern(self, psi, mu):
        """One sub-step: U_z^mu, right-canonise, truncate (dQA_mps.py:92-106)."""
        uz = make_Uz(self.N, self.Uk_FT[:, self.pp], self.dataset[mu])
        psi = apply_mpsmpo(psi, uz)
        psi = right_canonize(psi, 1, self.sgn0fix)
        self._snap(mu, "canon", psi)
        psi = compress_svd_normalized(psi, self.max_bond, self.svd)
        self._snap(mu, "trunc", psi)
        return psi

    def single_step(self):
        """dQA_mps.py:84-119."""
        psi = self.psi
        s_p = (self.pp + 1) / self.P
        beta_p = (1 - s_p) * self.dt
        for mu in range(self.N_xi):
            psi = self.apply_pattern(psi, mu)
            self.bd_monitor.append(psi[int(self.N / 2)].shape[0])
        psi = apply_mpsmpo(psi, make_Ux(self.N, beta_p))
        self._snap(-1, "mixer", psi)
        expv = self.compute_loss(psi)
        self.psi = psi
        self.loss.append((s_p, expv))
        self.pp += 1
        return expv

    def run(self, steps=None):
        self.reset()
        for _ in range(self.P if steps is None else steps):
            self.single_step()


# ----------------------------------------------------------------------------
# gauge-invariant comparison helpers (protocol P1)
# ----------------------------------------------------------------------------


def mps_norm2(mps):
    return braket(mps, mps)[0, 0].real


def overlap(a, b):
    """<b|a> (braket conjugates its second argument)."""
    return braket(a, b)[0, 0]


def fidelity_defect(a, b):
    """1 - |<a|b>| / (|a||b|)."""
    return 1.0 - abs(overlap(a, b)) / math.sqrt(mps_norm2(a) * mps_norm2(b))


def schmidt_spectra(mps):
    """Singular values at every bond of a (not necessarily canonical) MPS."""
    work = [np.array(t) for t in mps]
    work = right_canonize(work, 1, True)
    out = []
    for l in range(len(work) - 1):
        a = work[l]
        u, s, vh = np.linalg.svd(a.reshape(-1, a.shape[2]), full_matrices=False)
        out.append(s)
        work[l] = u.reshape(a.shape[0], a.shape[1], -1)
        work[l + 1] = np.tensordot(vh * s.reshape(-1, 1), work[l + 1], axes=(1, 0))
    return out


def to_dense(mps):
    """State vector of an MPS, site 0 = most significant bit."""
    v = mps[0]
    for t in mps[1:]:
        v = np.tensordot(v, t, axes=(v.ndim - 1, 0))
    return v.reshape(-1)


# ----------------------------------------------------------------------------
# second oracle: exact (untruncated) state-vector dQA, N <= ~16
# ----------------------------------------------------------------------------


def statevector_dqa(xi, P, dt, steps=None):
    """Dense 2^N Trotterised dQA: psi <- exp(-i gamma_p f(x_mu(sigma))) psi for each mu, then
    exp(+i beta_p sigma^x) on every qubit; eps = <sum_mu f(x_mu)>/N.  Site 0 is the most
    significant bit, s=0 <-> sigma=+1 (conventions of lib/HammingEvolution.py:133-148 and
    SURVEY section 8c).  Returns eps(s) for s = 0, 1/P, ... (length steps+1)."""
    xi = check_patterns(xi).astype(np.int64)
    n_xi, n = xi.shape
    dim = 1 << n
    idx = np.arange(dim)
    bits = (idx[:, None] >> (n - 1 - np.arange(n))[None, :]) & 1  # s_i of site i
    sigma = 1 - 2 * bits
    fx = np.stack([f_perceptron((n - (sigma * xi[mu][None, :]).sum(1)) // 2, n) for mu in range(n_xi)])
    hz = fx.sum(0)
    psi = np.full(dim, 2.0 ** (-n / 2), dtype=CDT)
    out = [np.vdot(psi, hz * psi).real / n]
    for p in range(P if steps is None else steps):
        s_p = (p + 1) / P
        gamma, beta = s_p * dt, (1 - s_p) * dt
        psi = psi * np.exp(-1j * gamma * hz)
        t = psi.reshape((2,) * n)
        c, s = np.cos(beta), 1j * np.sin(beta)
        for i in range(n):
            t = np.moveaxis(t, i, 0)
            t = np.stack([c * t[0] + s * t[1], s * t[0] + c * t[1]])
            t = np.moveaxis(t, 0, i)
        psi = t.reshape(dim)
        out.append(np.vdot(psi, hz * psi).real / n)
    return np.array(out)
