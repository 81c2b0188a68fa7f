"""Host-side operator builders with the reference's names (lib/dQA_utils.py).

These return lists of numpy site tensors exactly like the reference's
`PerceptronHamiltonian`, for callers that use the op-level API
(`apply_mpsmpo(psi, make_Uz(...))`).  The hot path itself never materialises these
tensors: libdqa.so applies U_z in the sector representation and folds the phases of
H^{mu,k} into the energy kernel.  Pure table building -- no step arithmetic happens here.
"""
from __future__ import annotations

import numpy as np

__all__ = ["PerceptronHamiltonian", "create_dataset", "as_patterns"]


def as_patterns(dataset):
    """+-1 patterns as int8 after validation (the reference's files are int64, float64 or
    float32: SURVEY App. B Q11)."""
    a = np.asarray(dataset)
    if a.ndim not in (2, 3):
        raise ValueError("patterns must be (N_xi, N) or (batch, N_xi, N)")
    if not np.all((a == 1) | (a == -1)):
        raise ValueError("patterns must be +-1")
    return a.astype(np.int8)


class PerceptronHamiltonian:
    """Same static-style interface as lib/dQA_utils.py:23-109."""

    @staticmethod
    def make_Ux(N, beta_p, dtype=np.complex128):
        """lib/dQA_utils.py:26-30."""
        c, s = np.cos(beta_p), 1j * np.sin(beta_p)
        tb = np.array([[c, s], [s, c]], dtype=dtype)
        return [tb.reshape(1, 1, 2, 2).copy() for _ in range(N)]

    @staticmethod
    def Wz(N, Uk, xi, marginal=None, dtype=np.complex128):
        """Eq. 17 site tensor (lib/dQA_utils.py:32-61): diagonal in the bond index k and
        in the spin, value (Uk[k]/sqrt(N+1))**(1/N) * exp(i pi k (1 -+ xi)/(N+1))."""
        Uk = np.asarray(Uk)
        D = len(Uk)
        k = np.arange(D)
        coeff = np.power(Uk / np.sqrt(N + 1), 1 / N)
        ang = 1j * k * np.pi / (N + 1)
        w = np.stack([coeff * np.exp(ang * (1 - xi)), coeff * np.exp(ang * (1 + xi))], axis=1)  # (D,2)
        if marginal == "l":
            t = np.zeros((1, D, 2, 2), dtype=dtype)
            t[0, k, 0, 0], t[0, k, 1, 1] = w[:, 0], w[:, 1]
        elif marginal == "r":
            t = np.zeros((D, 1, 2, 2), dtype=dtype)
            t[k, 0, 0, 0], t[k, 0, 1, 1] = w[:, 0], w[:, 1]
        else:
            t = np.zeros((D, D, 2, 2), dtype=dtype)
            t[k, k, 0, 0], t[k, k, 1, 1] = w[:, 0], w[:, 1]
        return t

    @staticmethod
    def make_Uz(N, Uk, xi, dtype=np.complex128):
        """lib/dQA_utils.py:63-75."""
        assert len(xi) == N, "not matching dims"
        W = PerceptronHamiltonian.Wz
        return [W(N, Uk, xi[0], "l", dtype)] + [W(N, Uk, xi[i], None, dtype) for i in range(1, N - 1)] + [W(N, Uk, xi[N - 1], "r", dtype)]

    @staticmethod
    def h_perceptron(m, N):
        """lib/dQA_utils.py:78-85."""
        m = np.array(m)
        return np.where(m >= 0, 0, -m / np.sqrt(N)).squeeze()

    @staticmethod
    def f_perceptron(x, N):
        """lib/dQA_utils.py:87-95."""
        return PerceptronHamiltonian.h_perceptron(N - 2 * np.asarray(x), N)

    @staticmethod
    def Hz_mu_singleK(N, mu, K, f_FT_, patterns):
        """lib/dQA_utils.py:98-109."""
        base = np.power(f_FT_[K] / np.sqrt(N + 1), 1 / N)
        out = []
        for i in range(N):
            t = np.zeros((1, 1, 2, 2), dtype=np.complex128)
            for s in range(2):
                t[0, 0, s, s] = base * np.exp(1.0j * (np.pi / (N + 1)) * K * (1 - patterns[mu, i] * (-1) ** s))
            out.append(t)
        return out


def create_dataset(N, features):
    """lib/dQA_utils.py:112-116 (N patterns of `features` +-1 entries)."""
    x = np.random.randint(2, size=(N, features))
    x[x == 0] = -1
    return x
