"""The reference's tensor-network functions (lib/tenn.py) with the same names, argument meaning and
layouts -- MPS site (l, phys, r), MPO site (l, r, up, down), lists of per-site numpy arrays -- computed
by libdqa.so's op-level kernels (csrc/dqa_ops.cuh).  Host code here only loops over sites and reshapes;
no arithmetic happens in Python.  The driver's hot loop uses the fused sector kernels instead.

Deviation: `qr_stabilized` uses sgn(0) := +1, so an exactly zero pivot never deletes a
column of Q.
"""
from __future__ import annotations

from typing import List

import numpy as np

from ._lib import Ops

_ops = {}  # one op context (device scratch + handle) per CUDA device
_forced = None


def ops(device: int = 0) -> Ops:
    if _forced is not None:
        return _forced
    o = _ops.get(device)
    if o is None:
        o = _ops[device] = Ops(device=device)
    return o


def set_ops(o):
    """Use a specific op context (the CPU tests hand in one backed by the SIMT-emulator build)."""
    global _forced
    _forced = o


# ---- layout helpers (pure reshapes, lib/tenn.py:31-61) --------------------------------------------
def convert(mps):
    return [np.transpose(t, (0, 2, 1)) for t in mps]


def remove_margin_extra_dim_in_lrp(mps) -> None:
    mps[0] = mps[0][0]
    mps[-1] = mps[-1][:, 0, :]


def add_margin_extra_dim_in_lrp(mps) -> None:
    mps[0] = np.expand_dims(mps[0], 0)
    mps[-1] = np.expand_dims(mps[-1], 1)


# ---- MPS / MPO operations ----------------------------------------------------------------------------
def contract_mps_mpo_single_site(state_site, operator_site):
    """lib/tenn.py:67-81."""
    return ops().contract_site(state_site, operator_site)


def apply_mpsmpo(mps, mpo):
    """lib/tenn.py:83-90."""
    assert len(mps) == len(mpo), "must have same number of sites"
    return [contract_mps_mpo_single_site(ss, oo) for ss, oo in zip(mps, mpo)]


def braket(bra, ket):
    """<bra|ket> transfer sweep; conjugates `ket` (lib/tenn.py:93-109)."""
    assert len(bra) == len(ket)
    return ops().braket(bra, ket)


def make_rand_mpo_complex(N: int, bond_dim=1):
    """lib/tenn.py:112-125 (host random numbers, as in the reference)."""
    assert N > 1
    shapes = [(1, bond_dim, 2, 2)] + [(bond_dim, bond_dim, 2, 2)] * (N - 2) + [(bond_dim, 1, 2, 2)]
    return [np.random.uniform(low=-1.0, size=s) + 1.0j * np.random.uniform(low=-1.0, size=s) for s in shapes]


def make_rand_mps_complex(N: int, bond_dim=1):
    """lib/tenn.py:128-140."""
    assert N > 1
    shapes = [(1, 2, bond_dim)] + [(bond_dim, 2, bond_dim)] * (N - 2) + [(bond_dim, 2, 1)]
    return [np.random.uniform(low=-1.0, size=s) + 1.0j * np.random.uniform(low=-1.0, size=s) for s in shapes]


# ---- canonisation ------------------------------------------------------------------------------------
def qr_stabilized(x):
    """lib/tenn.py:294-304."""
    return ops().qr_stabilized(x)


def right_canonize_twosites(A, B):
    """lib/tenn.py:152-176."""
    return ops().right_canonize_twosites(A, B)


def right_canonize(mps: List[np.ndarray], site: int):
    """lib/tenn.py:179-190 (mutates and returns the list, like the reference)."""
    assert site < len(mps)
    assert site >= 1
    for n in reversed(range(site, len(mps))):
        mps[n - 1], mps[n] = right_canonize_twosites(mps[n - 1], mps[n])
    return mps


# ---- SVD compression ---------------------------------------------------------------------------------
def svd_truncated(x, cutoff=-1.0, cutoff_mode=3, max_bond=-1, absorb=0, renorm=0):
    """lib/tenn.py:372-405: returns (U, V) with the singular values absorbed, or (U, s, VH) for absorb=None."""
    u, s, vh = ops().svd_truncated(x, cutoff, cutoff_mode, max_bond, absorb, renorm)
    if absorb is None:
        return u, s, vh
    return u, vh


def compress_normalize_onesite(A, A_next, max_bd: int):
    """lib/tenn.py:229-252."""
    return ops().compress_onesite(A, A_next, max_bd)


def compress_svd_normalized(mps, max_bd: int):
    """lib/tenn.py:255-262 (mutates and returns the list)."""
    N = len(mps)
    for l in range(N):
        if l == N - 1:
            mps[l] = ops().normalize(mps[l])
        else:
            mps[l], mps[l + 1] = compress_normalize_onesite(mps[l], mps[l + 1], max_bd=max_bd)
    return mps


def normalize_onesite(A, A_next):
    raise DeprecationWarning("this function is not updated to work on lpr tensors")  # lib/tenn.py:203-209


def normalize_mps_via_svd(mps):
    raise DeprecationWarning("this function is not updated to work on lpr tensors")  # lib/tenn.py:212-219
