"""Shared test helpers: golden fixture access and gauge-invariant comparisons."""
import os

import numpy as np

import dqa_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def unpack_mps(bonds, flat):
    out, off = [], 0
    for i in range(len(bonds) - 1):
        l, r = int(bonds[i]), int(bonds[i + 1])
        out.append(np.array(flat[off:off + l * 2 * r]).reshape(l, 2, r))
        off += l * 2 * r
    assert off == len(flat)
    return out


def snapshots(g):
    """[(p, mu, pre_mps, post_mps)] of a golden file."""
    res = []
    for i, (p, mu) in enumerate(g["snap_index"]):
        pre = unpack_mps(g[f"snap{i}_pre_bonds"], g[f"snap{i}_pre_flat"])
        post = unpack_mps(g[f"snap{i}_post_bonds"], g[f"snap{i}_post_flat"])
        res.append((int(p), int(mu), pre, post))
    return res


def compare_states(got, want, xi, fxft, tol_fid=1e-12, tol_eps=1e-10, tol_s=1e-12):
    """Gauge-invariant comparison: bond dims exact, fidelity defect,
    Schmidt spectra at every bond, energy density."""
    assert [t.shape for t in got] == [t.shape for t in want], ([t.shape for t in got], [t.shape for t in want])
    fid = O.fidelity_defect(got, want)
    assert abs(fid) < tol_fid, f"fidelity defect {fid}"
    assert abs(O.mps_norm2(got) - 1.0) < 1e-12
    sg, sw = O.schmidt_spectra(got), O.schmidt_spectra(want)
    for a, b in zip(sg, sw):
        n = min(len(a), len(b))
        assert np.abs(a[:n] - b[:n]).max() < tol_s
    eg = O.compute_loss_fast(got, xi, fxft)
    ew = O.compute_loss_fast(want, xi, fxft)
    assert abs(eg - ew) < tol_eps, f"eps {eg} vs {ew}"
    return fid
