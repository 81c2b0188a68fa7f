"""TEST INFRASTRUCTURE ONLY -- `jax.numpy` as plain numpy (complex128).

Two switches (module attributes, set by oracle/ref_runner.py; no reference file
is touched):
  SGN0FIX   sign(0) := +1 inside qr_stabilized (lib/tenn.py:298-303 uses
            jnp.sign, whose sign(0)=0 deletes a column of Q).
  SVD_DRIVER  'gesdd' (numpy default) or 'gesvd' (scipy) for numpy.linalg.svd
            (lib/tenn.py:402), to measure the oracle's own roundoff envelope.
"""
import numpy as _np
from numpy import *  # noqa: F401,F403

linalg = _np.linalg
SGN0FIX = False


def einsum(*args, **kw):
    # XLA lowers these contractions to GEMMs; optimize=True makes numpy do the same.
    kw.setdefault("optimize", True)
    return _np.einsum(*args, **kw)


def sign(x):
    s = _np.sign(x)
    if SGN0FIX:
        s = _np.where(x == 0, _np.ones_like(s), s)
    return s
