"""TEST INFRASTRUCTURE ONLY -- run the reference's own files, unmodified, in complex128.

Works only where /root/reference exists (the build container).  Used by
oracle/make_golden.py to produce the fixtures under tests/golden/ and by
tests/test_oracle_vs_reference.py to pin oracle/dqa_oracle.py (the travelling
numpy restatement) against the real reference.  Never imported by the product.

Recipe: SURVEY.md App. E.  The shim directory shadows jax/quimb/matplotlib; the
reference's dQA_mps.py, lib/tenn.py, lib/dQA_utils.py, lib/dQA_loss.py are then
imported as they are.
"""
import contextlib
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("DQA_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dQA_mps.py"))


def load(sgn0fix=True, svd_driver="gesdd"):
    """Import the reference behind the shim and return its `dQA_mps` module.

    sgn0fix     sign(0):=1 in qr_stabilized; False = verbatim.
    svd_driver  'gesdd' = numpy.linalg.svd as the reference calls it; 'gesvd'
                swaps in scipy's driver (only to measure the oracle's own spread).
    """
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import jax.numpy as jnp  # the shim

    jnp.SGN0FIX = bool(sgn0fix)
    import numpy as np

    if svd_driver == "gesvd":
        import scipy.linalg as sla

        if not hasattr(np.linalg, "_dqa_orig_svd"):
            np.linalg._dqa_orig_svd = np.linalg.svd

        def _svd(x, full_matrices=True, **kw):
            return sla.svd(x, full_matrices=full_matrices, lapack_driver="gesvd")

        np.linalg.svd = _svd
    elif hasattr(np.linalg, "_dqa_orig_svd"):
        np.linalg.svd = np.linalg._dqa_orig_svd
    import dQA_mps  # noqa: E402  the reference driver, verbatim

    return dQA_mps


def load_ancilla():
    load_mod = load()
    import lib.dQA_loss as dqa_loss  # reference lib/dQA_loss.py, verbatim

    return load_mod, dqa_loss


@contextlib.contextmanager
def quiet():
    """Silence tqdm / print_info of the reference."""
    old = sys.stderr
    sys.stderr = io.StringIO()
    try:
        yield
    finally:
        sys.stderr = old
