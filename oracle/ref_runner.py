"""TEST INFRASTRUCTURE ONLY -- run the reference's own files, unmodified, in complex128.

Works where /root/reference exists (the build container) or where oracle/make_ref.py has staged the four
hot-path files into oracle/_ref/ (the GPU box: bench.py's CPU legs only).  Used by
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

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref")  # verbatim copies made by oracle/make_ref.py (git-ignored, travels with gpurun)


def _pick_root():
    env = os.environ.get("DQA_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(cand, "dQA_mps.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()
_SHIM = os.path.join(_HERE, "shim")


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
