"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/dqa.h declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g

    g.build()
    return g.LIB


def declared_functions():
    src = open(os.path.join(ROOT, "include", "dqa.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dqa_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("dqa_create", "dqa_step", "dqa_apply_pattern", "dqa_apply_ux", "dqa_energy", "dqa_set_mps",
                 "dqa_get_mps", "dqa_get_index_tensors", "dqa_ensemble_moments", "dqa_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from perceptron_dqa_b200 import _lib

    assert sorted(_lib.EXPORTS) == declared_functions()


def test_sass_is_sm100a(lib_path):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_product_fails_loudly_without_gpu(lib_path):
    """No CPU fallback: with no CUDA device, building a plan must raise, not compute."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from perceptron_dqa_b200._lib import Plan

    with pytest.raises(Exception):
        Plan(7, 5, 4, 1.0, 4)


def test_argument_errors_without_gpu(lib_path):
    from perceptron_dqa_b200 import _lib

    lib = _lib.declare(ctypes.CDLL(lib_path))
    h = ctypes.c_void_p()
    cfg = _lib.dqa_config(N=1, n_xi=1, P=1, chi=1, batch=1, device=0, max_sweeps=0, reserved=0, dt=1.0, cutoff=0.0, stream=None)
    rc = lib.dqa_create(ctypes.byref(h), ctypes.byref(cfg))
    assert rc == -1 and b"N>=2" in lib.dqa_last_error(h)
    lib.dqa_destroy(h)
    assert lib.dqa_create(None, None) == -1
