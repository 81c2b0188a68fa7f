"""TEST INFRASTRUCTURE ONLY -- numpy-backed stand-in for the `jax` package.

The reference (baronefr/perceptron-dqa) imports jax only for `jit`, `devices`,
`default_device`, `jax.config.config.update` and `jax.numpy` (dQA_mps.py:26-27,
lib/tenn.py:21-22).  None of those is installed in this image, so the oracle
runs the reference's *unmodified* files on top of this shim: `jit` is the
identity, devices are dummies, and `jax.numpy` is numpy (complex128 end to end).
Nothing in the product package imports this.
"""
import contextlib

from . import numpy  # noqa: F401  (jax.numpy)
from . import config  # noqa: F401 (jax.config)


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


class _Device:
    def __init__(self, kind):
        self.platform = kind

    def __repr__(self):
        return f"ShimDevice({self.platform})"


def devices(kind=None):
    return [_Device(kind or "cpu")]


@contextlib.contextmanager
def default_device(_dev):
    yield
