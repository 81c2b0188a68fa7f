"""perceptron-dqa_b200: B200-native engine for the digitized-Quantum-Annealing MPS step.

Drop-in surface (same names as the reference, baronefr/perceptron-dqa):
    mydQA                      dQA_mps.py:38-185
    mydQA_ancilla              lib/dQA_loss.py:53-84
    PerceptronHamiltonian      lib/dQA_utils.py:23-109
    apply_mpsmpo, braket, right_canonize, compress_svd_normalized, svd_truncated,
    qr_stabilized              lib/tenn.py
Everything numeric runs in libdqa.so (hand-written sm_100a CUDA behind a C ABI,
include/dqa.h); importing this package never falls back to a CPU implementation.
"""
from ._lib import DqaError, DqaNumericError, Plan, load  # noqa: F401

__all__ = ["Plan", "load", "DqaError", "DqaNumericError"]
