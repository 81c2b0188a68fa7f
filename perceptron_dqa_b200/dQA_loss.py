"""`mydQA_ancilla` (lib/dQA_loss.py:53-84): compute_loss of an externally produced MPS that carries
`n_ancilla` trailing ancilla sites, with optional endian flip of the patterns -- the entry point the
reference's circuit driver uses (dQA_circuit.py:152-156,177-180,219).  The energy sweep runs in
libdqa.so (k_energy_ext: the transfer-matrix kernel with identity factors on the ancilla sites)."""
from __future__ import annotations

import numpy as np

from . import tenn
from .dQA_mps import _device_index, mydQA
from .dQA_utils import as_patterns

__all__ = ["mydQA_ancilla", "Hz_mu_singleK_with_ancilla"]


def Hz_mu_singleK_with_ancilla(N, mu, K, f_FT_, patterns, extra_ancillary):
    """Host-side builder with the reference's name (lib/dQA_loss.py:28-46), for op-level callers."""
    from .dQA_utils import PerceptronHamiltonian

    out = PerceptronHamiltonian.Hz_mu_singleK(N, mu, K, f_FT_, patterns)
    for _ in range(extra_ancillary):
        out.append(np.eye(2, dtype=np.complex128).reshape(1, 1, 2, 2).copy())
    return out


class mydQA_ancilla(mydQA):
    def __init__(self, dataset, P: int, dt: float, device=None, n_ancilla: int = 0, flip_endian: bool = False):
        assert n_ancilla >= 0, "number of ancillas must be non negative"
        dataset = np.load(dataset) if isinstance(dataset, str) else np.asarray(dataset)
        if flip_endian:
            dataset = dataset.copy()[:, ::-1]
        super().__init__(dataset, P, dt, 10, device)   # max_bond pinned to 10 as in lib/dQA_loss.py:62
        super().init_fourier()
        self.n_ancilla = n_ancilla

    def compute_loss(self, psi):
        return tenn.ops(_device_index(self.device)).energy(list(psi), as_patterns(self.dataset), self.fxft, self.n_ancilla)

    def run(self) -> None:
        raise Exception("do not use this class to run dQA!")

    def single_step(self) -> None:
        raise Exception("do not use this class to run dQA!")
