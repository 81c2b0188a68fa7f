"""Ensemble sharding over the GPUs of one node.

Independent pattern realisations share nothing but (N, P, dt, chi) and the Fourier tables
(the reference runs them as separate script invocations: benchmarker-mps.py:91-95 and
averages offline, data/plotter.py:175-195).  One process per GPU evolves its own block of
realisations with libdqa.so; the only exchange is ONE all-reduce of 3*(P+1) doubles
[sum eps(s), sum eps(s)^2, #healthy] at the end -- NCCL over NVLink on GPUs, gloo in the CPU
tests of this host-side logic.  No other collective exists on this path.
"""
from __future__ import annotations

import numpy as np

__all__ = ["shard_bounds", "moments_from_eps", "finalize_moments", "EnsembleDQA"]


def shard_bounds(n_real, world, rank):
    """Contiguous block partition: realisation r lives on rank floor(r*world/n_real)."""
    lo = -(-rank * n_real // world)
    hi = -(-(rank + 1) * n_real // world)
    return lo, hi


def moments_from_eps(eps, status):
    """Host mirror of k_ensemble_moments (same layout), for tests and for traces that are
    already on the host: eps (B,P+1) complex, status (B,) int -> (3*(P+1),) float64."""
    eps = np.asarray(eps)
    ok = np.asarray(status) == 0
    re = eps.real[ok]
    return np.concatenate([re.sum(0), (re * re).sum(0), np.full(eps.shape[1], float(ok.sum()))])


def finalize_moments(mom):
    """-> mean eps(s), standard error, number of realisations that contributed."""
    mom = np.asarray(mom, dtype=np.float64)
    n = mom.shape[0] // 3
    s1, s2, cnt = mom[:n], mom[n:2 * n], mom[2 * n:]
    c = np.maximum(cnt, 1.0)
    mean = s1 / c
    var = np.maximum(s2 / c - mean * mean, 0.0)
    sem = np.sqrt(var / np.maximum(cnt - 1.0, 1.0))
    return mean, sem, cnt


class EnsembleDQA:
    """R realisations sharded over torch.distributed ranks, one GPU each.

        ens = EnsembleDQA(xi_all, P, dt, max_bond)      # xi_all: (R, N_xi, N), same on all ranks
        ens.run(); ens.mean, ens.sem, ens.count         # identical on all ranks
    """

    def __init__(self, xi_all, P, dt, max_bond=10, device=None):
        import torch.distributed as dist

        from .dQA_mps import mydQA

        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        xi_all = np.asarray(xi_all)
        self.lo, self.hi = shard_bounds(xi_all.shape[0], self.world, self.rank)
        self.local = mydQA(xi_all[self.lo:self.hi], P=P, dt=dt, max_bond=max_bond, device=device)
        self.P = P

    def run(self, chunk=0):
        import torch

        self.local.init_fourier()
        self.local.run(print_info=False, chunk=chunk)
        plan = self.local.plan
        mom = torch.zeros(3 * (self.P + 1), dtype=torch.float64, device=f"cuda:{plan.device}")
        plan.ensemble_moments_into(mom.data_ptr())
        if self.dist:
            self.dist.all_reduce(mom)
        self.mean, self.sem, self.count = finalize_moments(mom.cpu().numpy())
        return self.mean
