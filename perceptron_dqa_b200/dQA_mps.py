"""`mydQA`: the reference driver's interface (dQA_mps.py:38-185) on top of libdqa.so.

    obj = mydQA('patterns.npy', P=100, dt=1.0, max_bond=10)
    obj.init_fourier(); obj.run(); obj.loss[-1][1]

Same constructor, methods and public attributes (`loss`, `bd_monitor`, `psi`, `Uk_FT`,
`fxft`, `N`, `N_xi`, `P`, `dt`, `tau`, `max_bond`, `pp`, `dataset`).  Extras that default
to the reference behaviour: a 3-d `dataset` (batch, N_xi, N) evolves `batch` independent
realisations at once on this GPU (`.eps` then holds the (batch, P+1) traces); with such a
batch `dt` may be a sequence of `batch` values (the dt grid of benchmarker-mps.py:50-57 on the
ensemble dimension: one Fourier table per distinct dt, `Uk_FT` is then (n_dt, N+1, P)).
Host code here only builds the two small Fourier tables (as dQA_mps.py:62-66 does) and
moves results; all step arithmetic is CUDA.
"""
from __future__ import annotations

import numpy as np
import numpy.fft as fft

from ._lib import DqaNumericError, Plan
from .dQA_utils import PerceptronHamiltonian, as_patterns

__all__ = ["mydQA"]


def _device_index(device):
    if device is None:
        return 0
    if isinstance(device, int):
        return device
    idx = getattr(device, "index", None)  # torch.device
    if idx is not None:
        return idx
    return int(getattr(device, "id", 0))  # a jax-like device object


class mydQA:
    def __init__(self, dataset, P: int, dt: float, max_bond: int = 10, device=None, _plan_kw=None):
        self.device = device
        data = np.load(dataset) if isinstance(dataset, str) else np.asarray(dataset)
        self.dataset = data
        self._xi = as_patterns(data)
        self.batch = 1 if self._xi.ndim == 2 else self._xi.shape[0]
        self.N_xi, self.N = self._xi.shape[-2], self._xi.shape[-1]
        self.P = P
        if np.ndim(dt) == 0:
            self.dt, self._dts = dt, None
            self.tau = dt * P
        else:  # one dt per realisation of the batch
            self._dts = np.asarray(dt, dtype=np.float64).reshape(-1)
            if self._dts.shape[0] != self.batch:
                raise ValueError("dt must be a scalar or one value per realisation of the batch")
            self._dt_values, self._dt_index = np.unique(self._dts, return_inverse=True)
            self.dt = self._dts
            self.tau = self._dts * P
        self.max_bond = max_bond
        self.pp = 0
        self._plan_kw = dict(_plan_kw or {})
        self._plan = None
        self._loss_plans = {}
        self._psi_cache = None
        self.loss = []
        self.bd_monitor = []

    # -- tables (dQA_mps.py:62-66) -------------------------------------------------------
    def init_fourier(self) -> None:
        f = PerceptronHamiltonian.f_perceptron(range(self.N + 1), self.N)

        def table(dt):
            uk = np.zeros((self.N + 1, self.P), dtype=np.complex128)
            for p in range(self.P):
                uk[:, p] = fft.fft(np.exp(-1.0j * ((p + 1) / self.P) * dt * f), norm="ortho")
            return uk

        self.Uk_FT = table(self.dt) if self._dts is None else np.stack([table(v) for v in self._dt_values])
        self.fxft = fft.fft(f, norm="ortho")
        if self._plan is not None:
            self._set_tables(self._plan, main=True)

    def _set_tables(self, pl, main):
        if self._dts is None:
            pl.set_fourier(self.Uk_FT, self.fxft)
        elif main:
            pl.set_fourier_tables(self.Uk_FT, self.fxft, self._dt_index, self._dts)
        else:  # single-instance helper plans (compute_loss) only need fxft; any table will do
            pl.set_fourier(self.Uk_FT[0], self.fxft)

    # -- plan handling ---------------------------------------------------------------------
    def _make_plan(self, chi, batch, xi):
        if not hasattr(self, "Uk_FT"):
            raise RuntimeError("call init_fourier() first")
        kw = dict(self._plan_kw)
        kw.setdefault("device", _device_index(self.device))
        main = batch == self.batch and self._dts is not None
        dt0 = self.dt if self._dts is None else float(self._dts[0])
        pl = Plan(self.N, self.N_xi, self.P, dt0, chi, batch=batch, n_tables=len(self._dt_values) if main else 1, **kw)
        pl.set_patterns(xi)
        self._set_tables(pl, main)
        return pl

    @property
    def plan(self) -> Plan:
        if self._plan is None:
            self._plan = self._make_plan(self.max_bond, self.batch, self._xi)
        return self._plan

    # -- state -----------------------------------------------------------------------------
    @property
    def psi(self):
        """Current MPS as the reference keeps it: list of N arrays (l,2,r) (instance 0;
        `psi_of(i)` for the others)."""
        if self._psi_cache is None:
            self._psi_cache = self.plan.get_mps(0)
        return self._psi_cache

    @psi.setter
    def psi(self, mps):
        self.plan.set_mps(list(mps), 0)
        self._psi_cache = None

    def psi_of(self, inst):
        return self.plan.get_mps(inst)

    # -- energy (dQA_mps.py:68-81) -----------------------------------------------------------
    def compute_loss(self, psi):
        """eps of an arbitrary N-site MPS given as the reference's list of (l,2,r) arrays."""
        chi = max(max(t.shape[0], t.shape[2]) for t in psi)
        pl = self._loss_plans.get(chi)
        if pl is None:
            xi0 = self._xi if self._xi.ndim == 2 else self._xi[0]
            pl = self._loss_plans[chi] = self._make_plan(chi, 1, xi0)
        pl.set_mps(list(psi), 0)
        return complex(pl.energy()[0])

    # -- stepping (dQA_mps.py:84-174) ---------------------------------------------------------
    def _pull(self, first_slot):
        eps = self.plan.eps_trace()
        self.eps = eps
        bd = self.plan.bd_monitor()
        for i in range(first_slot, self.pp + 1):
            s = i / self.P if i else 0
            self.loss.append((s, eps[0, i] if self.batch == 1 else eps[:, i].copy()))
        n0 = (first_slot - 1) * self.N_xi if first_slot > 0 else 0
        vals = bd[0, n0:self.pp * self.N_xi] if self.batch == 1 else bd[:, n0:self.pp * self.N_xi].T
        self.bd_monitor.extend(vals.tolist())

    def single_step(self):
        """One dQA step; returns eps(s_p) like the reference (dQA_mps.py:84-119)."""
        if self.plan.pp != self.pp:
            self.plan.pp = self.pp
        self._step_and_pull(1)
        return self.loss[-1][1]

    def _step_and_pull(self, n):
        """n steps, then the traces.  If an instance fails numerically the library has still completed (and
        recorded) the steps: keep `pp`, `loss` and `bd_monitor` in line with the device before re-raising."""
        first = self.pp + 1
        try:
            self.plan.step(n)
        except DqaNumericError:
            self.pp = self.plan.pp
            self._psi_cache = None
            if self.pp >= first:
                self._pull(first)
            raise
        self.pp += n
        self._psi_cache = None
        self._pull(first)

    def run(self, skip_jit=0, print_info: bool = True, chunk: int = 0) -> None:
        """Run all P steps from |+>^N.  `skip_jit` is accepted and ignored (it is a no-op in
        the reference too: SURVEY App. B Q3)."""
        if print_info:
            print("dQA info ---")
            print(" tau = {}, P = {}, dt = {}".format(self.tau, self.P, self.dt))
            print(" max bd =", self.max_bond)
            print(" dataset :  N = {}, N_xi = {}".format(self.N, self.N_xi))
        self.plan.reset_plus()
        self.pp = 0
        self.loss, self.bd_monitor = [], []
        self._psi_cache = None
        self._pull(0)
        chunk = chunk or self.P
        while self.pp < self.P:
            self._step_and_pull(min(chunk, self.P - self.pp))

    def plot_loss(self):
        """dQA_mps.py:177-185."""
        import matplotlib.pyplot as plt

        plt.plot(*zip(*np.real_if_close(self.loss)))
        plt.yscale("log")
        plt.title("dQA")
        return plt
