"""ctypes binding of libdqa.so (include/dqa.h).

The product path is CUDA only: `load()` raises if the nvcc-built library is missing -- there
is no CPU fallback.  `Plan` is a thin object wrapper over one `dqa_t*`; PyTorch is used for
exactly one thing, owning the device workspace (a uint8 tensor) and naming the stream.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DQA_LIB") or os.path.join(_HERE, "libdqa.so")  # DQA_LIB: another build of the same library (kernel A/B runs)
NFAM = 6
FAMILIES = ("sector_qr", "theta", "lq", "mixer", "energy", "jacobi")


class DqaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdqa error {code}: {msg}")
        self.code = code


class DqaNumericError(DqaError):
    """Some instance went non-finite / lost its norm / Jacobi did not converge (DQA_ENUMERIC)."""


class dqa_config(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("n_xi", C.c_int32), ("P", C.c_int32), ("chi", C.c_int32),
        ("batch", C.c_int32), ("device", C.c_int32), ("max_sweeps", C.c_int32), ("n_tables", C.c_int32),
        ("dt", C.c_double), ("cutoff", C.c_double), ("stream", C.c_void_p),
        ("policy_set", C.c_int32), ("cutoff_mode", C.c_int32), ("renorm", C.c_int32), ("absorb", C.c_int32),
    ]


_P = C.c_void_p
_SIGS = {
    "dqa_create": ([C.POINTER(_P), C.POINTER(dqa_config)], C.c_int),
    "dqa_destroy": ([_P], None),
    "dqa_last_error": ([_P], C.c_char_p),
    "dqa_workspace_bytes": ([_P, C.POINTER(C.c_size_t)], C.c_int),
    "dqa_bind_workspace": ([_P, _P, C.c_size_t], C.c_int),
    "dqa_set_patterns": ([_P, _P], C.c_int),
    "dqa_set_fourier": ([_P, _P, _P], C.c_int),
    "dqa_set_fourier_tables": ([_P, C.c_int, _P, _P, _P, _P], C.c_int),
    "dqa_clear_status": ([_P], C.c_int),
    "dqa_get_truncation_error": ([_P, C.c_int, _P], C.c_int),
    "dqa_get_index_tensors": ([_P, C.c_int, _P, _P, _P], C.c_int),
    "dqa_reset_plus": ([_P], C.c_int),
    "dqa_step": ([_P, C.c_int], C.c_int),
    "dqa_step_async": ([_P, C.c_int], C.c_int),
    "dqa_get_step": ([_P, C.POINTER(C.c_int)], C.c_int),
    "dqa_set_step": ([_P, C.c_int], C.c_int),
    "dqa_apply_pattern": ([_P, C.c_int, C.c_int], C.c_int),
    "dqa_canonize": ([_P, C.c_int, C.c_int], C.c_int),
    "dqa_truncate": ([_P, C.c_int, C.c_int], C.c_int),
    "dqa_apply_ux": ([_P, C.c_double], C.c_int),
    "dqa_energy": ([_P, _P], C.c_int),
    "dqa_get_bonds": ([_P, C.c_int, _P], C.c_int),
    "dqa_get_mps": ([_P, C.c_int, C.c_int, _P, C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
    "dqa_set_mps": ([_P, C.c_int, C.c_int, _P, C.c_int, C.c_int], C.c_int),
    "dqa_state_bytes": ([_P, C.POINTER(C.c_size_t)], C.c_int),
    "dqa_download_state": ([_P, _P], C.c_int),
    "dqa_upload_state": ([_P, _P], C.c_int),
    "dqa_sync": ([_P], C.c_int),
    "dqa_get_launch_count": ([_P, C.POINTER(C.c_int64)], C.c_int),
    "dqa_get_schmidt": ([_P, C.c_int, C.c_int, _P, C.POINTER(C.c_int)], C.c_int),
    "dqa_get_eps": ([_P, _P], C.c_int),
    "dqa_get_bd_monitor": ([_P, _P], C.c_int),
    "dqa_get_status": ([_P, _P], C.c_int),
    "dqa_get_counters": ([_P, _P, C.c_int], C.c_int),
    "dqa_ensemble_moments": ([_P, _P], C.c_int),
    "dqa_profile_begin": ([_P], C.c_int),
    "dqa_profile_end": ([_P, _P, _P], C.c_int),
    "dqa_measure_fp64_peak": ([_P, C.POINTER(C.c_double)], C.c_int),
    "dqa_measure_dmma_peak": ([_P, C.POINTER(C.c_double)], C.c_int),
    "dqa_version": ([], C.c_char_p),
    "dqa_ops_create": ([C.POINTER(_P), _P, C.c_size_t, C.c_int, _P], C.c_int),
    "dqa_ops_destroy": ([_P], None),
    "dqa_ops_last_error": ([_P], C.c_char_p),
    "dqa_op_contract_site": ([_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P], C.c_int),
    "dqa_op_braket": ([_P, C.c_int, _P, _P, _P, _P, _P], C.c_int),
    "dqa_op_qr_stabilized": ([_P, _P, C.c_int, C.c_int, _P, _P], C.c_int),
    "dqa_op_right_canonize_twosites": ([_P, _P, C.c_int, _P, C.c_int, C.c_int, _P, _P, C.POINTER(C.c_int)], C.c_int),
    "dqa_op_svd_truncated": ([_P, _P, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P,
                              C.POINTER(C.c_int)], C.c_int),
    "dqa_op_compress_onesite": ([_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P, C.POINTER(C.c_int)], C.c_int),
    "dqa_op_normalize": ([_P, _P, C.c_long], C.c_int),
    "dqa_op_energy": ([_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P], C.c_int),
    "dqa_sv_run": ([_P, C.c_int, C.c_int, _P, C.c_int, C.c_double, C.c_int, _P, _P], C.c_int),
}
EXPORTS = tuple(_SIGS)


def declare(lib):
    for name, (args, res) in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    return lib


_lib = None


def load():
    """Load the CUDA library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback for the dQA hot path.")
        _lib = declare(C.CDLL(LIB_PATH))
    return _lib


def _ptr(a):
    return a.ctypes.data_as(_P)


def _torch_workspace(nbytes, device):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the dQA hot path runs on the GPU only")
    t = torch.empty(nbytes + 256, dtype=torch.uint8, device=f"cuda:{device}")
    off = (-t.data_ptr()) % 256
    return t, t.data_ptr() + off


class Plan:
    """One dqa_t*: N sites, n_xi patterns, P steps, bond chi, `batch` realisations on one GPU."""

    def __init__(self, N, n_xi, P, dt, chi, batch=1, device=0, cutoff=0.0, max_sweeps=0, stream=None,
                 lib=None, workspace=None, n_tables=1, cutoff_mode=None, renorm=None, absorb=None):
        """cutoff_mode / renorm / absorb: the truncation policy of lib/tenn.py:238 (None = the reference's
        hard-coded 4 / 2 / 1); n_tables > 1: instances with different dt (set_fourier_tables)."""
        self.lib = lib or load()
        self.N, self.n_xi, self.P, self.dt, self.chi, self.batch, self.device = N, n_xi, P, dt, chi, batch, device
        self.D = N + 1
        self.n_tables = n_tables
        policy = not (cutoff_mode is None and renorm is None and absorb is None)
        cfg = dqa_config(N=N, n_xi=n_xi, P=P, chi=chi, batch=batch, device=device, max_sweeps=max_sweeps,
                         n_tables=n_tables, dt=dt, cutoff=cutoff, stream=stream, policy_set=int(policy),
                         cutoff_mode=4 if cutoff_mode is None else int(cutoff_mode),
                         renorm=2 if renorm is None else int(renorm), absorb=1 if absorb is None else int(absorb))
        self._h = _P()
        rc = self.lib.dqa_create(C.byref(self._h), C.byref(cfg))
        if rc != 0:
            msg = self.lib.dqa_last_error(self._h).decode() if self._h else "dqa_create failed"
            if self._h:
                self.lib.dqa_destroy(self._h)
                self._h = None
            raise DqaError(rc, msg)
        nbytes = C.c_size_t()
        self._ck(self.lib.dqa_workspace_bytes(self._h, C.byref(nbytes)))
        self.workspace_bytes = nbytes.value
        alloc = workspace or _torch_workspace
        self._ws_owner, ptr = alloc(self.workspace_bytes, device)
        self._ck(self.lib.dqa_bind_workspace(self._h, ptr, self.workspace_bytes))

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.dqa_last_error(self._h).decode()
            raise (DqaNumericError if rc == -4 else DqaError)(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dqa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # inputs --------------------------------------------------------------------------
    def set_patterns(self, xi):
        xi = np.ascontiguousarray(xi, dtype=np.int8).reshape(self.batch, self.n_xi, self.N)
        self._ck(self.lib.dqa_set_patterns(self._h, _ptr(xi)))

    def set_fourier(self, uk_ft, fxft):
        uk = np.ascontiguousarray(uk_ft, dtype=np.complex128)
        fx = np.ascontiguousarray(fxft, dtype=np.complex128)
        assert uk.shape == (self.D, self.P) and fx.shape == (self.D,)
        self._ck(self.lib.dqa_set_fourier(self._h, _ptr(uk), _ptr(fx)))

    def set_fourier_tables(self, uk_fts, fxft, table_of_instance, dt_of_instance):
        """uk_fts (n_tables, D, P): one init_fourier table per distinct dt; instance b uses table_of_instance[b]."""
        uk = np.ascontiguousarray(uk_fts, dtype=np.complex128)
        fx = np.ascontiguousarray(fxft, dtype=np.complex128)
        tab = np.ascontiguousarray(table_of_instance, dtype=np.int32)
        dts = np.ascontiguousarray(dt_of_instance, dtype=np.float64)
        assert uk.shape == (self.n_tables, self.D, self.P) and fx.shape == (self.D,)
        assert tab.shape == (self.batch,) and dts.shape == (self.batch,)
        self._ck(self.lib.dqa_set_fourier_tables(self._h, self.n_tables, _ptr(uk), _ptr(fx), _ptr(tab), _ptr(dts)))

    def index_tensors(self, inst=0):
        hbit = np.empty((self.n_xi, self.N), dtype=np.uint8)
        e = np.empty((self.n_xi, self.N, 2), dtype=np.int32)
        q = np.empty((self.n_xi, self.N, 2, self.D), dtype=np.int32)
        self._ck(self.lib.dqa_get_index_tensors(self._h, inst, _ptr(hbit), _ptr(e), _ptr(q)))
        return hbit, e, q

    # driver path -----------------------------------------------------------------------
    def reset_plus(self):
        self._ck(self.lib.dqa_reset_plus(self._h))

    def step(self, nsteps=1):
        self._ck(self.lib.dqa_step(self._h, nsteps))

    def step_async(self, nsteps=1):
        self._ck(self.lib.dqa_step_async(self._h, nsteps))

    @property
    def pp(self):
        v = C.c_int()
        self._ck(self.lib.dqa_get_step(self._h, C.byref(v)))
        return v.value

    @pp.setter
    def pp(self, v):
        self._ck(self.lib.dqa_set_step(self._h, int(v)))

    # sub-steps ---------------------------------------------------------------------------
    def apply_pattern(self, p, mu):
        self._ck(self.lib.dqa_apply_pattern(self._h, p, mu))

    def canonize(self, p, mu):
        self._ck(self.lib.dqa_canonize(self._h, p, mu))

    def truncate(self, p, mu):
        self._ck(self.lib.dqa_truncate(self._h, p, mu))

    def apply_ux(self, beta):
        self._ck(self.lib.dqa_apply_ux(self._h, float(beta)))

    def energy(self):
        out = np.empty(self.batch, dtype=np.complex128)
        self._ck(self.lib.dqa_energy(self._h, _ptr(out)))
        return out

    # state -------------------------------------------------------------------------------
    def bonds(self, inst=0):
        b = np.empty(self.N + 1, dtype=np.int32)
        self._ck(self.lib.dqa_get_bonds(self._h, inst, _ptr(b)))
        return b

    def get_mps(self, inst=0):
        out = []
        buf = np.empty(2 * self.chi * self.chi, dtype=np.complex128)
        l, r = C.c_int(), C.c_int()
        for site in range(self.N):
            self._ck(self.lib.dqa_get_mps(self._h, inst, site, _ptr(buf), C.byref(l), C.byref(r)))
            out.append(buf[: l.value * 2 * r.value].reshape(l.value, 2, r.value).copy())
        return out

    def set_mps(self, mps, inst=0):
        assert len(mps) == self.N
        for site, t in enumerate(mps):
            t = np.ascontiguousarray(t, dtype=np.complex128)
            assert t.ndim == 3 and t.shape[1] == 2
            self._ck(self.lib.dqa_set_mps(self._h, inst, site, _ptr(t), t.shape[0], t.shape[2]))

    def state_bytes(self):
        v = C.c_size_t()
        self._ck(self.lib.dqa_state_bytes(self._h, C.byref(v)))
        return v.value

    def download_state(self, host_ptr):
        self._ck(self.lib.dqa_download_state(self._h, host_ptr))

    def upload_state(self, host_ptr):
        self._ck(self.lib.dqa_upload_state(self._h, host_ptr))

    def sync(self):
        self._ck(self.lib.dqa_sync(self._h))

    def launch_count(self):
        v = C.c_int64()
        self._ck(self.lib.dqa_get_launch_count(self._h, C.byref(v)))
        return v.value

    def schmidt(self, bond, inst=0):
        s = np.empty(2 * self.chi, dtype=np.float64)
        n = C.c_int()
        self._ck(self.lib.dqa_get_schmidt(self._h, inst, bond, _ptr(s), C.byref(n)))
        return s[: n.value].copy()

    # traces ------------------------------------------------------------------------------
    def eps_trace(self):
        out = np.empty((self.batch, self.P + 1), dtype=np.complex128)
        self._ck(self.lib.dqa_get_eps(self._h, _ptr(out)))
        return out

    def bd_monitor(self):
        out = np.empty((self.batch, self.P * self.n_xi), dtype=np.int32)
        self._ck(self.lib.dqa_get_bd_monitor(self._h, _ptr(out)))
        return out

    def status(self):
        out = np.empty(self.batch, dtype=np.int32)
        self._ck(self.lib.dqa_get_status(self._h, _ptr(out)))
        return out

    def clear_status(self):
        self._ck(self.lib.dqa_clear_status(self._h))

    def truncation_error(self, inst=0):
        """relative discarded weight of the last truncation at every cut (N-1 values)"""
        out = np.empty(self.N - 1, dtype=np.float64)
        self._ck(self.lib.dqa_get_truncation_error(self._h, inst, _ptr(out)))
        return out

    def counters(self, reset=False):
        out = np.empty((self.batch, 16), dtype=np.uint64)
        self._ck(self.lib.dqa_get_counters(self._h, _ptr(out), int(reset)))
        return out

    def ensemble_moments_into(self, device_ptr):
        self._ck(self.lib.dqa_ensemble_moments(self._h, device_ptr))

    def profile_begin(self):
        self._ck(self.lib.dqa_profile_begin(self._h))

    def profile_end(self):
        ms = np.zeros(NFAM, dtype=np.float64)
        n = np.zeros(NFAM, dtype=np.int64)
        self._ck(self.lib.dqa_profile_end(self._h, _ptr(ms), _ptr(n)))
        return dict(zip(FAMILIES, ms.tolist())), dict(zip(FAMILIES, n.tolist()))

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._ck(self.lib.dqa_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def dmma_peak_tflops(self):
        v = C.c_double()
        self._ck(self.lib.dqa_measure_dmma_peak(self._h, C.byref(v)))
        return v.value


class Ops:
    """Op-level context (include/dqa.h, "op-level API"): the reference's lib/tenn.py functions on dense
    site tensors, computed on the GPU.  Owns nothing but a caller-allocated device scratch."""

    def __init__(self, scratch_bytes=256 << 20, device=0, stream=None, lib=None, workspace=None):
        self.lib = lib or load()
        alloc = workspace or _torch_workspace
        self._owner, ptr = alloc(scratch_bytes, device)
        self._h = _P()
        rc = self.lib.dqa_ops_create(C.byref(self._h), ptr, scratch_bytes, device, stream)
        if rc != 0:
            raise DqaError(rc, self.lib.dqa_ops_last_error(self._h).decode() if self._h else "dqa_ops_create failed")

    def _ck(self, rc):
        if rc != 0:
            raise (DqaNumericError if rc == -4 else DqaError)(rc, self.lib.dqa_ops_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dqa_ops_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _c(a):
        return np.ascontiguousarray(a, dtype=np.complex128)

    def contract_site(self, a, w):
        a, w = self._c(a), self._c(w)
        al, _, ar = a.shape
        wl, wr = w.shape[0], w.shape[1]
        out = np.empty((al * wl, 2, ar * wr), dtype=np.complex128)
        self._ck(self.lib.dqa_op_contract_site(self._h, _ptr(a), al, ar, _ptr(w), wl, wr, _ptr(out)))
        return out

    @staticmethod
    def _pack(mps):
        bonds = np.array([mps[0].shape[0]] + [t.shape[2] for t in mps], dtype=np.int32)
        flat = np.concatenate([np.ascontiguousarray(t, dtype=np.complex128).ravel() for t in mps])
        return bonds, flat

    def braket(self, bra, ket):
        bb, bf = self._pack(bra)
        kb, kf = self._pack(ket)
        out = np.empty((int(bb[-1]), int(kb[-1])), dtype=np.complex128)
        self._ck(self.lib.dqa_op_braket(self._h, len(bra), _ptr(bb), _ptr(bf), _ptr(kb), _ptr(kf), _ptr(out)))
        return out

    def qr_stabilized(self, x):
        x = self._c(x)
        m, n = x.shape
        k = min(m, n)
        q = np.empty((m, k), dtype=np.complex128)
        r = np.empty((k, n), dtype=np.complex128)
        self._ck(self.lib.dqa_op_qr_stabilized(self._h, _ptr(x), m, n, _ptr(q), _ptr(r)))
        return q, r

    def right_canonize_twosites(self, a, b):
        a, b = self._c(a), self._c(b)
        al, bl, br = a.shape[0], b.shape[0], b.shape[2]
        kq = min(2 * br, bl)
        an = np.empty((al, 2, kq), dtype=np.complex128)
        bn = np.empty((kq, 2, br), dtype=np.complex128)
        k = C.c_int()
        self._ck(self.lib.dqa_op_right_canonize_twosites(self._h, _ptr(a), al, _ptr(b), bl, br, _ptr(an), _ptr(bn), C.byref(k)))
        return an, bn

    def svd_truncated(self, x, cutoff=-1.0, cutoff_mode=3, max_bond=-1, absorb=0, renorm=0):
        x = self._c(x)
        m, n = x.shape
        k = min(m, n)
        u = np.empty((m, k), dtype=np.complex128)
        s = np.empty(k, dtype=np.float64)
        vh = np.empty((k, n), dtype=np.complex128)
        nc = C.c_int()
        ab = 2 if absorb is None else int(absorb)
        self._ck(self.lib.dqa_op_svd_truncated(self._h, _ptr(x), m, n, float(cutoff), int(cutoff_mode), int(max_bond), ab, int(renorm),
                                               _ptr(u), _ptr(s), _ptr(vh), C.byref(nc)))
        c = nc.value
        return u[:, :c].copy(), s[:c].copy(), vh[:c].copy()

    def compress_onesite(self, a, a_next, max_bd):
        a, a_next = self._c(a), self._c(a_next)
        l, _, r = a.shape
        r2 = a_next.shape[2]
        k = min(2 * l, r)
        ao = np.empty(2 * l * k, dtype=np.complex128)
        no = np.empty(k * 2 * r2, dtype=np.complex128)
        nc = C.c_int()
        self._ck(self.lib.dqa_op_compress_onesite(self._h, _ptr(a), l, r, _ptr(a_next), r2, int(max_bd), _ptr(ao), _ptr(no), C.byref(nc)))
        c = nc.value
        return ao[: 2 * l * c].reshape(l, 2, c).copy(), no[: c * 2 * r2].reshape(c, 2, r2).copy()

    def normalize(self, x):
        x = self._c(x).copy()
        self._ck(self.lib.dqa_op_normalize(self._h, _ptr(x), x.size))
        return x

    def energy(self, mps, xi, fxft, n_ancilla=0):
        bonds, flat = self._pack(mps)
        xi = np.ascontiguousarray(xi, dtype=np.int8)
        fx = np.ascontiguousarray(fxft, dtype=np.complex128)
        out = np.empty(1, dtype=np.complex128)
        n_phys = len(mps) - n_ancilla
        assert xi.shape[1] == n_phys and fx.shape[0] == n_phys + 1
        self._ck(self.lib.dqa_op_energy(self._h, len(mps), n_phys, xi.shape[0], _ptr(bonds), _ptr(flat), _ptr(xi), _ptr(fx), _ptr(out)))
        return complex(out[0])

    def statevector_dqa(self, xi, P, dt, steps=None, timings=False):
        """Exact Trotterised dQA on the device (2^N amplitudes): eps(s) for s = 0..steps/P."""
        xi = np.ascontiguousarray(xi, dtype=np.int8)
        n_xi, n = xi.shape
        steps = P if steps is None else steps
        eps = np.empty(steps + 1, dtype=np.float64)
        ms = np.zeros(3, dtype=np.float64)
        self._ck(self.lib.dqa_sv_run(self._h, n, n_xi, _ptr(xi), int(P), float(dt), int(steps), _ptr(eps), _ptr(ms) if timings else None))
        return (eps, dict(zip(("phase", "mixer", "energy"), ms.tolist()))) if timings else eps
