"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):
    OMP_NUM_THREADS=1 python oracle/make_golden.py
Everything stored here was produced by the reference's own files (dQA_mps.py, lib/tenn.py,
lib/dQA_utils.py) executed verbatim behind the import shim (oracle/ref_runner.py) in
complex128, with the sign(0):=+1 shim switch on unless the key says
`verbatim`.  The fixtures travel to the GPU box, where /root/reference does not exist.

Per config the file holds: the +-1 patterns, the Fourier tables, the integer Hamming
tables computed by the SURVEY 8a-2 formulas, eps(s), bd_monitor, and MPS snapshots
(before and after selected sub-steps, for teacher-forced parity).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_runner  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
DATA = os.path.join(ref_runner.REFERENCE_ROOT, "data")


def pack_mps(mps):
    """list of (l,2,r) -> (bonds[N+1], flat complex array)"""
    bonds = [mps[0].shape[0]] + [t.shape[2] for t in mps]
    return np.array(bonds, dtype=np.int32), np.concatenate([np.asarray(t).ravel() for t in mps])


def reference_substep(mod, obj, psi, mu):
    """The body of the mu-loop of mydQA.single_step (dQA_mps.py:94-103), calling the
    reference's own functions."""
    uz = mod.PerceptronHamiltonian.make_Uz(obj.N, obj.Uk_FT[:, obj.pp], obj.dataset[mu])
    psi = mod.apply_mpsmpo(psi, uz)
    pre = mod.right_canonize(psi, 1)
    return mod.compress_svd_normalized(pre, max_bd=obj.max_bond)


def run_config(name, xi, P, dt, chi, steps, snap_at, sgn0fix=True, svd_driver="gesdd", extra=None):
    mod = ref_runner.load(sgn0fix=sgn0fix, svd_driver=svd_driver)
    obj = mod.mydQA(xi, P=P, dt=dt, max_bond=chi)
    obj.init_fourier()
    out = {
        "xi": np.asarray(xi).astype(np.int8), "P": P, "dt": dt, "chi": chi, "steps": steps,
        "Uk_FT": obj.Uk_FT, "fxft": obj.fxft,
    }
    sgn = np.array([1, -1], dtype=np.int64)
    xi64 = np.asarray(xi).astype(np.int64)
    e = np.rint(1 - xi64[:, :, None] * sgn[None, None, :]).astype(np.int64)
    out["e"] = e
    out["q"] = e[:, :, :, None] * np.arange(obj.N + 1, dtype=np.int64)[None, None, None, :]
    out["hbit"] = ((1 - xi64) // 2).astype(np.uint8)
    t0 = time.time()
    with ref_runner.quiet():
        obj.psi = [np.array([[[2**-0.5], [2**-0.5]]], dtype=np.complex128)] * obj.N
        obj.pp = 0
        obj.loss = [(0, obj.compute_loss(obj.psi))]
        obj.bd_monitor = []
        snaps = []
        for p in range(steps):
            if p in snap_at:
                # replay this step by hand with the reference's functions to capture sub-step states
                psi = obj.psi
                for mu in range(obj.N_xi):
                    if mu in snap_at[p]:
                        b0, f0 = pack_mps(psi)
                    psi = reference_substep(mod, obj, [np.array(t) for t in psi], mu)
                    if mu in snap_at[p]:
                        b1, f1 = pack_mps(psi)
                        snaps.append((p, mu, b0, f0, b1, f1))
            obj.single_step()
            if p in snap_at:
                b, f = pack_mps(obj.psi)
                out[f"state_after_step{p}_bonds"], out[f"state_after_step{p}_flat"] = b, f
    out["eps"] = np.array([l[1] for l in obj.loss])
    out["s"] = np.array([l[0] for l in obj.loss], dtype=np.float64)
    out["bd_monitor"] = np.array(obj.bd_monitor, dtype=np.int32)
    out["snap_index"] = np.array([(p, mu) for p, mu, *_ in snaps], dtype=np.int32).reshape(-1, 2)
    for i, (_p, _mu, b0, f0, b1, f1) in enumerate(snaps):
        out[f"snap{i}_pre_bonds"], out[f"snap{i}_pre_flat"] = b0, f0
        out[f"snap{i}_post_bonds"], out[f"snap{i}_post_flat"] = b1, f1
    if extra:
        out.update(extra)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: {steps} steps in {time.time() - t0:.1f}s, eps[-1]={out['eps'][-1]:.12g}", flush=True)
    return out


def run_saturated_substep(name, xi, P, dt, chi, warm_substeps):
    """One teacher-forced sub-step on a state whose bonds have reached chi: the reference's own sub-step functions are
    applied `warm_substeps` times from |+> (no energy evaluations: the reference's compute_loss builds chi^4 tensors), then
    the next sub-step is stored as (pre, post)."""
    mod = ref_runner.load(sgn0fix=True, svd_driver="gesdd")
    obj = mod.mydQA(xi, P=P, dt=dt, max_bond=chi)
    obj.init_fourier()
    t0 = time.time()
    psi = [np.array([[[2**-0.5], [2**-0.5]]], dtype=np.complex128)] * obj.N
    p = mu = 0
    with ref_runner.quiet():
        for _ in range(warm_substeps):
            obj.pp = p
            psi = reference_substep(mod, obj, [np.array(t) for t in psi], mu)
            mu += 1
            if mu == obj.N_xi:
                # the mixer of the reference's single_step (dQA_mps.py:108-109)
                beta = (1.0 - (p + 1) / P) * dt
                psi = mod.apply_mpsmpo(psi, mod.PerceptronHamiltonian.make_Ux(obj.N, beta_p=beta))
                mu, p = 0, p + 1
        obj.pp = p
        b0, f0 = pack_mps(psi)
        post = reference_substep(mod, obj, [np.array(t) for t in psi], mu)
        b1, f1 = pack_mps(post)
    xi64 = np.asarray(xi).astype(np.int64)
    out = {"xi": np.asarray(xi).astype(np.int8), "P": P, "dt": dt, "chi": chi, "steps": 0, "Uk_FT": obj.Uk_FT, "fxft": obj.fxft,
           "hbit": ((1 - xi64) // 2).astype(np.uint8), "snap_index": np.array([(p, mu)], dtype=np.int32),
           "snap0_pre_bonds": b0, "snap0_pre_flat": f0, "snap0_post_bonds": b1, "snap0_post_flat": f1}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: {warm_substeps} warm-up sub-steps in {time.time() - t0:.1f}s, bonds {b0.tolist()} -> {b1.tolist()}", flush=True)


def main():
    which = sys.argv[1:] or ["n7", "c1", "c1_spread", "c2", "n15"]
    if "n7" in which:
        xi = np.load(os.path.join(DATA, "patterns_5-7.npy"))
        run_config("n7_chi8_untruncated", xi, P=20, dt=0.7, chi=8, steps=20, snap_at={0: {0, 4}, 10: {2}})
        run_config("n7_chi4", xi, P=20, dt=0.7, chi=4, steps=20, snap_at={0: {0, 4}, 7: {1, 3}})
    if "c1" in which:
        xi = np.load(os.path.join(DATA, "patterns_9-12.npy"))
        run_config("c1_n12_chi10", xi, P=100, dt=0.5, chi=10, steps=100,
                   snap_at={0: {0, 8}, 5: set(range(9)), 30: {0, 4}, 60: {3}, 95: {8}})
    if "c1_spread" in which:
        # the oracle's own roundoff envelope: same run with LAPACK gesvd instead of gesdd
        xi = np.load(os.path.join(DATA, "patterns_9-12.npy"))
        run_config("c1_n12_chi10_gesvd", xi, P=100, dt=0.5, chi=10, steps=100, snap_at={}, svd_driver="gesvd")
        ref_runner.load(svd_driver="gesdd")
    if "c2" in which:
        xi = np.load(os.path.join(DATA, "patterns_17-21.1.npy"))
        run_config("c2_n21_chi10", xi, P=100, dt=1.0, chi=10, steps=24, snap_at={0: {0}, 3: {0, 8, 16}, 20: {5}})
        run_config("c2_n21_chi10_gesvd", xi, P=100, dt=1.0, chi=10, steps=24, snap_at={}, svd_driver="gesvd")
        ref_runner.load(svd_driver="gesdd")
    if "n15" in which:
        xi = np.load(os.path.join(DATA, "patterns_12-15.npy"))
        run_config("n15_chi10", xi, P=100, dt=1.2, chi=10, steps=12, snap_at={4: {0, 11}})
    if "c3" in which:
        # N=21 at the headline bond dimension: a few steps only (the reference needs minutes per step)
        xi = np.load(os.path.join(DATA, "patterns_17-21.1.npy"))
        run_config("c3_n21_chi32", xi, P=100, dt=1.0, chi=32, steps=2, snap_at={1: {0, 16}})
    if "n12_chi64" in which:
        # untruncated regime at the largest supported bond: N=12, chi=64 = 2^(N/2)
        xi = np.load(os.path.join(DATA, "patterns_9-12.npy"))
        run_config("n12_chi64_untruncated", xi, P=100, dt=0.5, chi=64, steps=5, snap_at={4: {0, 8}})
    if "n32" in which:
        # scaled-down analogue of the N=64 instance: synthetic iid +-1 patterns, default_rng(64)
        rng = np.random.default_rng(64)
        xi = (2 * rng.integers(0, 2, size=(4, 32)) - 1).astype(np.int8)
        run_config("n32_chi8_synthetic", xi, P=50, dt=1.0, chi=8, steps=3, snap_at={2: {0, 3}})
    if "n14_chi64" in which:
        # saturated bonds at the largest supported chi (the 2chi = 128-column SVD): synthetic iid +-1 patterns, default_rng(14)
        rng = np.random.default_rng(14)
        xi = (2 * rng.integers(0, 2, size=(int(os.environ.get("NXI", "8")), 14)) - 1).astype(np.int8)
        run_saturated_substep("n14_chi64_saturated", xi, P=20, dt=1.0, chi=64, warm_substeps=int(os.environ.get("WARM", "32")))
    if "n12_chi48" in which:
        # saturated bonds in the 12-slot sector-QR / shared-memory-resident Jacobi range (32 < chi <= 56), default_rng(48)
        rng = np.random.default_rng(48)
        xi = (2 * rng.integers(0, 2, size=(8, 12)) - 1).astype(np.int8)
        run_saturated_substep("n12_chi48_saturated", xi, P=20, dt=1.0, chi=48, warm_substeps=32)
    if "c3_16" in which:
        xi = np.load(os.path.join(DATA, "patterns_17-21.1.npy"))
        run_config("c3_n21_chi16", xi, P=100, dt=1.0, chi=16, steps=4, snap_at={3: {0, 16}})


if __name__ == "__main__":
    main()
