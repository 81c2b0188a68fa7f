#!/usr/bin/env python
"""bench.py -- dQA steps/s at N=21, chi=32 (BASELINE.json's metric) on N GPUs of one node.

    python bench.py --gpus 1 --steps K --warmup W              # this repo (CUDA, libdqa.so)
    python bench.py --impl reference --gpus 1 --steps K ...    # the reference algorithm on the host CPUs
    torchrun ... bench.py --gpus N ...                         # one rank per GPU, ensemble sharded

A "step" is one dQA step (N_xi x {U_z^mu, right-canonise, truncate to chi} + U_x + energy,
dQA_mps.py:84-119) of every realisation in the job; `value` = realisation-steps per second,
whole job.  Workload: an ensemble of `--batch` independent pattern realisations per GPU
(realisation 0 is data/patterns_17-21.1.npy, the rest follow the C4 law of SURVEY 8d),
N=21, N_xi=17, chi=32, P=1000, dt=1.0, steps taken at p ~ P/2 with saturated bonds.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

METRIC = "dQA MPO-apply+SVD-truncate steps/sec at N=21, chi=32"
UNIT = "realisation-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=21)
    ap.add_argument("--nxi", type=int, default=17)
    ap.add_argument("--chi", type=int, default=32)
    ap.add_argument("--P", type=int, default=1000)
    ap.add_argument("--dt", type=float, default=1.0)
    ap.add_argument("--batch", type=int, default=888, help="realisations per GPU (888 = 148 SMs x 6: whole waves at 2 and 3 CTAs/SM)")
    ap.add_argument("--cpu-seconds", type=float, default=200.0, help="budget of the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    return ap.parse_args()


def ensemble_patterns(a, rank, batch):
    """C4 law: iid uniform +-1, default_rng(20230317); realisation r of the job is
    row r of that table; global realisation 0 is the reference's patterns_17-21.1.npy."""
    rng = np.random.default_rng(20230317)
    table = (2 * rng.integers(0, 2, size=(1024, 17, 21)) - 1).astype(np.int8)
    if (a.n, a.nxi) != (21, 17):
        table = (2 * rng.integers(0, 2, size=(1024, a.nxi, a.n)) - 1).astype(np.int8)
    idx = (rank * batch + np.arange(batch)) % 1024
    xi = table[idx].copy()
    if rank == 0 and (a.n, a.nxi) == (21, 17):
        g = np.load(os.path.join(ROOT, "tests", "golden", "c2_n21_chi10.npz"))
        xi[0] = g["xi"]
    return xi


def workload_name(a, batch, world):
    return (f"ensemble of {batch * world} pattern realisations ({batch}/GPU; #0 = patterns_17-21.1.npy, rest iid +-1 "
            f"rng 20230317), N={a.n}, N_xi={a.nxi}, chi={a.chi}, P={a.P}, dt={a.dt}, complex128, steps at p~P/2")


# ----------------------------------------------------------------------------------------
# CPU legs: the oracle (numpy restatement of the reference algorithm) on the host cores
# ----------------------------------------------------------------------------------------
def _random_saturated_mps(n, chi, seed):
    """A normalised random MPS with the bond profile min(chi, 2^i, 2^(N-i)) -- the shape the
    state has at p ~ P/2; LAPACK cost does not depend on the values."""
    rng = np.random.default_rng(seed)
    bonds = [min(chi, 2 ** i, 2 ** (n - i)) for i in range(n + 1)]
    mps = [rng.normal(size=(bonds[i], 2, bonds[i + 1])) + 1j * rng.normal(size=(bonds[i], 2, bonds[i + 1])) for i in range(n)]
    return mps


def _cpu_worker(args):
    """One sample on one core: one pattern sub-step (U_z apply, right-canonise, truncate:
    dQA_mps.py:94-103) plus the energy of two of the N_xi patterns (dQA_mps.py:68-81)."""
    (n, nxi, chi, P, dt, seed, mps) = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dqa_oracle as O

    rng = np.random.default_rng(seed)
    xi = (2 * rng.integers(0, 2, size=(nxi, n)) - 1).astype(np.int8)
    o = O.OracleDQA(xi, P, dt, chi)
    o.init_fourier()
    o.pp = P // 2
    if mps is None:
        mps = _random_saturated_mps(n, chi, seed)
        mps = O.right_canonize(mps, 1)
        mps = O.compress_svd_normalized(mps, chi)
    t0 = time.perf_counter()
    psi = o.apply_pattern([t.copy() for t in mps], 0)
    t1 = time.perf_counter()
    O.compute_loss(psi, xi[:2], o.fxft)
    t2 = time.perf_counter()
    return (t1 - t0, (t2 - t1) / 2.0)


def cpu_sample(a, cores, state=None):
    """Run one sample on each of `cores` processes at once; returns the estimated seconds one
    full dQA step takes on one core = N_xi * (t_substep + t_energy_per_pattern)."""
    import multiprocessing as mp

    jobs = [(a.n, a.nxi, a.chi, a.P, a.dt, 1000 + i, state) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    per_step = [a.nxi * (ts + te) for ts, te in res]
    return float(np.mean(per_step)), wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(a):
    """--impl reference: the reference's algorithm (oracle port; the reference's own Python files
    cannot travel to the GPU box) on all host cores, one realisation per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(host_cores(), 64)
    # calibrate with one warm-up sample, then size K so the run stays inside the budget
    t_step, wall = cpu_sample(a, cores)
    k = max(1, min(a.steps, int(max(1.0, (a.cpu_seconds - wall)) / max(wall, 1e-3))))
    steps = []
    for _ in range(k):
        t_step, wall = cpu_sample(a, cores)
        steps.append(t_step)
    t_step = float(np.mean(steps))
    value = cores / t_step
    sample = (f"{k} x (1 of {a.nxi} pattern sub-steps + energy of 2 of {a.nxi} patterns) per core on a random "
              f"saturated MPS, scaled by {a.nxi}; {cores} processes, 1 BLAS thread each")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": k,
        "warmup": 1, "ms_per_step": 1e3 * t_step / cores, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": {"workload": workload_name(a, a.batch, a.gpus), "note": "CPU arm: one realisation per host core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 7 and r[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def run_ours(a):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the dQA hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from perceptron_dqa_b200.dQA_mps import mydQA

    B = a.batch
    xi = ensemble_patterns(a, rank, B)
    obj = mydQA(xi if B > 1 else xi[0], P=a.P, dt=a.dt, max_bond=a.chi, device=local)
    obj.init_fourier()
    pl = obj.plan
    pl.reset_plus()
    pl.pp = a.P // 2
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(3, a.warmup)
    pl.step(W)
    barrier()

    # ---- timed region: K steps, device time, L2 flushed between steps ---------------------
    sampler = ClockSampler(local)
    sampler.start()
    l0 = pl.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    for k in range(a.steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        pl.step(1)
        ev[k][1].record()
    barrier()
    ms = sum(x.elapsed_time(y) for x, y in ev)
    launches = pl.launch_count() - l0
    clocks = sampler.summary()
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = B * world * a.steps / (ms_max * 1e-3)

    # ---- ensemble reduction (the one collective of the path): sums of eps(s) over all ranks --
    mom = torch.zeros(3 * (a.P + 1), dtype=torch.float64, device=f"cuda:{local}")
    pl.ensemble_moments_into(mom.data_ptr())
    if dist is not None:
        dist.all_reduce(mom)
    torch.cuda.synchronize()
    s_idx = a.P // 2 + W + a.steps
    n_ok = float(mom[2 * (a.P + 1) + s_idx].item())
    eps_mean = float(mom[s_idx].item()) / max(n_ok, 1.0)

    # ---- end to end: host buffers in, host buffers out, every step ---------------------------
    nbytes = pl.state_bytes()
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    eps_host = np.empty((B, a.P + 1), dtype=np.complex128)
    pl.download_state(host.data_ptr())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_e2e = max(1, min(a.steps, 3))
    t0 = time.perf_counter()
    e0.record()
    for _ in range(k_e2e):
        pl.set_patterns(xi)                      # H2D: the step's patterns
        pl.upload_state(host.data_ptr())         # H2D: the step's input state (pinned)
        pl.step(1)
        pl.download_state(host.data_ptr())       # D2H: the new state
        eps_host[:] = pl.eps_trace()             # D2H: eps(s) trace
    e1.record()
    barrier()
    t_e2e = time.perf_counter() - t0
    ms_e2e = e0.elapsed_time(e1)
    t2 = torch.tensor([max(ms_e2e * 1e-3, t_e2e)], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = B * world * k_e2e / float(t2.item())
    h2d = nbytes + xi.nbytes
    d2h = nbytes + eps_host.nbytes + 4 * B

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
        "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex128", "data": "synthetic",
        "config": {"workload": workload_name(a, B, world), "batch_per_gpu": B,
                   "l2": "flushed (256 MiB write) between timed steps; per-step working set "
                         f"{pl.workspace_bytes / 2**20:.0f} MiB/GPU", "parallelism": f"ensemble x{world}"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": k_e2e},
        "gpu_launches": int(launches), "clocks": clocks,
        "single_realisation": None,
        "ensemble": {"eps_mean_at_last_step": eps_mean, "healthy": n_ok, "allreduce": "nccl" if dist is not None else "none"},
        "status_nonzero": int(np.count_nonzero(pl.status())),
    }

    if rank == 0 and not a.no_roofline:
        # per-kernel-family device time (CUDA events inside libdqa on the launching stream) and the
        # FP64 flops each family executed (exact loop counts kept by the kernels themselves)
        pl.counters(reset=True)
        pl.profile_begin()
        pl.step(1)
        fam_ms, fam_n = pl.profile_end()
        cnt = pl.counters().sum(0).astype(np.float64)
        peak = pl.fp64_peak_tflops()
        bonds = pl.bonds(0)
        e_fl = sum(16.0 * (bonds[i] ** 2 * bonds[i + 1] + bonds[i] * bonds[i + 1] ** 2) for i in range(a.n))
        flops = {"jacobi": cnt[5], "lq": cnt[4], "sector_qr": cnt[6], "theta": cnt[7],
                 "energy": float(e_fl) * a.nxi * (a.n // 2 + 1 + (a.n + 1) % 2 * 0) * B}
        fams = {}
        for f, fl in flops.items():
            ms_f = fam_ms.get(f, 0.0)
            ach = fl / (ms_f * 1e-3) / 1e12 if ms_f > 0 else 0.0
            fams[f] = {"ms_per_step": ms_f, "launches": fam_n.get(f, 0), "flops_per_step": fl, "achieved_tflops": ach,
                       "frac": ach / peak if peak else None, "share_of_step": ms_f / max(1e-9, sum(fam_ms.values()))}
        dom = max(fams, key=lambda f: fams[f]["ms_per_step"])
        line["roofline"] = {
            "kernel": {"jacobi": "k_jacobi", "lq": "k_lq", "sector_qr": "k_sector_qr", "theta": "k_theta", "energy": "k_energy"}[dom],
            "bound": "fp64", "achieved": fams[dom]["achieved_tflops"], "peak": peak, "unit": "TFLOP/s",
            "frac": fams[dom]["frac"], "traffic": None,
            "peak_source": "measured on this GPU by libdqa's DFMA probe (MEASURED_PEAKS.json has no FP64 entry; "
                           "B200 FP64 vector = FP64 tensor peak)",
            "flops_per_launch": fams[dom]["flops_per_step"] / max(1, fams[dom]["launches"]),
            "ms_per_launch": fams[dom]["ms_per_step"] / max(1, fams[dom]["launches"]),
            "share_of_step": fams[dom]["share_of_step"], "families": fams,
            "whole_step": {"flops": float(sum(flops.values())), "achieved_tflops": float(sum(flops.values())) / (sum(fam_ms.values()) * 1e-3) / 1e12,
                           "frac": float(sum(flops.values())) / (sum(fam_ms.values()) * 1e-3) / 1e12 / peak if peak else None},
            "rotations_per_svd": float(cnt[0]) / max(1.0, float(cnt[2])), "sweeps_per_svd": float(cnt[1]) / max(1.0, float(cnt[2])),
            # what the reference's own (dense) algorithm would execute for the same steps: context for the speed-up, not a
            # roofline claim (the kernels execute ~100x fewer flops in the sector representation)
            "reference_algorithm": {"flops_per_realisation_step": reference_flops(a.n, a.nxi, a.chi),
                                    "executed_flops_per_realisation_step": float(sum(flops.values())) / B},
        }
        # DRAM traffic of the dominant kernel: not measurable here without a profiler, so it is read from the committed ncu
        # capture of the same kernel (one mid-chain launch at this batch), when that summary is present
        try:
            tr = 0.0
            with open(os.path.join(ROOT, "profiles", f"r01_{line['roofline']['kernel']}_ncu_summary.csv")) as f:
                for ln in f:
                    parts = ln.strip().split(",")
                    if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        tr += float(parts[2]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[parts[1]]
            if tr > 0:
                line["roofline"]["traffic"] = tr
                line["roofline"]["traffic_source"] = (f"profiles/r01_{line['roofline']['kernel']}_ncu_summary.csv: dram read+write of one "
                                                      "launch under ncu --set full (batch 888); the kernel is FP64-bound, not HBM-bound")
        except Exception:
            pass
        # the HBM side of the roofline, to show the step is not bandwidth bound: algorithmic bytes = the MPS read and
        # written once per pattern sub-step and once for mixer+energy, against the pool's measured copy bandwidth
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
        s_bytes = sum(16 * 2 * int(bonds[i]) * int(bonds[i + 1]) for i in range(a.n))
        alg_bytes = 2.0 * (a.nxi + 1) * s_bytes * B
        step_s = sum(fam_ms.values()) * 1e-3
        line["roofline"]["hbm_view"] = {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": alg_bytes / step_s / 1e9,
                                        "peak_gbs": hbm_peak, "peak_source": hbm_src, "frac": alg_bytes / step_s / 1e9 / hbm_peak}
    if rank == 0 and B > 1:
        # latency view: one realisation alone on the GPU
        o1 = mydQA(xi[0], P=a.P, dt=a.dt, max_bond=a.chi, device=local)
        o1.init_fourier()
        p1 = o1.plan
        p1.reset_plus()
        p1.pp = a.P // 2
        p1.step(3)
        torch.cuda.synchronize()
        x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x.record()
        p1.step(2)
        y.record()
        torch.cuda.synchronize()
        line["single_realisation"] = {"steps_per_s": 2.0 / (x.elapsed_time(y) * 1e-3), "ms_per_step": x.elapsed_time(y) / 2.0}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = min(host_cores(), 64)
        state = None
        t_step, wall = cpu_sample(a, cores, state)
        line["cpu_baseline"] = {
            "value": cores / t_step, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 of {a.nxi} pattern sub-steps + energy of 2 of {a.nxi} patterns per core on a random saturated MPS, "
                      f"scaled by {a.nxi}; {cores} processes x 1 BLAS thread; {wall:.0f} s wall",
            "s_per_step_per_core": t_step}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)



def reference_flops(N, n_xi, chi):
    """real FP64 flops of one step of the REFERENCE's algorithm (dense bond chi*(N+1) tensors,
    LAPACK operation counts: QR of every enlarged site, R-absorb GEMM, U_z apply, SVD, (S V^H) A GEMM, energy).
    1 226 GFLOP at N=21, chi=32.  Reported next to the flops the kernels execute; never used for `achieved`."""
    D = N + 1
    b = [min(chi, 2 ** i, 2 ** (N - i)) for i in range(N + 1)]
    m = [1] + [b[n] * D for n in range(1, N)] + [1]

    def geqrf(M, n):
        return 2 * M * n * n - 2 / 3 * n ** 3 if M >= n else 2 * n * M * M - 2 / 3 * M ** 3

    q = [0] * (N + 1)
    q[N] = 1
    r, qr, gemm = 1, 0.0, 0.0
    for n in range(N - 1, 0, -1):
        M = 2 * r
        k = min(M, m[n])
        qr += 4 * (geqrf(M, m[n]) + 2 * M * k * k - 2 / 3 * k ** 3)
        q[n] = k
        gemm += 8 * (2 * b[n - 1] * b[n] * D) * q[n]
        r = k
    uz = sum(6 * 2 * b[i] * b[i + 1] * D for i in range(N))
    svd, absorb, c = 0.0, 0.0, 1
    for l in range(N - 1):
        mx, mn = max(2 * c, q[l + 1]), min(2 * c, q[l + 1])
        svd += 4 * (2 * (2 * mx * mn * mn - 2 / 3 * mn ** 3) + 20 * mn ** 3)
        c = min(chi, 2 * c, q[l + 1])
        absorb += 8 * c * q[l + 1] * 2 * q[l + 2]
    energy = n_xi * D * sum(8 * (2 * b[i] ** 2 * b[i + 1] + 2 * b[i] * b[i + 1] ** 2) for i in range(N))
    return n_xi * (qr + gemm + uz + svd + absorb) + energy

def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
