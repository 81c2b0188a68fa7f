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
    ap.add_argument("--ensemble-total", type=int, default=0,
                    help="C4 mode (BASELINE.json config 4): a FIXED ensemble of this many distinct realisations split over the "
                         "ranks (strong scaling), full P-step run from |+>, all-reduced eps(s) inside the timed region")
    ap.add_argument("--save", default="", help="C4 mode: write mean/sem of eps(s) to this .npz")
    ap.add_argument("--no-roofline", action="store_true")
    return ap.parse_args()


def ensemble_patterns(a, rank, batch):
    """C4 law: iid uniform +-1, default_rng(20230317); realisation r of the job is
    row r of that table; global realisation 0 is the reference's patterns_17-21.1.npy."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rows = max(1024, batch * world)  # the first 1024 rows are the C4 table; larger jobs continue the same stream (no repeats)
    rng = np.random.default_rng(20230317)
    table = (2 * rng.integers(0, 2, size=(rows, 17, 21)) - 1).astype(np.int8)
    if (a.n, a.nxi) != (21, 17):
        table = (2 * rng.integers(0, 2, size=(rows, a.nxi, a.n)) - 1).astype(np.int8)
    idx = rank * batch + np.arange(batch)
    xi = table[idx].copy()
    if rank == 0 and (a.n, a.nxi) == (21, 17):
        g = np.load(os.path.join(ROOT, "tests", "golden", "c2_n21_chi10.npz"))
        xi[0] = g["xi"]
    return xi


def workload_name(a, batch, world):
    return (f"ensemble of {batch * world} pattern realisations ({batch}/GPU; #0 = patterns_17-21.1.npy, rest iid +-1 "
            f"rng 20230317), N={a.n}, N_xi={a.nxi}, chi={a.chi}, P={a.P}, dt={a.dt}, complex128, steps at p~P/2")


# ----------------------------------------------------------------------------------------
# CPU legs.  The baseline is the reference's OWN implementation: dQA_mps.py / lib/tenn.py / lib/dQA_utils.py,
# unmodified, imported through the numpy-backed jax shim (oracle/ref_runner.py) in complex128 -- from /root/reference
# where it exists, else from the verbatim copies oracle/make_ref.py staged in the git-ignored oracle/_ref/ (they travel
# to the GPU box).  Only if neither is present does the numpy port (oracle/dqa_oracle.py) stand in (kind "port").
# One process per host core, one realisation per process, 1 BLAS thread each (OpenBLAS threading slows these sizes).
# ----------------------------------------------------------------------------------------
def _evolved_state(n, nxi, chi):
    """A real evolved state with saturated bonds: the reference's own MPS after two steps at N=21, chi=32
    (tests/golden/c3_n21_chi32.npz, written by the unmodified reference).  None for other sizes."""
    if (n, nxi, chi) != (21, 17, 32):
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import golden, unpack_mps

    g = golden("c3_n21_chi32")
    return unpack_mps(g["snap1_post_bonds"], g["snap1_post_flat"])


def _cpu_worker(args):
    """Whole dQA steps of one realisation on one core.  Returns (kind, [seconds per step], untimed warm-up steps)."""
    (n, nxi, chi, P, dt, xi, nsteps, want) = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dqa_oracle as O
    import ref_runner

    state = _evolved_state(n, nxi, chi)
    warm = 0 if state is not None else 1
    times = []
    if want == "reference" and ref_runner.available():
        mod = ref_runner.load(sgn0fix=False)  # verbatim, stock code path
        obj = mod.mydQA(np.asarray(xi).astype(np.float64), P=P, dt=dt, max_bond=chi)
        obj.init_fourier()
        obj.psi = state if state is not None else [np.array([[[2 ** -0.5], [2 ** -0.5]]], dtype=np.complex128)] * n
        obj.pp = P // 2
        obj.loss, obj.bd_monitor = [], []  # the two traces run() would have created (dQA_mps.py:141-147)
        with ref_runner.quiet():
            for k in range(warm + nsteps):
                t0 = time.perf_counter()
                obj.single_step()  # dQA_mps.py:84-119
                if k >= warm:
                    times.append(time.perf_counter() - t0)
        return "reference", times, warm
    o = O.OracleDQA(xi, P, dt, chi)
    o.init_fourier()
    o.psi = state if state is not None else O.plus_state(n)
    o.pp = P // 2
    o.loss, o.bd_monitor = [], []
    for k in range(warm + nsteps):
        t0 = time.perf_counter()
        o.single_step()
        if k >= warm:
            times.append(time.perf_counter() - t0)
    return "port", times, warm


def cpu_steps(a, cores, nsteps, want="reference"):
    """`nsteps` whole steps on each of `cores` processes at once.  Returns (kind, mean s per step per core, wall, warm)."""
    import multiprocessing as mp

    xi = ensemble_patterns(a, 0, max(cores, 1))
    jobs = [(a.n, a.nxi, a.chi, a.P, a.dt, xi[i], nsteps, want) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    per_step = [t for _k, ts, _w in res for t in ts]
    return res[0][0], float(np.mean(per_step)), wall, res[0][2]


def cpu_same_algorithm(a):
    """One step of the SECTOR algorithm (what the kernels run) in numpy on one core (oracle/sector_model.py): splits the
    GPU-vs-reference ratio into an algorithmic and a hardware factor.  None if no evolved state is at hand."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dqa_oracle as O
    import sector_model as S

    psi = _evolved_state(a.n, a.nxi, a.chi)
    if psi is None:
        return None
    xi = ensemble_patterns(a, 0, 1)[0]
    hbit, _, _ = O.hamming_tables(xi)
    uk, fx = O.fourier_tables(a.n, a.P, a.dt)
    p = a.P // 2
    u = S.uz_diagonal(uk[:, p], a.n)
    t0 = time.perf_counter()
    for mu in range(a.nxi):
        psi = S.apply_pattern_sector(psi, u, np.stack([hbit[mu], 1 - hbit[mu]], axis=1), a.chi)
    psi = S.mixer(psi, (1 - (p + 1) / a.P) * a.dt)
    O.compute_loss_fast(psi, xi, fx)
    return time.perf_counter() - t0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline_block(a, nsteps):
    cores = min(host_cores(), 64)
    kind, t_step, wall, warm = cpu_steps(a, cores, nsteps)
    what = ("mydQA.single_step() of the unmodified reference (dQA_mps.py:84-119 through the numpy jax shim, complex128)"
            if kind == "reference" else "single_step() of the numpy port oracle/dqa_oracle.py (reference files not staged on this box)")
    start = ("the reference's own evolved MPS (after 2 steps, bonds saturated at chi)" if warm == 0
             else f"|+>^N and {warm} untimed step")
    blk = {"value": cores / t_step, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"{nsteps} whole dQA step(s) per core at p=P/2: {what}; start state {start}; {cores} processes x 1 BLAS "
                     f"thread, one realisation each; {wall:.0f} s wall",
           "s_per_step_per_core": t_step, "whole_steps": nsteps}
    t_alg = cpu_same_algorithm(a)
    if t_alg:
        blk["same_algorithm"] = {"s_per_step_per_core": t_alg, "value": cores / t_alg, "cores_assumed": cores,
                                 "what": "one step of oracle/sector_model.py (numpy statement of the sector algorithm the kernels "
                                         "execute) on one core: reference/this = algorithmic factor, this/GPU = hardware factor"}
    return blk


def run_reference(a):
    """--impl reference: whole steps of the reference's own single_step() on all host cores, one realisation per core.
    A step of the reference at N=21, chi=32 takes ~1.5-2 min per core, so the run is bounded to the number of whole steps
    that fits --cpu-seconds (at least 2 when --steps >= 2); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k = max(1, min(a.steps, 2 if a.cpu_seconds < 400 else int(a.cpu_seconds // 100)))
    blk = cpu_baseline_block(a, k)
    value, cores = blk["value"], blk["cores"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": k,
        "warmup": 0, "ms_per_step": 1e3 * blk["s_per_step_per_core"] / cores, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": {"workload": workload_name(a, a.batch, a.gpus),
                   "note": f"CPU arm: {cores} host cores, one realisation per core, {k} whole steps each, no extrapolation "
                           "(numpy has no JIT or cache to warm; start state: see cpu_baseline.sample)"},
        "cpu_baseline": blk,
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
    ms_steps = [x.elapsed_time(y) for x, y in ev]
    ms = sum(ms_steps)
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
    mom_host = torch.empty(3 * (a.P + 1), dtype=torch.float64).pin_memory()
    for _ in range(k_e2e):
        pl.set_patterns(xi)                      # H2D: the step's patterns
        pl.upload_state(host.data_ptr())         # H2D: the step's input state (pinned)
        pl.step(1)
        pl.download_state(host.data_ptr())       # D2H: the new state
        eps_host[:] = pl.eps_trace()             # D2H: eps(s) trace
        pl.ensemble_moments_into(mom.data_ptr())  # the path's one collective: sum eps, sum eps^2, #healthy over the job
        if dist is not None:
            dist.all_reduce(mom)
        mom_host.copy_(mom, non_blocking=False)  # D2H: the ensemble result of this step
    e1.record()
    barrier()
    t_e2e = time.perf_counter() - t0
    ms_e2e = e0.elapsed_time(e1)
    t2 = torch.tensor([max(ms_e2e * 1e-3, t_e2e)], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = B * world * k_e2e / float(t2.item())
    h2d = nbytes + xi.nbytes
    d2h = nbytes + eps_host.nbytes + 4 * B + mom_host.numel() * 8

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
        "ms_per_step": ms_max / a.steps, "ms_of_each_step_rank0": [round(v, 2) for v in ms_steps],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex128", "data": "synthetic",
        "config": {"workload": workload_name(a, B, world), "batch_per_gpu": B,
                   "l2": "flushed (256 MiB write) between timed steps; per-step working set "
                         f"{pl.workspace_bytes / 2**20:.0f} MiB/GPU", "parallelism": f"ensemble x{world}"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": k_e2e},
        "gpu_launches": int(launches), "clocks": clocks,
        "single_realisation": None,
        "ensemble": {"eps_mean_at_last_step": eps_mean, "healthy": n_ok, "distinct_realisations": B * world,
                     "allreduce": ("nccl, inside the e2e region every step" if dist is not None else "none (1 rank); moments inside the e2e region")},
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
        peaks64 = fp64_peaks(pl, local)
        peak_name = max(peaks64, key=peaks64.get)
        peak = peaks64[peak_name]
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
            "peak_source": f"largest of three FP64 peaks measured on this GPU in this run: {peak_name} (MEASURED_PEAKS.json has no "
                           "FP64 entry): cuBLAS DGEMM 8192^3 via torch.matmul (best of 10, CUDA events), libdqa's register-resident "
                           "DFMA loop, libdqa's mma.sync m8n8k4 f64 loop",
            "fp64_peaks_tflops": peaks64,
            "flops_per_launch": fams[dom]["flops_per_step"] / max(1, fams[dom]["launches"]),
            "ms_per_launch": fams[dom]["ms_per_step"] / max(1, fams[dom]["launches"]),
            "share_of_step": fams[dom]["share_of_step"], "families": fams,
            "whole_step": {"flops": float(sum(flops.values())), "achieved_tflops": float(sum(flops.values())) / (sum(fam_ms.values()) * 1e-3) / 1e12,
                           "frac": float(sum(flops.values())) / (sum(fam_ms.values()) * 1e-3) / 1e12 / peak if peak else None,
                           # the family times come from a profiled step (one kernel family at a time, whole batch per launch);
                           # the timed steps run the batch as sub-batches on parallel branches of the step graph (DQA_SPLIT)
                           "timed_step": {"ms": ms_max / a.steps, "achieved_tflops": float(sum(flops.values())) / (ms_max / a.steps * 1e-3) / 1e12,
                                          "frac": float(sum(flops.values())) / (ms_max / a.steps * 1e-3) / 1e12 / peak if peak else None}},
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
            prof = next(pth for pth in (os.path.join(ROOT, "profiles", f"{r}_{line['roofline']['kernel']}_ncu_summary.csv") for r in ("r02", "r01"))
                        if os.path.isfile(pth))
            with open(prof) as f:
                for ln in f:
                    parts = ln.strip().split(",")
                    if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        tr += float(parts[2]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[parts[1]]
            if tr > 0:
                line["roofline"]["traffic"] = tr
                line["roofline"]["traffic_source"] = (f"profiles/{os.path.basename(prof)}: dram read+write of one "
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
        line["cpu_baseline"] = cpu_baseline_block(a, 1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)



def run_c4(a):
    """BASELINE.json config 4 as stated: a FIXED ensemble of --ensemble-total distinct synthetic realisations (C4 law) split
    over the ranks, the whole P-step run from |+>^N, and the ensemble mean / standard error of eps(s) all-reduced over
    NCCL -- everything inside the timed region.  Strong scaling: total work is fixed as ranks are added."""
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
    from perceptron_dqa_b200.ensemble import finalize_moments

    total = a.ensemble_total
    if total % world:
        raise SystemExit("--ensemble-total must be a multiple of the number of ranks")
    B = total // world
    xi = ensemble_patterns(a, rank, B)  # rows rank*B .. rank*B+B-1 of the C4 table: all distinct
    dev = f"cuda:{local}"
    obj = mydQA(xi if B > 1 else xi[0], P=a.P, dt=a.dt, max_bond=a.chi, device=local)
    obj.init_fourier()
    pl = obj.plan
    mom = torch.zeros(3 * (a.P + 1), dtype=torch.float64, device=dev)
    res_host = torch.empty(3 * (a.P + 1), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def whole_run():
        pl.reset_plus()
        pl.step(a.P)
        pl.ensemble_moments_into(mom.data_ptr())
        if dist is not None:
            dist.all_reduce(mom)

    W = max(3, a.warmup)
    pl.reset_plus()
    pl.step(min(W, a.P))
    if dist is not None:
        dist.all_reduce(mom)  # NCCL communicator set-up outside the timed region
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # ---- device-timed: the whole ensemble run + the collective -----------------------------------------------
    flush.fill_(1)
    l0 = pl.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    whole_run()
    e1.record()
    barrier()
    launches = pl.launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.summary()
    # ---- end to end: host patterns and tables in, ensemble curve out ----------------------------------------------
    flush.fill_(2)
    barrier()
    t0 = time.perf_counter()
    pl.set_patterns(xi)
    pl.set_fourier(obj.Uk_FT, obj.fxft)
    whole_run()
    res_host.copy_(mom)
    mean, sem, n_ok = finalize_moments(res_host.numpy())
    torch.cuda.synchronize()
    t2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    s_e2e = float(t2.item())
    status = int(np.count_nonzero(pl.status()))
    if rank == 0:
        line = {
            "metric": METRIC.replace("chi=32", f"chi={a.chi}") + " (C4 ensemble run)", "value": total * a.P / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": a.P, "warmup": W, "ms_per_step": ms / a.P, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
            "config": {"workload": f"C4: fixed ensemble of {total} distinct realisations (default_rng(20230317), iid +-1), N={a.n}, "
                                   f"N_xi={a.nxi}, chi={a.chi}, P={a.P}, dt={a.dt}, full run from |+>^N, {B} realisations per GPU; timed "
                                   "region = reset + P steps + ensemble moments + NCCL all-reduce of 3(P+1) doubles",
                       "l2": f"flushed before the run; working set {pl.workspace_bytes / 2**20:.0f} MiB/GPU", "parallelism": f"ensemble x{world}"},
            "e2e": {"value": total * a.P / s_e2e, "unit": UNIT, "h2d_bytes_per_step": int((xi.nbytes + obj.Uk_FT.nbytes + obj.fxft.nbytes) / a.P),
                    "d2h_bytes_per_step": int(res_host.numel() * 8 / a.P), "seconds": s_e2e,
                    "what": "host patterns + Fourier tables in, whole run, all-reduce, mean/sem of eps(s) out (wall clock, max over ranks)"},
            "gpu_launches": int(launches), "clocks": clocks, "run_seconds": ms * 1e-3,
            "ensemble": {"healthy_at_end": float(n_ok[-1]), "eps_mean_first": float(mean[0]), "eps_mean_last": float(mean[-1]),
                         "eps_sem_last": float(sem[-1]), "status_nonzero_rank0": status,
                         "allreduce": "nccl (inside the timed region)" if dist is not None else "none (1 rank)"},
        }
        if a.save:
            np.savez_compressed(a.save, s=np.arange(a.P + 1) / a.P, eps_mean=mean, eps_sem=sem, healthy=n_ok, total=total, P=a.P,
                                dt=a.dt, chi=a.chi, n_gpus=world)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def fp64_peaks(pl, local):
    """FP64 roofline denominators measured live: a library DGEMM (independent of this repo's code) and libdqa's two probes."""
    import torch

    out = {"dfma_probe": pl.fp64_peak_tflops(), "dmma_probe": pl.dmma_peak_tflops()}
    try:
        n = 8192
        x = torch.randn(n, n, dtype=torch.float64, device=f"cuda:{local}")
        y = torch.randn(n, n, dtype=torch.float64, device=f"cuda:{local}")
        torch.matmul(x, y)
        best = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(x, y)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out["cublas_dgemm_8192"] = best
        del x, y
    except Exception as exc:  # pragma: no cover
        out["cublas_dgemm_8192"] = 0.0
        out["cublas_error"] = str(exc)[:100]
    return {k: v for k, v in out.items() if isinstance(v, float)}


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
    elif a.ensemble_total:
        run_c4(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
