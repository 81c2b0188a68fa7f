"""Sweep harness with the reference's on-disk formats (benchmarker-mps.py:29-117).

The reference script is module-level code: a list of parameter dicts -> `function_to_run(**settings)` ->
one CSV row per run (`P,dt,max_bond_dim,datafile,output`, header written once), failures logged to
`errors.log` and recorded as `None`, abort after MAX_FAILURE_TOLERANCE failures, optional `.npy` eps(s)
curves of length P+1 (complex).  Same behaviour here as functions, with the GPU driver underneath.
Runs that share (N, N_xi, P, max_bond) but differ in the pattern file and/or in dt are batched into one
plan (one realisation each): that is how an ensemble, a set of realisations (data/plotter.py:175-195) or the
dt grid of benchmarker-mps.py:50-57 maps onto the GPU (`run_batched`, `run_sweep_batched`).
"""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np

from .dQA_mps import mydQA

MAX_FAILURE_TOLERANCE = 4  # benchmarker-mps.py:68


def function_to_run(P, dt, max_bond_dim, datafile, save_curves=None, **ignore):
    """benchmarker-mps.py:32-43: returns Re eps(1); optionally saves the complex eps(s) curve."""
    obj = mydQA(datafile, P=P, dt=dt, max_bond=max_bond_dim)
    obj.init_fourier()
    obj.run(skip_jit=4, print_info=False)
    if save_curves:
        np.save(save_curves, np.array([el[1] for el in obj.loss]))
    return float(np.real(obj.loss[-1][1]))


def run_batched(settings_list, save_dir=None):
    """Runs that share P and max_bond_dim (and the pattern shape) as one batched plan -> list of Re eps(1).
    `datafile` and `dt` may differ from run to run (one Fourier table per distinct dt)."""
    first = settings_list[0]
    for s in settings_list:
        if s["P"] != first["P"] or s["max_bond_dim"] != first["max_bond_dim"]:
            raise ValueError("run_batched: P and max_bond_dim must be common to the batch")
    xi = np.stack([np.load(s["datafile"]) if isinstance(s["datafile"], str) else np.asarray(s["datafile"]) for s in settings_list])
    dts = [float(s["dt"]) for s in settings_list]
    obj = mydQA(xi, P=first["P"], dt=dts[0] if len(set(dts)) == 1 else dts, max_bond=first["max_bond_dim"])
    obj.init_fourier()
    obj.run(print_info=False)
    if save_dir:
        for i, s in enumerate(settings_list):
            np.save(os.path.join(save_dir, "dt{}_bd{}_{}.npy".format(s["dt"], s["max_bond_dim"], i)), obj.eps[i])
    return [float(v) for v in np.real(obj.eps[:, -1])]


def run_sweep_batched(parameter_combinations, benchmark_file, error_log_file="errors.log", max_failures=MAX_FAILURE_TOLERANCE):
    """The sweep of benchmarker-mps.py:88-117 with the parameter sets grouped by (P, max_bond_dim, pattern shape):
    every group is one batched plan (dt and datafile vary inside it).  Rows are written in the original order;
    a failing group logs one error and records None for each of its rows."""
    write_header(benchmark_file, parameter_combinations[0].keys())
    groups = {}
    for i, s in enumerate(parameter_combinations):
        shape = np.load(s["datafile"]).shape if isinstance(s["datafile"], str) else np.asarray(s["datafile"]).shape
        groups.setdefault((s["P"], s["max_bond_dim"], shape), []).append(i)
    values, failures = [None] * len(parameter_combinations), 0
    for idx in groups.values():
        sets = [parameter_combinations[i] for i in idx]
        try:
            for i, v in zip(idx, run_batched(sets)):
                values[i] = v
        except Exception as e:  # noqa: BLE001 -- as the reference: log and go on
            failures += 1
            log_error(error_log_file, sets, e)
        if failures >= max_failures:
            break
    for s, v in zip(parameter_combinations, values):
        append_row(benchmark_file, s, v)
    if failures >= max_failures:
        raise Exception("multiple failures occurred in benchmarks, aborting...")
    return values


def write_header(benchmark_file, keys):
    """benchmarker-mps.py:74-82."""
    if not os.path.isfile(benchmark_file):
        with open(benchmark_file, "w") as f:
            f.write(",".join("{}".format(k) for k in keys))
            f.write(",output\n")
        return True
    return False


def append_row(benchmark_file, settings, value):
    """benchmarker-mps.py:110-114."""
    with open(benchmark_file, "a") as f:
        f.write(",".join("{}".format(v) for v in settings.values()))
        f.write("," + str(value) + "\n")


def log_error(error_log_file, settings, exc):
    """benchmarker-mps.py:101-108."""
    if error_log_file is None:
        return
    with open(error_log_file, "a") as f:
        f.write("\n[{}] params {} failed ----------\n".format(datetime.now().strftime("%d/%m/%Y %H:%M:%S"), str(settings)))
        f.write(str(exc))


def run_sweep(parameter_combinations, benchmark_file, error_log_file="errors.log", runner=function_to_run,
              max_failures=MAX_FAILURE_TOLERANCE):
    """The EXECUTE loop of benchmarker-mps.py:88-117.  Returns the list of logged values."""
    write_header(benchmark_file, parameter_combinations[0].keys())
    failures, out = 0, []
    for settings in parameter_combinations:
        try:
            value = runner(**settings)
        except Exception as e:  # noqa: BLE001 -- the reference catches everything and logs it
            value = None
            failures += 1
            log_error(error_log_file, settings, e)
        append_row(benchmark_file, settings, value)
        out.append(value)
        if failures >= max_failures:
            raise Exception("multiple failures occurred in benchmarks, aborting...")
    return out
