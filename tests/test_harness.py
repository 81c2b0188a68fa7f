"""CPU: the sweep harness keeps the reference's CSV / error-log formats (benchmarker-mps.py:74-117)."""
import os

import pytest

from perceptron_dqa_b200 import benchmarker as bm


def test_csv_rows_and_error_log(tmp_path):
    csv, log = str(tmp_path / "test15.csv"), str(tmp_path / "errors.log")
    combos = [{"P": 100, "dt": dt, "max_bond_dim": 10, "datafile": "data/patterns_12-15.npy"} for dt in (0.1, 0.2, 0.3)]

    def fake(P, dt, max_bond_dim, datafile, **_):
        if dt == 0.2:
            raise RuntimeError("boom")
        return dt * 2

    out = bm.run_sweep(combos, csv, log, runner=fake)
    assert out == [0.2, None, 0.6]
    lines = open(csv).read().splitlines()
    assert lines[0] == "P,dt,max_bond_dim,datafile,output"
    assert lines[1] == "100,0.1,10,data/patterns_12-15.npy,0.2"
    assert lines[2].endswith(",None")
    assert "params {'P': 100, 'dt': 0.2" in open(log).read() and "boom" in open(log).read()
    # header is not repeated on a second sweep into the same file
    bm.run_sweep(combos[:1], csv, log, runner=fake)
    assert open(csv).read().count("P,dt,max_bond_dim") == 1


def test_abort_after_too_many_failures(tmp_path):
    csv = str(tmp_path / "x.csv")

    def bad(**_):
        raise ValueError("no")

    combos = [{"P": 1, "dt": 0.1, "max_bond_dim": 2, "datafile": "f"}] * 6
    with pytest.raises(Exception, match="multiple failures"):
        bm.run_sweep(combos, csv, None, runner=bad)
    assert len(open(csv).read().splitlines()) == 1 + bm.MAX_FAILURE_TOLERANCE


@pytest.mark.gpu
def test_sweep_on_gpu(tmp_path):
    import numpy as np

    from helpers import golden

    g = golden("n7_chi4")
    f = str(tmp_path / "patterns_5-7.npy")
    np.save(f, g["xi"])
    curve = str(tmp_path / "curve.npy")
    v = bm.function_to_run(int(g["P"]), float(g["dt"]), int(g["chi"]), f, save_curves=curve)
    c = np.load(curve)
    assert c.shape == (int(g["P"]) + 1,) and c.dtype == np.complex128 and abs(v - c[-1].real) == 0
    assert abs(c[3] - g["eps"][3]) < 1e-10
    vals = bm.run_batched([{"P": int(g["P"]), "dt": float(g["dt"]), "max_bond_dim": int(g["chi"]), "datafile": f}] * 2)
    assert vals[0] == vals[1] == v


@pytest.mark.gpu
def test_dt_grid_sweep_batched_on_gpu(tmp_path):
    """benchmarker-mps.py:50-57 sweeps dt with one run per value; run_sweep_batched maps the grid onto the ensemble
    dimension (one plan per (P, max_bond_dim, pattern shape)) and writes the same CSV rows in the same order."""
    import numpy as np

    from helpers import golden

    g = golden("n7_chi4")
    f = str(tmp_path / "patterns_5-7.npy")
    np.save(f, g["xi"])
    P, chi = int(g["P"]), int(g["chi"])
    combos = [{"P": P, "dt": dt, "max_bond_dim": chi, "datafile": f} for dt in (0.3, float(g["dt"]), 1.1)]
    combos.append({"P": P, "dt": 0.3, "max_bond_dim": chi + 2, "datafile": f})  # a second group (another bond dimension)
    csv = str(tmp_path / "grid.csv")
    vals = bm.run_sweep_batched(combos, csv, str(tmp_path / "errors.log"))
    lines = open(csv).read().splitlines()
    assert lines[0] == "P,dt,max_bond_dim,datafile,output" and len(lines) == 5
    for s_, v, ln in zip(combos, vals, lines[1:]):
        one = bm.function_to_run(**s_)
        assert abs(v - one) < 1e-12 and ln.endswith(str(v))
