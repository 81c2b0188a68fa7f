#!/bin/bash
# Round-2 final records on ONE B200 (run through gpurun from the repo root); outputs under gpurun_out/.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02.log
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
python tools/chi_sweep.py 10:888 16:888 20:888 24:888 32:888 40:296 48:296 56:296 64:296 > gpurun_out/chi_sweep_r02.jsonl 2> gpurun_out/chi_sweep_r02.err
python bench.py --ensemble-total 1024 --chi 10 --P 100 --no-cpu-baseline --save gpurun_out/r02_c4_eps_mean_1gpu.npz > gpurun_out/r02_c4_1gpu.json 2> gpurun_out/r02_c4_1gpu.err; echo "c4 rc=$?"
# launch list of one bench step (3 warm-up steps of 3*1431+1 launches each are skipped, plus the set-up kernels)
ncu --metrics gpu__time_duration.sum --clock-control none -s 12888 -c 4294 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_launches_r02.log 2>&1; echo "ncu rc=$?"
tail -c 1500 gpurun_out/bench_r02.json; echo; cat gpurun_out/chi_sweep_r02.jsonl; tail -c 600 gpurun_out/r02_c4_1gpu.json; echo; head -5 gpurun_out/launches_r02.csv | cut -c1-200
