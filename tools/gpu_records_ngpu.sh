#!/bin/bash
# Round-2 multi-GPU records (run through `gpurun --gpus N -- bash tools/gpu_records_ngpu.sh N`); outputs under gpurun_out/.
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "weak rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --ensemble-total 1024 --chi 10 --P 100 --no-cpu-baseline --save gpurun_out/r02_c4_eps_mean_${N}gpu.npz > gpurun_out/r02_c4_${N}gpu.json 2> gpurun_out/r02_c4_${N}gpu.err; echo "c4 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r02_bench_${N}gpu.json", "gpurun_out/r02_c4_${N}gpu.json"):
    d = json.loads(open(f).read().strip().split("\n")[-1])
    print(f, d["n_gpus"], round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", d["e2e"].get("value"), d["e2e"].get("seconds"))
PY
