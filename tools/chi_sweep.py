"""Throughput vs bond dimension at N=21, N_xi=17 (SURVEY config C3): realisation-steps/s of a saturated step.
usage: python tools/chi_sweep.py 10:888 16:888 ... (chi:batch)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for arg in sys.argv[1:]:
    chi, batch = arg.split(":")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_step.py"), "--chi", chi, "--batch", batch, "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True).stdout.strip().split("\n")[-1]
    d = json.loads(out)
    print(json.dumps({"chi": int(chi), "batch": int(batch), "s_per_step": round(d["wall_s_per_step"], 4),
                      "realisation_steps_per_s": round(int(batch) / d["wall_s_per_step"], 1),
                      "ms": {k: round(v, 1) for k, v in d["ms_per_step"].items()}, "bonds_mid": d["bonds"][len(d["bonds"]) // 2]}), flush=True)
