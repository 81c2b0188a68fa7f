"""Sweep block sizes through the library's env knobs.
usage: python tools/block_size_sweep.py theta_energy 10:32 16:64 ...   (DQA_THETA_THREADS = DQA_ENERGY_THREADS = threads)
       python tools/block_size_sweep.py lq 10:128 20:256 ...           (DQA_LQ_THREADS)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
what = sys.argv[1]
for arg in sys.argv[2:]:
    chi, t = arg.split(":")
    env = dict(os.environ)
    if what == "lq":
        env["DQA_LQ_THREADS"] = t
    else:
        env["DQA_THETA_THREADS"] = env["DQA_ENERGY_THREADS"] = t
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_step.py"), "--chi", chi, "--batch", os.environ.get("SWEEP_BATCH", "888"), "--steps", "2",
                          "--warmup", "1"], env=env, capture_output=True, text=True).stdout.strip().split("\n")[-1]
    d = json.loads(out)
    keys = ("lq",) if what == "lq" else ("theta", "energy")
    print(chi, t, {k: round(v, 1) for k, v in d["ms_per_step"].items() if k in keys}, round(d["wall_s_per_step"], 4), flush=True)
