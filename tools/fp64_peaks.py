"""FP64 peaks of this GPU, three ways -> profiles/r02_fp64_peaks.json (when run with --write):
cuBLAS DGEMM 8192^3 through torch.matmul (best of 10, CUDA events; independent of this repo's code), and libdqa.so's
register-resident DFMA loop and mma.sync m8n8k4 f64 loop.  bench.py measures the same three live and divides by the
largest."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from perceptron_dqa_b200.dQA_mps import mydQA  # noqa: E402

xi = (2 * np.random.default_rng(0).integers(0, 2, size=(64, 17, 21)) - 1).astype(np.int8)
o = mydQA(xi, P=10, dt=1.0, max_bond=32)
o.init_fourier()
out = bench.fp64_peaks(o.plan, 0)
out["largest"] = max(out, key=out.get)
try:
    import torch

    out["gpu"] = torch.cuda.get_device_name(0)
except Exception:
    pass
print(json.dumps(out))
if "--write" in sys.argv:
    with open(os.path.join(ROOT, "gpurun_out", "r02_fp64_peaks.json"), "w") as f:
        json.dump(out, f, indent=1)
