"""FP64 peaks of this GPU as libdqa.so measures them: register-resident DFMA loop and mma.sync m8n8k4 f64 loop."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceptron_dqa_b200.dQA_mps import mydQA  # noqa: E402

xi = (2 * np.random.default_rng(0).integers(0, 2, size=(64, 17, 21)) - 1).astype(np.int8)
o = mydQA(xi, P=10, dt=1.0, max_bond=32)
o.init_fourier()
print({"dfma_tflops": o.plan.fp64_peak_tflops(), "dmma_tflops": o.plan.dmma_peak_tflops()})
