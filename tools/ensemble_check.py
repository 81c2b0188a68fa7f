"""torchrun --nproc-per-node 2 tools/ensemble_check.py : EnsembleDQA (NCCL all-reduce of the eps moments)
against the same ensemble evolved in one process."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from perceptron_dqa_b200.dQA_mps import mydQA  # noqa: E402
from perceptron_dqa_b200.ensemble import EnsembleDQA  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rng = np.random.default_rng(20230317)
xi = (2 * rng.integers(0, 2, size=(10, 5, 7)) - 1).astype(np.int8)
ens = EnsembleDQA(xi, P=12, dt=0.7, max_bond=4, device=local)
mean = ens.run()
if dist.get_rank() == 0:
    one = mydQA(xi, P=12, dt=0.7, max_bond=4, device=local)
    one.init_fourier()
    one.run(print_info=False)
    want = one.eps.real.mean(0)
    err = float(np.abs(mean - want).max())
    print(f"ensemble_check world={dist.get_world_size()} count={ens.count[0]:.0f} max|mean-ref|={err:.2e}", flush=True)
    assert err < 1e-13 and ens.count[0] == 10
dist.barrier()
dist.destroy_process_group()
