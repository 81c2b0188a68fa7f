"""CPU, world_size 2 over gloo: the host-side logic of the ensemble path (sharding, the
moments layout of k_ensemble_moments, the single all-reduce, masking of failed instances)."""
import os
import socket

import numpy as np
import pytest

from perceptron_dqa_b200.ensemble import finalize_moments, moments_from_eps, shard_bounds


def test_shard_bounds_partition():
    for R in (1, 7, 16, 1024):
        for W in (1, 2, 3, 4, 8):
            covered = []
            for r in range(W):
                lo, hi = shard_bounds(R, W, r)
                covered += list(range(lo, hi))
            assert covered == list(range(R))
    assert shard_bounds(1024, 8, 3) == (384, 512)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, R, P, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    eps_all = rng.normal(size=(R, P + 1)) + 1e-16j * rng.normal(size=(R, P + 1))
    status_all = np.zeros(R, dtype=np.int32)
    status_all[3] = 1  # one diverged realisation must not poison the average
    lo, hi = shard_bounds(R, world, rank)
    mom = torch.from_numpy(moments_from_eps(eps_all[lo:hi], status_all[lo:hi]))
    dist.all_reduce(mom)
    mean, sem, cnt = finalize_moments(mom.numpy())
    ok = status_all == 0
    want = eps_all.real[ok].mean(0)
    out.put((rank, float(np.abs(mean - want).max()), float(cnt[0]), float(sem[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_of_moments_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    R, P = 11, 6
    procs = [ctx.Process(target=_worker, args=(r, 2, port, R, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _rank, err, cnt, sem in res:
        assert err < 1e-14 and cnt == R - 1 and sem > 0


def test_moments_layout_matches_kernel_contract():
    eps = np.array([[1 + 0j, 2], [3, 4], [5, 6]], dtype=np.complex128)
    st = np.array([0, 4, 0])
    m = moments_from_eps(eps, st)
    assert m.tolist() == [6.0, 8.0, 26.0, 40.0, 2.0, 2.0]
