"""CPU, world_size = 2, gloo: the N > 1 host logic.  Ladders are sharded by global id; because the random slots are
keyed by the global ladder id, two ranks running 2 ladders each must reproduce, bit for bit, one process running
all 4 — and the pooled statistics must equal the single-process ones.  (The chains themselves come from the CPU
oracle here; on GPUs the same sharding drives the engine, bench.py.)"""
import os
import socket

import numpy as np
import pytest

import models
from nimbleapt_b200.sharding import local_moments, pooled_mean_var, shard_ladders


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, niter, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch.distributed as dist
    from helpers import all_reduce_sum, build_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ladder0, nl = shard_ladders(total, world, rank)
    orc = build_oracle(models.c2_mixture(n=20, K=6, tmax=100.0), nLadders=nl, ladder0=ladder0)
    orc.run(niter, thin=2)
    s = np.stack([orc.samples(l) for l in range(nl)])
    pooled = all_reduce_sum(local_moments(s), dist)
    np.save(os.path.join(out, "s%d.npy" % rank), s)
    np.save(os.path.join(out, "p%d.npy" % rank), pooled)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_partition_is_a_partition():
    for total, world in ((4096, 8), (10, 4), (3, 8), (65536, 8)):
        ids = []
        for r in range(world):
            l0, n = shard_ladders(total, world, r)
            ids += list(range(l0, l0 + n))
        assert ids == list(range(total))
    with pytest.raises(ValueError):
        shard_ladders(4, 2, 2)


def test_two_ranks_reproduce_one_process(tmp_path):
    import torch.multiprocessing as mp
    from helpers import build_oracle
    total, niter, world = 4, 300, 2
    mp.spawn(_worker, args=(world, _free_port(), total, niter, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / ("s%d.npy" % r)) for r in range(world)])
    orc = build_oracle(models.c2_mixture(n=20, K=6, tmax=100.0), nLadders=total)
    orc.run(niter, thin=2)
    ref = np.stack([orc.samples(l) for l in range(total)])
    assert np.array_equal(got, ref)  # bit-identical: sharding does not change any ladder's chain
    p0, p1 = np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy")
    assert np.array_equal(p0, p1)
    np.testing.assert_allclose(p0, local_moments(ref), rtol=1e-13)
    mean, var = pooled_mean_var(p0, 2)
    np.testing.assert_allclose(mean, ref.reshape(-1, 2).mean(0), rtol=1e-12)
    np.testing.assert_allclose(var, ref.reshape(-1, 2).var(0, ddof=1), rtol=1e-10)
