"""-m gpu, needs 2 GPUs (skipped on a single-GPU box): the N > 1 path on real devices.  Two ranks share 5 ladders,
3 + 2 (sharding.shard_ladders: the counts differ on purpose; global ladder ids key the Philox stream); the traces are
gathered with NCCL inside the library -- while the run is still sampling (apt_run_args.gather, to either root), and
afterwards (apt_comm_gather_samples to either root, apt_comm_allgather_samples to everyone) -- and the pooled moments
are summed with apt_comm_allreduce_sum.  All of it must equal, bit for bit, one process running the 6 ladders on one GPU
(ladders never exchange anything while sampling).  The rank arithmetic itself is covered on CPU with
gloo in tests/test_sharding_gloo.py."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import models

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spec():
    return models.c2_mixture(n=20, K=8, tmax=200.0)


def _worker(rank, world, port, total, niter, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch
    import torch.distributed as dist
    from helpers import build_engine
    from nimbleapt_b200 import _lib
    from nimbleapt_b200.sharding import local_moments, shard_ladders
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # out-of-band channel for the ncclUniqueId only
    L = _lib.lib()
    ladder0, nl = shard_ladders(total, world, rank)
    eng = build_engine(_spec(), nLadders=nl, ladder0=ladder0, device=rank)
    obj = [eng.commUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    eng.commInit(world, rank, obj[0])
    rows, cols = niter // 2, len(eng.monitorNames)
    nall = total * rows * cols
    # (1) the gather inside the run: pieces travel to the root while the next piece is being sampled
    for root_rank in (0, 1):
        dest = np.full(nall, np.nan) if rank == root_rank else None
        eng.run(niter, thin=2, samplesOut=dest, gatherRoot=root_rank)
        if rank == root_rank:
            np.save(os.path.join(out, "run_root%d.npy" % root_rank), dest)
    # a root whose buffer is too small must fail without leaving the other rank waiting
    small = np.full(7, np.nan) if rank == 0 else None
    try:
        eng.run(niter, thin=2, samplesOut=small, gatherRoot=0)
        failed = False
    except Exception as e:
        failed = "buffer too small" in str(e)
    assert failed == (rank == 0)
    # (2) whole-trace gathers after a run
    eng.run(niter, thin=2)
    local = eng.mvSamples.matrix(-1)
    dp = C.POINTER(C.c_double)
    got = C.c_int64(-1)
    gathered = np.full(nall, np.nan)
    if rank == 0:
        _lib.check(L.apt_comm_gather_samples(eng._h, 0, 0, gathered.ctypes.data_as(dp), gathered.size, C.byref(got)))
        assert got.value == nall
    else:
        _lib.check(L.apt_comm_gather_samples(eng._h, 0, 0, None, 0, C.byref(got)))
        assert got.value == 0
    allg = np.full(nall, np.nan)
    _lib.check(L.apt_comm_allgather_samples(eng._h, 0, allg.ctypes.data_as(dp), allg.size, C.byref(got)))
    assert got.value == nall
    mom = eng.commAllReduceSum(local_moments(local))
    if rank == 0:  # a second gather, to another root, after the first one (buffers are reused)
        _lib.check(L.apt_comm_gather_samples(eng._h, 0, 1, None, 0, C.byref(got)))
    else:
        g1 = np.full(nall, np.nan)
        _lib.check(L.apt_comm_gather_samples(eng._h, 0, 1, g1.ctypes.data_as(dp), g1.size, C.byref(got)))
        np.save(os.path.join(out, "g_root1.npy"), g1)
    np.save(os.path.join(out, "local%d.npy" % rank), local)
    np.save(os.path.join(out, "allg%d.npy" % rank), allg)
    np.save(os.path.join(out, "mom%d.npy" % rank), mom)
    if rank == 0:
        np.save(os.path.join(out, "g_root0.npy"), gathered)
    eng.commFinalize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
def test_two_gpus_reproduce_one_and_nccl_gathers(tmp_path):
    import torch.multiprocessing as mp
    from helpers import build_engine
    from nimbleapt_b200.sharding import local_moments
    world, total, niter = 2, 5, 400
    mp.spawn(_worker, args=(world, _free_port(), total, niter, str(tmp_path)), nprocs=world, join=True)
    # one process, all 5 ladders, the same four runs as the workers (every run continues the random slots of the one before
    # and starts from the state the one before left in the model, APT:548; the workers' third run samples too before the
    # root reports its short buffer)
    eng = build_engine(_spec(), nLadders=total)
    refs = []
    for _ in range(4):
        eng.run(niter, thin=2)
        refs.append(eng.mvSamples.matrix(-1).copy())
    ref = refs[3]
    loc = [np.load(tmp_path / ("local%d.npy" % r)) for r in range(world)]
    assert [a.shape[0] for a in loc] == [3, 2]
    assert np.array_equal(np.concatenate(loc), ref)  # sharding does not change any ladder's chain
    for name, want in (("run_root0", refs[0]), ("run_root1", refs[1]), ("g_root0", ref), ("g_root1", ref), ("allg0", ref), ("allg1", ref)):
        assert np.array_equal(np.load(tmp_path / (name + ".npy")).reshape(want.shape), want), name
    m0, m1 = np.load(tmp_path / "mom0.npy"), np.load(tmp_path / "mom1.npy")
    assert np.array_equal(m0, m1)
    np.testing.assert_allclose(m0, local_moments(ref), rtol=1e-13)


def test_gather_within_a_run_on_one_rank(tmp_path):
    """The same code path with a communicator of one rank (runs on the single-GPU box): rows packed piece by piece on the
    device, handed to the front library, copied to the caller's buffer on the second stream."""
    from helpers import build_engine
    eng = build_engine(_spec(), nLadders=4)
    eng.commInit(1, 0, eng.commUniqueId())
    niter = 403  # pieces of unequal size, the last one short
    dest = np.full(4 * (niter // 3) * 2, np.nan)
    eng.run(niter, thin=3, samplesOut=dest, gatherRoot=0)
    assert np.array_equal(dest.reshape(4, niter // 3, 2), eng.mvSamples.matrix(-1))
    assert eng.getTimes()["run_wall_ms"] > 0
    np.testing.assert_array_equal(eng.commAllReduceSum(np.arange(5.0)), np.arange(5.0))
    eng.commFinalize()
