"""-m gpu, needs 2 GPUs (skipped on a single-GPU box): the N > 1 path on real devices.  Two ranks own 3 ladders each
(global ladder ids key the Philox stream); after sampling, the traces are gathered with NCCL inside the library
(apt_comm_gather_samples to rank 0, apt_comm_allgather_samples to everyone) and the pooled moments are summed with
apt_comm_allreduce_sum.  All of it must equal, bit for bit, one process running the 6 ladders on one GPU
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


def _worker(rank, world, port, per_rank, niter, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch
    import torch.distributed as dist
    from helpers import build_engine
    from nimbleapt_b200 import _lib
    from nimbleapt_b200.sharding import local_moments
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # out-of-band channel for the ncclUniqueId only
    L = _lib.lib()
    eng = build_engine(_spec(), nLadders=per_rank, ladder0=rank * per_rank, device=rank)
    uid = (C.c_ubyte * 128)()
    if rank == 0:
        _lib.check(L.apt_comm_unique_id(uid))
    obj = [bytes(uid)]
    dist.broadcast_object_list(obj, src=0)
    uid = (C.c_ubyte * 128)(*obj[0])
    _lib.check(L.apt_comm_init(eng._h, world, rank, uid))
    eng.run(niter, thin=2)
    local = eng.mvSamples.matrix(-1)
    n = local.size
    dp = C.POINTER(C.c_double)
    got = C.c_int64(-1)
    gathered = np.full(world * n, np.nan)
    if rank == 0:
        _lib.check(L.apt_comm_gather_samples(eng._h, 0, 0, gathered.ctypes.data_as(dp), gathered.size, C.byref(got)))
        assert got.value == world * n
    else:
        _lib.check(L.apt_comm_gather_samples(eng._h, 0, 0, None, 0, C.byref(got)))
        assert got.value == 0
    allg = np.full(world * n, np.nan)
    _lib.check(L.apt_comm_allgather_samples(eng._h, 0, allg.ctypes.data_as(dp), allg.size, C.byref(got)))
    assert got.value == world * n
    mom = np.ascontiguousarray(local_moments(local))
    _lib.check(L.apt_comm_allreduce_sum(eng._h, mom.ctypes.data_as(dp), mom.size))
    if rank == 0:  # a second gather, to another root, after the first one (buffers are reused)
        _lib.check(L.apt_comm_gather_samples(eng._h, 0, 1, None, 0, C.byref(got)))
    else:
        g1 = np.full(world * n, np.nan)
        _lib.check(L.apt_comm_gather_samples(eng._h, 0, 1, g1.ctypes.data_as(dp), g1.size, C.byref(got)))
        np.save(os.path.join(out, "g_root1.npy"), g1)
    np.save(os.path.join(out, "local%d.npy" % rank), local)
    np.save(os.path.join(out, "allg%d.npy" % rank), allg)
    np.save(os.path.join(out, "mom%d.npy" % rank), mom)
    if rank == 0:
        np.save(os.path.join(out, "g_root0.npy"), gathered)
    _lib.check(L.apt_comm_finalize(eng._h))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
def test_two_gpus_reproduce_one_and_nccl_gathers(tmp_path):
    import torch.multiprocessing as mp
    from helpers import build_engine
    from nimbleapt_b200.sharding import local_moments
    world, per_rank, niter = 2, 3, 400
    mp.spawn(_worker, args=(world, _free_port(), per_rank, niter, str(tmp_path)), nprocs=world, join=True)
    eng = build_engine(_spec(), nLadders=world * per_rank)
    eng.run(niter, thin=2)
    ref = eng.mvSamples.matrix(-1)
    loc = [np.load(tmp_path / ("local%d.npy" % r)) for r in range(world)]
    assert np.array_equal(np.concatenate(loc), ref)  # sharding does not change any ladder's chain
    for name in ("g_root0", "g_root1", "allg0", "allg1"):
        assert np.array_equal(np.load(tmp_path / (name + ".npy")).reshape(ref.shape), ref), name
    m0, m1 = np.load(tmp_path / "mom0.npy"), np.load(tmp_path / "mom1.npy")
    assert np.array_equal(m0, m1)
    np.testing.assert_allclose(m0, local_moments(ref), rtol=1e-13)
