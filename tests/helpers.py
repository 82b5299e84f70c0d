"""Shared builders: the same workload spec (tests/models.py) drives the CUDA engine (through the C ABI) and the
CPU oracle (test infrastructure)."""
import numpy as np

import nimbleapt_b200 as nb
from nimbleapt_b200.workloads import build_engine  # noqa: F401  (re-exported for the tests)
from oracle.oracle import OracleAPT




def build_oracle(spec, nLadders=1, seed=11, ladder0=0):
    return OracleAPT(spec["code"], spec["constants"], spec["data"], spec["inits"], spec["samplers"], spec["monitors"],
                     spec.get("monitors2", ()), spec["Temps"], spec["ULT"], True, nLadders, seed, ladder0, functions=spec.get("functions"))


def make_tables(niter, L, K, NU, DMAX, seed=1):
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((niter, L, K, NU, DMAX))
    U = rng.random((niter, L, K, NU))
    S = rng.random((niter, L, K, 2))
    return Z, U, S


def run_pair(spec, niter, L=1, inject=True, seed=11, thin=1, record=True, build_kw=None, **run_kw):
    """Run engine and oracle on identical inputs; returns (engine, oracle, records)."""
    eng = build_engine(spec, nLadders=L, seed=seed, record=record, **(build_kw or {}))
    orc = build_oracle(spec, nLadders=L, seed=seed)
    K, NU, DMAX = eng.nTemps, eng.info(nb.api.INFO_N_UNITS), eng.info(nb.api.INFO_DMAX)
    recs = []
    inj = None
    if inject:
        Z, U, S = make_tables(niter, L, K, NU, DMAX, seed=seed + 100)
        inj = dict(Z=Z, U=U, S=S)
    for l in range(L):
        if inject:
            recs.append(orc.set_tables(l, Z[:, l].copy(), U[:, l].copy(), S[:, l].copy(), record=True, niter=niter))
        else:
            recs.append(orc.set_tables(l, record=True, niter=niter))
    eng.run(niter, thin=thin, _inject=inj, **run_kw)
    orc.run(niter, thin=thin, **run_kw)
    return eng, orc, recs


def assert_run_parity(eng, orc, recs, spec, name, niter, L, temps_rtol=1e-9):
    """What "parity with the oracle" means for a finished pair of runs (north-star check 2): accept / reject decisions
    of every sampler of every rung and iteration bit-identical, swap decisions bit-identical, accepted swap partners
    identical, and all state the run leaves behind (rung states, ladder, traces, swap counts, sampler scales) equal to
    rounding."""
    for l in range(L):
        dec_o, swap_o = recs[l]
        dec_e = eng.decisions(l, niter)
        swap_e = eng.swapRecord(l, niter)
        assert dec_e.shape == dec_o.shape
        nbad = int((dec_e != dec_o).sum())
        assert nbad == 0, "%s ladder %d: %d accept decisions differ (first at %s)" % (name, l, nbad, np.argwhere(dec_e != dec_o)[:3])
        # swap phase: accept/reject flags identical everywhere, partners identical wherever a swap was accepted.
        # The partner of a REJECTED proposal may differ in a vanishing fraction of cases: when every
        # exp(-|lp_i - lp_j|) of a row lies in the denormal range, CUDA's and glibc's exp() round differently
        # and the normalised row (APT:560-569) is not reproducible across math libraries.
        assert np.array_equal(swap_e >> 16, swap_o >> 16), "%s ladder %d: swap accept decisions differ" % (name, l)
        tgt_bad = (swap_e & 0xffff) != (swap_o & 0xffff)
        assert not np.any(tgt_bad & ((swap_o >> 16) == 1)), "%s ladder %d: accepted swap partners differ" % (name, l)
        assert tgt_bad.mean() <= 1e-3, "%s ladder %d: %d proposed partners differ" % (name, l, tgt_bad.sum())
        np.testing.assert_allclose(eng.rungStates(l), np.stack([orc.rung_state(l, k) for k in range(eng.nTemps)]), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(eng.TempsLadder(l), orc.Temps(l), rtol=temps_rtol)
        np.testing.assert_allclose(eng.mvSamples.matrix(l), orc.samples(l), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(eng.mvSamplesTmax.matrix(l), orc.samplesTmax(l), rtol=1e-8, atol=1e-10)
        if spec.get("monitors2"):  # logProb_* monitors of the demo model: per-node log-probabilities of the saved state
            assert eng.monitorNames2 == orc.mon2_labels
            np.testing.assert_allclose(eng.mvSamples2.matrix(l), orc.samples2(l), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(eng.logProbsLadder(l), orc.logProbs(l), rtol=1e-10)
        np.testing.assert_allclose(eng.tempTrajLadder(l), orc.tempTraj(l), rtol=temps_rtol)
        np.testing.assert_array_equal(eng.accCountSwapLadder(l), orc.accCountSwap(l))
        sc = eng.scales(l)
        for k in range(eng.nTemps):
            for s in range(sc.shape[1]):
                assert np.isclose(sc[k, s], orc.scale(l, k, s), rtol=1e-9)


def all_reduce_sum(vec, dist=None):
    """Sum a small host vector over ranks with torch.distributed (gloo on CPU); identity without dist.  Test-side stand-in
    for apt_comm_allreduce_sum (NCCL inside the library), so that the sharding logic is testable without GPUs."""
    if dist is None or not dist.is_initialized():
        return np.asarray(vec, dtype=np.float64)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(np.asarray(vec, dtype=np.float64), device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
