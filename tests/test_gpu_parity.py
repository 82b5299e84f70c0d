"""-m gpu: parity of the CUDA engine (called through the C ABI) against the CPU oracle on identical inputs.

The three north-star checks: (1) logProb on identical states within 1e-12 relative; (2) accept/reject and swap
decisions bit-identical when both sides are fed the same uniforms/normals (and also with the shared slot-addressed
Philox stream); (3) posterior summaries within 4 Monte-Carlo standard errors (test_gpu_posterior.py)."""
import numpy as np
import pytest

import models
import nimbleapt_b200 as nb
from helpers import assert_run_parity, build_engine, build_oracle, run_pair

pytestmark = pytest.mark.gpu

SMALL = {
    "c1": lambda: models.c1_vignette(nObs=100, K=8),
    "c2": lambda: models.c2_mixture(n=20, K=32),
    "c5": lambda: models.c5_small_logistic(N=256, p=8, K=8),
    # many rungs over shared data: each tile advances 4 chains together through one pass over the rows (G::MULTI,
    # apt_engine.cuh::rw_update_multi); N is not a multiple of the 4-row blocks, so the remainder path runs too
    "c5multi": lambda: models.c5_small_logistic(N=602, p=6, K=16),
    "c4": lambda: models.c4_glmm(G=20, per=10, K=4),
    "c3s": lambda: models.logistic(2000, 5, 4),
    "c3grid": lambda: models.logistic(2000, 5, 4),  # same model through the grid-wide engine (engine = 2)
    # grid engine, tile mode (warp tiles of 64 rows, lanes = row group x chain group): 16 rungs -> 8 rows x 4 chains per
    # thread, 4 rungs -> 4 rows x 2 chains; N is not a multiple of 64, so the remainder path runs too
    "c3tile16": lambda: models.logistic(3000, 15, 16, seed_off=5),
    "c3tile4": lambda: models.logistic(1500, 20, 4, seed_off=6),
    # ladder spanning a thread-block cluster (state in global memory, cluster barriers between sampler stages)
    "c4cluster": lambda: models.c4_glmm(G=20, per=10, K=4),
    # the reference's demo model (demo/APT_demo.R:83-92,173-177): dcat + dexp, a discrete node sampled by
    # sampler_slice_tempered, elements of constant arrays selected by that node, scalar + block RW samplers
    "mvn": lambda: models.mvn_prior(),  # dmnorm(mean, prec = ) and dmnorm(mean, cov = ) with constant matrices, n = 12
    "demo": lambda: models.demo_wrong_verbatim(),  # the demo's own text: wrongCode + the nimbleFunctions k2theta / k2sigma
    # sampler_RW_multinomial_tempered (APT:1005-1145) on a latent dmulti node + a log-scale scalar sampler
    "multinom": lambda: models.multinomial_latent(),
    "mixed": lambda: models.mixed_dists(),  # dbinom(size>1) dpois dgamma dbeta dexp dunif dnorm(tau) + all sampler options
}
BUILD_KW = {"c4cluster": dict(clusterCtas=4, lanesPerChain=32), "c3grid": dict(engine=2), "c3tile16": dict(engine=2), "c3tile4": dict(engine=2)}


@pytest.mark.parametrize("name", list(SMALL))
def test_logprob_matches_oracle(name):
    spec = SMALL[name]()
    eng = build_engine(spec, **BUILD_KW.get(name, {}))
    orc = build_oracle(spec)
    rng = np.random.default_rng(5)
    P = len(eng.paramNames)
    th0 = np.concatenate([np.ravel(spec["inits"][k], order="F") for k in spec["inits"]])
    assert th0.size == P
    states = th0[None, :] + 0.3 * rng.standard_normal((17, P))
    if name == "c4":
        states[:, 20] = np.abs(states[:, 20]) + 0.1  # sigma inside its support
    if name == "demo":
        states[:, 0] = rng.integers(1, 11, size=states.shape[0])  # k inside its support ...
        states[:3, 0] = [0.0, 11.0, 2.5]                           # ... and outside (dcat: -Inf on both sides)
        states[:, 2] = np.abs(states[:, 2]) + 0.05                 # sigma0 > 0
    if name == "multinom":
        states[:, :4] = rng.multinomial(30, [0.25] * 4, size=states.shape[0])  # compositions of N ...
        states[0, :4] = [30.0, 0.0, 0.0, 0.0]
        states[1, :4] = [10.0, 10.0, 5.0, 4.0]    # ... sum(x) != size, a non-integer and a negative cell: -Inf on both sides
        states[2, :4] = [7.5, 7.5, 7.0, 8.0]
        states[3, :4] = [-1.0, 11.0, 10.0, 10.0]
        states[:, 4] = np.abs(states[:, 4]) + 0.05
    if name == "mixed":
        states = np.abs(states) + 0.05
        states[:, 0] = np.clip(states[:, 0], 0.02, 0.98)  # p in (0,1)
        states[5, 0] = 1.5    # outside the support of dbeta: -Inf / NaN must agree too
        states[6, 3] = 25.0   # outside dunif(0.05, 20)
    tot, grp = eng.calculate(states)
    for i in range(states.shape[0]):
        ref, ref_d = orc.logprob(states[i])
        if not np.isfinite(ref):
            assert (np.isnan(ref) and np.isnan(tot[i])) or tot[i] == ref, (name, i, tot[i], ref)
            continue
        assert np.isclose(tot[i], ref, rtol=1e-12, atol=0), (name, i, tot[i], ref)
        assert np.allclose(np.sort(grp[i]), np.sort(ref_d[ref_d != 0]), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", list(SMALL))
@pytest.mark.parametrize("inject", [True, False])
def test_decisions_bit_identical(name, inject):
    spec = SMALL[name]()
    niter, L = 450, 2  # crosses two sampler adaptations (adaptInterval = 200)
    eng, orc, recs = run_pair(spec, niter, L=L, inject=inject, build_kw=BUILD_KW.get(name))
    assert_run_parity(eng, orc, recs, spec, name, niter, L)
    for l in range(L):
        if name == "multinom":  # u, timesRan ... RescaleThreshold of every rung (APT:1029-1040), adapted four times
            bs = eng.blockState(l)
            for k in range(eng.nTemps):
                np.testing.assert_allclose(bs[k, :1 + 8 * 16], orc.multinomialState(l, k, 0, 4), rtol=1e-9, atol=0)


@pytest.mark.parametrize("L,rpp", [(5, 0), (3, 2)])
def test_grid_engine_several_ladders_and_passes(L, rpp):
    """The persistent grid kernel when a pass holds several whole ladders and the last pass is short (5 ladders x 4 rungs, 16
    chains per pass: ladders 0-3, then ladder 4), and when a ladder spans two passes (3 ladders x 4 rungs, 2 chains per pass):
    control state mirrored per pass / read from global memory, swap phase of every ladder after the last pass, proposals for a
    different set of chains than the ones just finished.  Decisions bit for bit against the oracle, through an adaptation."""
    spec = models.logistic(600, 4, 4, seed_off=9)
    kw = dict(engine=2)
    if rpp:
        kw["rungsPerPass"] = rpp
    niter = 230
    eng, orc, recs = run_pair(spec, niter, L=L, inject=True, build_kw=kw)
    assert_run_parity(eng, orc, recs, spec, "grid L=%d" % L, niter, L)
