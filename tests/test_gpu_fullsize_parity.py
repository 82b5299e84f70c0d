"""-m gpu: oracle parity on BASELINE.json's configurations AT THEIR REAL SIZES (slow: the CPU oracle needs a few seconds
to a minute per case).  The small-model parity tests (test_gpu_parity.py) never reach the code paths that exist only at
these sizes:

  config 5  K = 64 rungs      : two swap-matrix columns per lane, CTA-wide team, no log table in the swap chain
                                 (apt_engine.cuh::ladder_step, CH = 2), two chains per tile (G::MULTI = 2)
  config 4  G = 1000, K = 24  : a sampler family of 1000 members, thread-block clusters, wide updates of the block sampler
  config 3  N = 1e6, p = 50   : the grid engine's tile kernel over 408 MB of data, 16 rungs per pass

Each is compared with the restated reference (oracle/apt_oracle.c after APT_functions.R:413-546, 556-576, 690-747,
821-872) decision by decision, like the small models."""
import numpy as np
import pytest

import models
from helpers import assert_run_parity, build_engine, build_oracle, run_pair

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


@pytest.mark.parametrize("multi", [2, 0])
@pytest.mark.parametrize("inject", [True, False])
def test_c5full_k64_decisions_bit_identical(inject, multi, monkeypatch):
    """BASELINE configs[4] per ladder: logistic N = 1024, p = 8, 64 temperatures (APT:452-484 with nTemps = 64)."""
    monkeypatch.setenv("APTB200_MULTI", str(multi))
    spec = models.c5_small_logistic()  # N = 1024, p = 8, K = 64
    niter, L = 450, 2                  # through two covariance adaptations (adaptInterval = 200, APT:855-871)
    eng, orc, recs = run_pair(spec, niter, L=L, inject=inject)
    assert eng.nTemps == 64
    assert ("MULTI = %d" % (2 if multi == 2 else 1)) in eng.lowered_source()
    assert_run_parity(eng, orc, recs, spec, "c5full", niter, L)


@pytest.mark.parametrize("variant,inject", [("cluster", True), ("one_cta", False)])
def test_c4full_g1000_k24_decisions_bit_identical(variant, inject):
    """BASELINE configs[3]: Poisson GLMM, 1000 random effects (N = 1e5), 24 temperatures; 1000 scalar samplers as one
    family + a block sampler + a log-scale sampler, 1002 tempered updates per rung and iteration (APT:435-448)."""
    spec = models.c4_glmm()  # G = 1000, per = 100, K = 24
    niter, L = 25, 1
    kw = dict(clusterCtas=1) if variant == "one_cta" else {}
    eng, orc, recs = run_pair(spec, niter, L=L, inject=inject, build_kw=kw)
    assert eng.nTemps == 24
    assert ("CLUSTER = 1" in eng.lowered_source()) == (variant == "one_cta")
    assert_run_parity(eng, orc, recs, spec, "c4full-" + variant, niter, L)


def test_c3full_logprob_matches_oracle_at_n_1e6():
    """North-star check 1 on BASELINE configs[2]: logistic N = 1e6, p = 50 — whole-model logProb of the device (wide
    reduction kernel, tree sums) against the oracle's node terms, 1e-12 relative.  model$getLogProb() adds the 1e6 node
    terms one after the other in double; that sequential sum is itself ~1e-11 off what the terms add up to (beta = 0:
    every node is log(1/2) exactly, the true total is known in closed form), so the device is held to 1e-12 against the
    oracle's node terms summed in long double, and the literal sequential sum is only required to sit within 1e-10."""
    spec = models.c3_logistic()
    eng = build_engine(spec, engine=2)
    orc = build_oracle(spec)
    rng = np.random.default_rng(5)
    P = len(eng.paramNames)
    assert P == 50
    states = np.stack([np.zeros(P), 0.25 * rng.standard_normal(P), 0.6 * rng.standard_normal(P)])
    tot, grp = eng.calculate(states)
    # closed form at beta = 0: N log(1/2) + P dnorm(0, 0, sd = 10, log = TRUE)
    exact0 = 1000000 * np.log(0.5) + P * (-0.5 * np.log(2 * np.pi) - np.log(10.0))
    assert abs(tot[0] - exact0) <= 1e-13 * abs(exact0), (tot[0], exact0)
    for i in range(states.shape[0]):
        ref, ref_d = orc.logprob(states[i], extended=True)
        lit, _ = orc.logprob(states[i])
        assert np.isfinite(ref)
        assert np.isclose(tot[i], ref, rtol=1e-12, atol=0), (i, tot[i], ref, (tot[i] - ref) / ref)
        assert np.allclose(np.sort(grp[i]), np.sort(ref_d[ref_d != 0]), rtol=1e-12, atol=0)
        assert np.isclose(lit, ref, rtol=1e-10, atol=0), (i, lit, ref)


def test_c3full_tile_kernel_decisions_bit_identical():
    """North-star check 2 on BASELINE configs[2]: 12 iterations x 16 rungs of sampler_RW_block_tempered through the grid
    engine's tile kernel (16 rungs per pass, one launch per iteration) with injected normals / uniforms, against the
    oracle's serial passes (APT:825: one pass over all 1e6 nodes per tempered update)."""
    spec = models.c3_logistic()
    niter = 12
    eng, orc, recs = run_pair(spec, niter, L=1, inject=True, build_kw=dict(engine=2))
    assert eng.info(15) == 2  # INFO_ENGINE: the grid engine
    # Decisions, swap partners and swap counts must be identical; states agree to 1e-8.  The ladder is compared at 1e-6:
    # its adaptation (APT:494-499) feeds on DIFFERENCES of log-probabilities of size 7e5, and the oracle adds the 1e6 node
    # terms one after the other like getLogProb() does, which costs ~1e-11 relative = ~7e-6 absolute on each sum (see the
    # test above) -- 1e-8 relative on a temperature after 12 iterations.
    assert_run_parity(eng, orc, recs, spec, "c3full", niter, 1, temps_rtol=1e-6)
    # the tracked log-probabilities (accumulated by the tile kernel) against the oracle's node terms added in long double
    lp = eng.logProbTemps
    ref = np.array([orc.logprob(orc.rung_state(0, k), extended=True)[0] for k in range(eng.nTemps)])
    np.testing.assert_allclose(lp, ref, rtol=1e-10)
