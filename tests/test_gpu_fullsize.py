"""-m gpu: BASELINE.json's full-size configurations, checked through size-independent properties (the oracle cannot
run them in seconds): invariants of the algorithm (APT:494-496 Temps[1] fixed, sorted ladder, swap counts), exact
posterior of config 2, independence of a ladder's trajectory from how many other ladders run beside it, agreement of
two different reduction kernels on the N = 1e6 likelihood, and invariance under the rungs-per-pass setting."""
import numpy as np
import pytest

import models
import nimbleapt_b200 as nb
from helpers import build_engine

pytestmark = pytest.mark.gpu


def test_config2_full_size_invariants_and_posterior():
    spec = models.c2_mixture(n=20, K=32)
    L, niter, thin = 4096, 3000, 10
    eng = build_engine(spec, nLadders=L)
    eng.run(niter, thin=thin)
    K = eng.nTemps
    T = np.stack([eng.TempsLadder(l) for l in range(0, L, 97)])
    assert np.all(T[:, 0] == 1.0)                       # Temps[1] never adapts (APT:494-496)
    assert np.all(np.diff(T, axis=1) > 0)               # the ladder stays sorted
    assert np.all(T <= spec["ULT"] + K)                 # clamp ULT + ii (APT:500)
    acc = eng.accCountSwapLadder(0)
    assert acc.shape == (K, K) and np.all(acc >= 0) and np.all(np.diag(acc) == 0) and acc.sum() <= K * niter
    s = eng.mvSamples.matrix(-1)[:, 100:, :]
    assert s.shape == (L, niter // thin - 100, 2) and np.all(np.isfinite(s))
    y = spec["data"]["y"]
    n = y.shape[0]
    shrink = 1.0 / (1.0 + 1.0 / (n * 100.0 ** 2))
    for est, truth in ((np.abs(s[:, :, 0]).mean(1), y[:, 0].mean() * shrink), (s[:, :, 1].mean(1), y[:, 1].mean() * shrink),
                       ((s[:, :, 0] > 0).mean(1), 0.5)):
        mcse = est.std(ddof=1) / np.sqrt(L)
        assert abs(est.mean() - truth) < 4 * mcse, (est.mean(), truth, mcse)
    # a ladder's trajectory depends on its global id only, not on the other 4095 (sharding invariant)
    small = build_engine(spec, nLadders=8, ladder0=1000)
    small.run(niter, thin=thin)
    for l in (0, 7):
        assert np.array_equal(small.mvSamples.matrix(l), eng.mvSamples.matrix(1000 + l))


def test_config3_full_size_two_kernels_and_rungs_per_pass():
    spec = models.c3_logistic()
    a = build_engine(spec, engine=2)                     # 16 rungs per pass: 4 x 4 register tiles
    b = build_engine(spec, engine=2, rungsPerPass=8)     # 8 rungs per pass: 4 x 2 tiles, two passes per iteration
    for e in (a, b):
        e.run(25)
    # the same arithmetic in a different summation order: decisions identical, states equal to rounding
    np.testing.assert_allclose(a.mvSamples.matrix(0), b.mvSamples.matrix(0), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a.logProbsLadder(0), b.logProbsLadder(0), rtol=1e-11)
    np.testing.assert_allclose(np.asarray(a.Temps), np.asarray(b.Temps), rtol=1e-9)
    # the tracked log-probabilities (tile kernel, per-CTA partials) against a fresh evaluation by the wide kernel
    th = a.rungStates(0)
    tot, _ = a.calculate(th)
    lp = a.logProbTemps
    np.testing.assert_allclose(lp, tot, rtol=1e-11)
    assert np.all(np.isfinite(tot)) and np.all(np.diff(np.asarray(a.Temps)) > 0) and np.asarray(a.Temps)[0] == 1.0


_C5 = {}


@pytest.mark.parametrize("variant", ["two_chains_per_tile", "one_chain_per_tile"])
def test_config5_full_size(variant, monkeypatch):
    """Config 5 (logistic N = 1024, p = 8, 64 rungs) on 2048 ladders: invariants, the tracked log-probabilities
    (accumulated over 150 iterations of new-minus-old sums by the tile code) against a fresh evaluation of the final
    states, independence of a ladder from its neighbours, and agreement between the two ways the likelihood is
    evaluated (G::MULTI = 2: two chains per tile through one pass over the rows; = 1: chain by chain)."""
    monkeypatch.setenv("APTB200_MULTI", "2" if variant == "two_chains_per_tile" else "0")
    spec = models.c5_small_logistic()
    L, niter, thin = 2048, 150, 10
    eng = build_engine(spec, nLadders=L)
    assert ("MULTI = %d" % (2 if variant == "two_chains_per_tile" else 1)) in eng.lowered_source()
    eng.run(niter, thin=thin)
    K = eng.nTemps
    assert K == 64
    for l in (0, 777, L - 1):
        T = eng.TempsLadder(l)
        assert T[0] == 1.0 and np.all(np.diff(T) > 0) and np.all(T <= spec["ULT"] + K)
        th = eng.rungStates(l)
        tot, _ = eng.calculate(th)
        lp = eng._get(nb.api.APT_LOGPROBTEMPS, l, K)
        np.testing.assert_allclose(lp, tot, rtol=1e-10)
        acc = eng.accCountSwapLadder(l)
        assert np.all(acc >= 0) and np.all(np.diag(acc) == 0) and 0 < acc.sum() <= K * niter
    s = eng.mvSamples.matrix(-1)
    assert s.shape == (L, niter // thin, 8) and np.all(np.isfinite(s))
    _C5[variant] = (s[[0, 777, L - 1]].copy(), np.stack([eng.TempsLadder(l) for l in (0, 777, L - 1)]))
    if variant == "two_chains_per_tile":
        small = build_engine(spec, nLadders=2, ladder0=776)
        small.run(niter, thin=thin)
        assert np.array_equal(small.mvSamples.matrix(1), s[777])
    if len(_C5) == 2:  # same decisions, sums equal to rounding
        a, b = _C5["two_chains_per_tile"], _C5["one_chain_per_tile"]
        np.testing.assert_allclose(a[0], b[0], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(a[1], b[1], rtol=1e-8)


_C4 = {}


@pytest.mark.parametrize("variant", ["cluster", "one_cta"])
def test_config4_full_size(variant):
    """Config 4 (Poisson GLMM, 1000 random effects, N = 1e5, 24 rungs: a family of 1000 scalar samplers, a block sampler,
    a log-scale sampler) with 2 ladders: a ladder spanning a thread-block cluster (wide updates of the block sampler)
    against one CTA per ladder; tracked log-probabilities against a fresh evaluation; invariants."""
    spec = models.c4_glmm()
    kw = dict(clusterCtas=1) if variant == "one_cta" else {}
    eng = build_engine(spec, nLadders=2, **kw)
    src = eng.lowered_source()
    assert ("CLUSTER = 1" in src) == (variant == "one_cta")
    niter = 40
    eng.run(niter)
    K = eng.nTemps
    assert K == 24
    for l in range(2):
        T = eng.TempsLadder(l)
        assert T[0] == 1.0 and np.all(np.diff(T) > 0)
        th = eng.rungStates(l)
        assert np.all(th[:, 1000] > 0) and np.all(th[:, 1000] < 5)  # sigma inside dunif(0, 5)
        tot, _ = eng.calculate(th)
        np.testing.assert_allclose(eng._get(nb.api.APT_LOGPROBTEMPS, l, K), tot, rtol=1e-10)
    s = eng.mvSamples.matrix(-1)
    assert s.shape == (2, niter, 3) and np.all(np.isfinite(s))
    _C4[variant] = (s.copy(), eng.rungStates(1).copy())
    if len(_C4) == 2:
        np.testing.assert_allclose(_C4["cluster"][0], _C4["one_cta"][0], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(_C4["cluster"][1], _C4["one_cta"][1], rtol=1e-7, atol=1e-10)
