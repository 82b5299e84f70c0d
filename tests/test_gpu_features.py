"""-m gpu: the rest of the reference's run()/control surface against the oracle: continuation (reset=FALSE,
resetMV, APT:378-385), thin/thin2 + monitors2 (APT:514-539), temperPriors=FALSE (APT:713-718, 827-830),
log=TRUE and reflective=TRUE proposals (APT:694-708), adaptScaleOnly / adaptive=FALSE (APT:726, 852-868),
fixed ladder (adaptTemps=FALSE), resetTempering (APT:365-367), and posterior summaries within 4 MCSE."""
import numpy as np
import pytest

import models
import nimbleapt_b200 as nb
from helpers import build_engine, build_oracle, make_tables

pytestmark = pytest.mark.gpu


def _both(spec, L=1, **kw):
    return build_engine(spec, nLadders=L, record=True, **kw), build_oracle(spec, nLadders=L)


def _run_both(eng, orc, niter, seed, **kw):
    K, NU, DMAX = eng.nTemps, eng.info(nb.api.INFO_N_UNITS), eng.info(nb.api.INFO_DMAX)
    L = eng.nLadders
    Z, U, S = make_tables(niter, L, K, NU, DMAX, seed=seed)
    recs = [orc.set_tables(l, Z[:, l].copy(), U[:, l].copy(), S[:, l].copy(), record=True, niter=niter) for l in range(L)]
    eng.run(niter, _inject=dict(Z=Z, U=U, S=S), **kw)
    orc.run(niter, **{k: v for k, v in kw.items()})
    for l in range(L):
        assert np.array_equal(eng.decisions(l, niter), recs[l][0])
        assert np.array_equal(eng.swapRecord(l, niter) >> 16, recs[l][1] >> 16)
    return recs


def test_continuation_and_thinning():
    spec = dict(models.c2_mixture(n=20, K=8, tmax=200.0))
    spec["monitors2"] = ["theta[2]"]
    eng, orc = _both(spec, L=2)
    _run_both(eng, orc, 300, 1, thin=3, thin2=7)
    _run_both(eng, orc, 250, 2, reset=False, thin=5, thin2=7)            # appends (APT:378-380)
    for l in range(2):
        assert eng.mvSamples.matrix(l).shape == orc.samples(l).shape == (100 + 50, 2)
        np.testing.assert_allclose(eng.mvSamples.matrix(l), orc.samples(l), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(eng.mvSamples2.matrix(l), orc.samples2(l), rtol=1e-8, atol=1e-10)
        assert eng.mvSamples2.matrix(l).shape == (300 // 7 + 250 // 7, 1)
        assert eng.tempTrajLadder(l).shape == (50, 8)                     # tempTraj only covers the last run (APT:386)
        np.testing.assert_allclose(eng.logProbsLadder(l), orc.logProbs(l), rtol=1e-10)
    _run_both(eng, orc, 120, 3, reset=False, resetMV=True, thin=4)        # sample store restarts (APT:381-384)
    assert eng.mvSamples.matrix(0).shape == (30, 2)
    np.testing.assert_allclose(eng.mvSamples.matrix(1), orc.samples(1), rtol=1e-8, atol=1e-10)
    _run_both(eng, orc, 100, 4, reset=True, resetTempering=True, adaptTemps=False)  # fixed ladder after reset
    np.testing.assert_allclose(eng.TempsLadder(0), orc.Temps(0), rtol=1e-9)
    np.testing.assert_allclose(eng.modelState(0), orc.rung_state(0, 0), rtol=1e-8)  # APT:548


def test_set_inits_all_ladders_and_caller_allocated_traces():
    """apt_set_inits with ladder = -1 puts the same values into every ladder's model (one upload, replicated on the
    device), ladder = l only into that one; every rung of a ladder starts from them (APT:370-372).  Traces can be
    fetched into caller-allocated memory (the C ABI never allocates for the caller)."""
    spec = models.c2_mixture(n=20, K=8, tmax=200.0)
    L = 5
    eng = build_engine(spec, nLadders=L)
    th = np.array([1.25, -0.5])
    eng.setInits(th)
    eng.run(0)
    for l in range(L):
        np.testing.assert_array_equal(eng.rungStates(l), np.tile(th, (8, 1)))
    th3 = np.array([-2.0, 0.75])
    eng.setInits(th3, ladder=3)
    eng.run(0)
    for l in range(L):
        np.testing.assert_array_equal(eng.rungStates(l), np.tile(th3 if l == 3 else th, (8, 1)))
        np.testing.assert_array_equal(eng.modelState(l), th3 if l == 3 else th)
    with pytest.raises(nb.AptError, match="ladder index out of range"):
        eng.setInits(th, ladder=L)
    eng.run(60, thin=3)
    ref = eng.mvSamples.matrix(-1)
    buf = np.full(L * 20 * 2 + 7, np.nan)
    got = eng.mvSamples.matrix(-1, out=buf)
    assert got.shape == (L, 20, 2) and np.shares_memory(got, buf)
    np.testing.assert_array_equal(got, ref)
    assert np.all(np.isnan(buf[L * 20 * 2:]))
    with pytest.raises(ValueError, match="out must be"):
        eng.mvSamples.matrix(-1, out=np.empty(10))


@pytest.mark.parametrize("engine", ["ladder", "grid"])
def test_asynchronous_copy_back_of_samples(engine):
    """apt_run_args.samples_out: the rows a run writes arrive in the caller's host buffer piece by piece on a second
    stream while the run goes on; when run() returns they equal what apt_get(APT_SAMPLES) returns -- for launch chunks
    that are not multiples of the thinning interval, for a continued run (only its own rows), on both engines."""
    if engine == "ladder":
        spec, L, kw = models.c2_mixture(n=20, K=8, tmax=200.0), 5, {}
    else:
        spec, L, kw = models.logistic(2000, 5, 4), 1, dict(engine=2)
    eng = build_engine(spec, nLadders=L, **kw)
    nmon = len(eng.monitorNames)
    out = np.full(L * 57 * nmon + 3, np.nan)
    eng.run(403, thin=7, samplesOut=out)                       # default: at least 4 launches
    ref = eng.mvSamples.matrix(-1)
    assert ref.shape == (L, 57, nmon) and eng.getTimes()["main_launches"] >= 4
    np.testing.assert_array_equal(out[:L * 57 * nmon].reshape(L, 57, nmon), ref)
    assert np.all(np.isnan(out[L * 57 * nmon:]))
    out2 = np.full(L * 18 * nmon, np.nan)
    eng.run(130, reset=False, thin=7, itersPerLaunch=10, samplesOut=out2)  # chunks of 10 iterations, rows every 7
    full = eng.mvSamples.matrix(-1)
    assert full.shape == (L, 57 + 18, nmon)
    np.testing.assert_array_equal(out2.reshape(L, 18, nmon), full[:, 57:, :])
    np.testing.assert_array_equal(full[:, :57, :], ref)
    with pytest.raises(nb.AptError, match="samples_out: buffer too small"):
        eng.run(70, thin=7, samplesOut=np.empty(L * 10 * nmon - 1))
    with pytest.raises(ValueError, match="samplesOut must be"):
        eng.run(70, thin=7, samplesOut=np.empty((L, 10 * nmon)))


@pytest.mark.parametrize("variant", ["temperPriorsFalse", "logScale", "reflective", "scaleOnly", "nonAdaptive", "identityPropCov",
                                     "sliceContinuous", "slicePartial", "multinomNoTempering", "multinomNonAdaptive",
                                     "multinomInLoop"])
def test_control_options(variant):
    if variant in ("multinomNoTempering", "multinomNonAdaptive"):
        # sampler_RW_multinomial_tempered with useTempering = FALSE (APT:1074) / adaptive = FALSE (APT:1077)
        spec = dict(models.multinomial_latent(K=4))
        ctl = dict(useTempering=variant != "multinomNoTempering", adaptive=variant != "multinomNonAdaptive", adaptInterval=100)
        spec["samplers"] = [dict(spec["samplers"][0], control=ctl), spec["samplers"][1]]
    elif variant == "multinomInLoop":
        # two dmulti nodes declared in a loop (rows of a matrix), each with its own multinomial sampler: the prior term of a
        # sampler is ONE instance of the declaration (a PARTIAL group), 5 cells, different totals
        code = """{
        for (g in 1:2) {
            x[g, 1:5] ~ dmulti(prob = prob[1:5], size = N[g])
            for (j in 1:5) {
                y[g, j] ~ dpois(0.25 + a * x[g, j])
            }
        }
        a ~ dgamma(shape = 4, rate = 4)
        }"""
        x0 = np.array([[4.0, 4.0, 4.0, 4.0, 4.0], [1.0, 2.0, 3.0, 4.0, 2.0]])
        spec = dict(code=code, constants=dict(N=np.array([20.0, 12.0]), prob=np.array([0.3, 0.1, 0.2, 0.25, 0.15])),
                    data=dict(y=np.array([[3.0, 0.0, 6.0, 9.0, 1.0], [0.0, 2.0, 5.0, 1.0, 4.0]])),
                    inits=dict(x=x0, a=np.array(1.0)),
                    samplers=[dict(target="x[1, 1:5]", type="sampler_RW_multinomial_tempered", control=dict(useTempering=True, adaptInterval=100)),
                              dict(target="x[2, 1:5]", type="sampler_RW_multinomial_tempered", control=dict(useTempering=True, adaptInterval=100)),
                              dict(target="a", type="sampler_RW_tempered", control=dict(temperPriors=True, log=True, scale=0.3))],
                    monitors=["x", "a"], Temps=np.array([1.0, 2.0, 4.0]), ULT=1e6)
    elif variant == "sliceContinuous":
        # sampler_slice_tempered on a continuous node (APT:898-995): stepping out, shrinkage, width adaptation
        spec = dict(models.c2_mixture(n=20, K=4, tmax=30.0))
        spec["samplers"] = [dict(target="theta[1]", type="sampler_slice_tempered", control=dict(width=0.5, sliceMaxSteps=10, adaptInterval=50)),
                            dict(target="theta[2]", type="sampler_RW_tempered", control=dict(temperPriors=True))]
    elif variant == "slicePartial":
        # slice updates whose dependants are a subset of a declaration (one random-effect group) + a non-adaptive one
        spec = models.c4_glmm(G=6, per=8, K=4)
        samplers = [dict(target="u[%d]" % (k + 1), type="sampler_slice_tempered", control=dict(adaptive=k != 1, adaptInterval=40)) for k in range(3)]
        samplers += [dict(target="u[%d]" % (k + 1), type="sampler_RW_tempered", control=dict(temperPriors=True)) for k in range(3, 6)]
        samplers.append(dict(target="b[1:2]", type="sampler_RW_block_tempered", control=dict(temperPriors=True)))
        samplers.append(dict(target="sigma", type="sampler_RW_tempered", control=dict(temperPriors=True, log=True)))
        spec = dict(spec, samplers=samplers)
    elif variant in ("temperPriorsFalse", "logScale", "reflective"):
        spec = models.c4_glmm(G=6, per=8, K=4)
        samplers = [dict(target="u[%d]" % (k + 1), type="sampler_RW_tempered", control=dict(temperPriors=variant != "temperPriorsFalse")) for k in range(6)]
        samplers.append(dict(target="b[1:2]", type="sampler_RW_block_tempered", control=dict(temperPriors=variant != "temperPriorsFalse")))
        ctl = dict(temperPriors=variant != "temperPriorsFalse")
        if variant == "logScale":
            ctl["log"] = True
        if variant == "reflective":
            ctl["reflective"] = True
        samplers.append(dict(target="sigma", type="sampler_RW_tempered", control=ctl))
        spec = dict(spec, samplers=samplers)
    else:
        spec = models.logistic(300, 4, 4, propcov=variant != "identityPropCov")
        ctl = dict(spec["samplers"][0]["control"])
        if variant == "scaleOnly":
            ctl["adaptScaleOnly"] = True
        if variant == "nonAdaptive":
            ctl["adaptive"] = False
        ctl["adaptInterval"] = 50
        spec = dict(spec, samplers=[dict(spec["samplers"][0], control=ctl)])
    eng, orc = _both(spec)
    _run_both(eng, orc, 220, 7)
    np.testing.assert_allclose(eng.mvSamples.matrix(0), orc.samples(0), rtol=1e-8, atol=1e-10)
    sc = eng.scales(0)
    for k in range(eng.nTemps):
        for s in range(sc.shape[1]):
            assert np.isclose(sc[k, s], orc.scale(0, k, s), rtol=1e-9)


def test_posterior_within_4_mcse_c2():
    """Exact posterior of config 2: |theta1| ~ N(ybar1 s, s/n), theta2 ~ N(ybar2 s, s/n), both signs equally likely."""
    spec = models.c2_mixture(n=20, K=8, tmax=200.0)
    L, niter = 256, 6000
    eng = build_engine(spec, nLadders=L)
    eng.run(niter, thin=2)
    s = eng.mvSamples.matrix(-1)[:, 500:, :]
    y = spec["data"]["y"]
    n = y.shape[0]
    shrink = 1.0 / (1.0 + 1.0 / (n * 100.0 ** 2))
    for est, truth in ((np.abs(s[:, :, 0]).mean(1), y[:, 0].mean() * shrink), (s[:, :, 1].mean(1), y[:, 1].mean() * shrink),
                       ((s[:, :, 0] > 0).mean(1), 0.5), (s[:, :, 1].var(1, ddof=1), shrink / n)):
        mcse = est.std(ddof=1) / np.sqrt(L)
        assert abs(est.mean() - truth) < 4 * mcse, (est.mean(), truth, mcse)
    q = np.quantile(s[:, :, 1].reshape(-1), [0.05, 0.5, 0.95])
    from scipy import stats
    ref = stats.norm.ppf([0.05, 0.5, 0.95], y[:, 1].mean() * shrink, np.sqrt(shrink / n))
    np.testing.assert_allclose(q, ref, atol=0.01)


def test_posterior_matches_oracle_within_4_mcse_c1():
    """Config 1 (vignette model): engine vs oracle posterior summaries from independent ladders (different ladder ids)."""
    spec = models.c1_vignette(nObs=100, K=8)
    Lg, Lo, niter = 128, 8, 4000
    eng = build_engine(spec, nLadders=Lg, ladder0=1000)
    eng.run(niter, thin=2)
    orc = build_oracle(spec, nLadders=Lo)
    orc.run(niter, thin=2, nthreads=8)
    g = eng.mvSamples.matrix(-1)[:, 500:, :]
    o = np.stack([orc.samples(l)[500:] for l in range(Lo)])
    for f in (lambda a: np.abs(a[:, :, 0]).mean(1), lambda a: np.abs(a[:, :, 1]).mean(1), lambda a: (a[:, :, 0] > 0).mean(1),
              lambda a: a[:, :, 0].var(1)):
        eg, eo = f(g), f(o)
        se = np.sqrt(eg.var(ddof=1) / Lg + eo.var(ddof=1) / Lo)
        assert abs(eg.mean() - eo.mean()) < 4 * se, (eg.mean(), eo.mean(), se)


def test_posterior_demo_model_slice_sampler_within_4_mcse():
    """The reference's demo model (demo/APT_demo.R): the discrete node k is sampled by sampler_slice_tempered; its
    exact marginal posterior is known (tests/models.py).  64 independent ladders give the Monte-Carlo error."""
    spec = models.demo_wrong_verbatim()
    exact = models.demo_wrong_k_posterior(spec)
    L, niter, burn = 64, 6000, 1000
    eng = build_engine(spec, nLadders=L)
    eng.run(niter)
    freq = np.zeros((L, len(exact)))
    for l in range(L):
        k = eng.mvSamples.matrix(l)[burn:, 0]
        assert np.all(k == np.floor(k)) and k.min() >= 1 and k.max() <= len(exact)
        freq[l] = np.bincount(k.astype(int) - 1, minlength=len(exact)) / len(k)
    mean, mcse = freq.mean(axis=0), freq.std(axis=0, ddof=1) / np.sqrt(L)
    assert np.all(np.abs(mean - exact) <= 4 * mcse + 2e-3), (mean, exact, mcse)
    w = eng.scales(0)[:, 2]  # adapted slice widths (APT:982)
    assert np.all(w > 0) and not np.allclose(w, 1.0)


def test_multinomial_sampler_posterior_and_reset():
    """sampler_RW_multinomial_tempered (APT:1005-1145): posterior means of the engine agree with the oracle's restatement
    within 4 Monte-Carlo standard errors -- both carry the reference's proposal ratio (APT:1093-1095), whose stationary
    law is measurably off the enumerated posterior (tests/test_oracle_engine.py), so the exact posterior is only a
    loose bracket here.  A second run with reset = TRUE restores the adaptation matrices but keeps the recycled u
    (sampler$reset, APT:1133-1142)."""
    spec = models.multinomial_latent()
    ex, ea = models.multinomial_latent_posterior(spec)
    Lg, Lo, niter, burn = 64, 32, 10000, 1000
    eng = build_engine(spec, nLadders=Lg)
    eng.run(niter)
    orc = build_oracle(spec, nLadders=Lo, seed=4321)
    orc.run(niter, nthreads=8)
    g = np.array([eng.mvSamples.matrix(l)[burn:].mean(axis=0) for l in range(Lg)])
    o = np.array([orc.samples(l)[burn:].mean(axis=0) for l in range(Lo)])
    x = eng.mvSamples.matrix(0)[:, :4]
    assert np.all(x == np.floor(x)) and x.min() >= 0 and np.all(x.sum(axis=1) == 30)
    se = np.sqrt(g.var(axis=0, ddof=1) / Lg + o.var(axis=0, ddof=1) / Lo)
    assert np.all(np.abs(g.mean(axis=0) - o.mean(axis=0)) <= 4 * se), (g.mean(axis=0), o.mean(axis=0), se)
    assert np.all(np.abs(g.mean(axis=0) - np.append(ex, ea)) < 0.25)
    bs = eng.blockState(0)
    u_before = bs[:, 0].copy()
    EN = bs[0, 1 + 5 * 16:1 + 6 * 16].reshape(4, 4, order="F")
    assert np.all((u_before >= 0) & (u_before <= np.pi)) and np.allclose(EN, EN.T) and EN.max() > 1 and EN.max() <= 30
    eng.run(0, reset=True)
    bs = eng.blockState(0)
    np.testing.assert_array_equal(bs[:, 0], u_before)
    np.testing.assert_array_equal(bs[:, 1:1 + 5 * 16], 0.0)
    np.testing.assert_array_equal(bs[:, 1 + 5 * 16:1 + 7 * 16], 1.0)
    np.testing.assert_array_equal(bs[:, 1 + 7 * 16:1 + 8 * 16], 0.2)


@pytest.mark.parametrize("multi", [0, 4])
def test_multi_chain_tiles_switch(multi, monkeypatch):
    """G::MULTI (chains advanced together by one tile, apt_engine.cuh::rw_update_multi): the automatic choice (2) is
    covered by tests/test_gpu_parity.py[c5multi]; here the same model with the feature off and with 4 chains per tile,
    each against the oracle (decisions bit-identical, traces, adapted scales and proposal covariances)."""
    monkeypatch.setenv("APTB200_MULTI", str(multi))
    spec = models.c5_small_logistic(N=602, p=6, K=16)
    eng, orc = _both(spec, L=2)
    assert ("MULTI = %d" % max(multi, 1)) in eng.lowered_source()
    _run_both(eng, orc, 450, 9)
    for l in range(2):
        np.testing.assert_allclose(eng.mvSamples.matrix(l), orc.samples(l), rtol=1e-8, atol=1e-10)
        bs = eng.blockState(l)
        for k in range(16):
            assert np.isclose(eng.scales(l)[k, 0], orc.scale(l, k, 0), rtol=1e-9)
            np.testing.assert_allclose(bs[k, :36].reshape(6, 6, order="F"), orc.propCov(l, k, 0, 6), rtol=1e-7, atol=1e-12)


_CHAIN = {}


@pytest.mark.parametrize("variant", ["logfree", "literal"])
def test_swap_chain_without_logs_decides_like_the_literal_expression(variant, monkeypatch):
    """The proposal ratio of the swap move (APT:462) is formed from per-row log(1 / rowsum) instead of two logarithms per
    proposal (apt_engine.cuh::ladder_step); -DAPT_CHAIN_LITERAL keeps the reference's expression.  Both kernels must take
    the same 2.4 million swap decisions (partners and accept flags) and end in the same states, on an adapted 32-rung
    ladder whose hot rungs sit on the uniform fallback (APT:571-573) and whose numerators underflow to 0."""
    if variant == "literal":
        monkeypatch.setenv("APTB200_EXTRA_NVCC_FLAGS", "-DAPT_CHAIN_LITERAL")
    spec = models.c2_mixture(n=20, K=32)
    L, niter = 25, 3000
    eng = build_engine(spec, nLadders=L, record=True)
    eng.run(niter, thin=10)
    rec = np.stack([eng.swapRecord(l, niter) for l in range(L)])
    assert rec.shape == (L, niter, 32)
    _CHAIN[variant] = (rec, eng.mvSamples.matrix(-1), np.stack([eng.TempsLadder(l) for l in range(L)]))
    if len(_CHAIN) == 2:
        a, b = _CHAIN["logfree"], _CHAIN["literal"]
        assert (a[0] >> 16).mean() > 0.05                      # swaps do get accepted
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
        np.testing.assert_array_equal(a[2], b[2])


def test_c4_family_medium():
    """Mixed scalar family / block / log-scale samplers at a size where the family spans several tiles."""
    spec = models.c4_glmm(G=60, per=20, K=6)
    eng, orc = _both(spec)
    _run_both(eng, orc, 210, 9)
    np.testing.assert_allclose(eng.mvSamples.matrix(0), orc.samples(0), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(eng.logProbsLadder(0), orc.logProbs(0), rtol=1e-9)


def test_c4_cluster_mode_wide_updates():
    """The same model with the ladder spanning a cluster of 4 CTAs: family members over all CTAs, the block sampler's
    1200 Poisson terms evaluated cluster-wide one chain at a time (rw_update_wide), state in global memory."""
    spec = models.c4_glmm(G=60, per=20, K=6)
    eng, orc = _both(spec, clusterCtas=4, lanesPerChain=32)
    assert "rw_update_wide" in eng.lowered_source()
    _run_both(eng, orc, 210, 9)
    np.testing.assert_allclose(eng.mvSamples.matrix(0), orc.samples(0), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(eng.logProbsLadder(0), orc.logProbs(0), rtol=1e-9)


def test_grid_engine_mid_size_against_oracle():
    """Grid engine on a 100k-row logistic model, 16 rungs in one pass: decisions vs the oracle for 30 iterations."""
    spec = models.logistic(100000, 50, 16, seed_off=3)
    eng, orc = _both(spec, engine=2)
    _run_both(eng, orc, 30, 11)
    np.testing.assert_allclose(eng.logProbsLadder(0), orc.logProbs(0), rtol=1e-11)
    th = eng.rungStates(0)
    tot, _ = eng.calculate(th)
    for k in (0, 7, 15):
        ref, _ = orc.logprob(th[k])
        assert abs(tot[k] - ref) <= 1e-12 * abs(ref)


@pytest.mark.parametrize("kind", ["scalar", "family"])
def test_start_outside_the_support_recovers_like_per_node_logprobs(kind):
    """An initial value outside a prior's support makes ONE node's logProb -Inf.  nimble stores logProbs per node, so
    the first accepted move into the support (logMHR = +Inf) leaves finite sums behind (APT:711,720-722); the engine
    stores one sum per declaration and must not turn -Inf + Inf into a NaN that poisons logProbTemps, the swap phase and
    the ladder adaptation (APT:460-462, 494-499).  Scalar samplers on single elements of a prior declaration (PARTIAL
    groups) and a family of 12 of them; compared with the oracle decision by decision."""
    rng = np.random.default_rng(models.DATA_SEED + 77)
    if kind == "scalar":
        n = 15
        y = np.stack([rng.standard_normal(n) + 1.0, rng.standard_normal(n)], axis=1)
        code = """{
        for (i in 1:n) {
            y[i,1] ~ dnorm(theta[1], sd=1)
            y[i,2] ~ dnorm(theta[2], sd=1)
        }
        for (j in 1:2) {
            theta[j] ~ dunif(-4, 4)
        }
        }"""
        spec = dict(code=code, constants=dict(n=n), data=dict(y=y), inits=dict(theta=np.array([4.5, -4.6])),  # both outside
                    samplers=[dict(target="theta[1]", type="sampler_RW_tempered", control=dict(temperPriors=True, scale=2.0)),
                              dict(target="theta[2]", type="sampler_RW_tempered", control=dict(temperPriors=True, scale=2.0))],
                    monitors=["theta"], Temps=np.exp(np.linspace(0.0, np.log(30.0), 6)), ULT=1e6)
    else:
        G, per = 12, 4
        g = np.repeat(np.arange(1, G + 1), per).astype(float)
        y = rng.standard_normal(G * per) * 0.7
        code = """{
        for (i in 1:N) {
            y[i] ~ dnorm(u[g[i]], sd=1)
        }
        for (k in 1:G) {
            u[k] ~ dunif(-3, 3)
        }
        }"""
        u0 = np.zeros(G); u0[[0, 5, 11]] = [3.4, -3.5, 4.0]  # three members start outside
        spec = dict(code=code, constants=dict(N=G * per, G=G, g=g), data=dict(y=y), inits=dict(u=u0),
                    samplers=[dict(target="u[%d]" % (k + 1), type="sampler_RW_tempered", control=dict(temperPriors=True, scale=1.5)) for k in range(G)],
                    monitors=["u"], Temps=np.exp(np.linspace(0.0, np.log(20.0), 5)), ULT=1e6)
    eng, orc = _both(spec, L=2)
    tot0, _ = eng.calculate(np.concatenate([np.ravel(spec["inits"][k]) for k in spec["inits"]]))
    assert tot0[0] == -np.inf
    # While two neighbouring rungs both sit at logProb -Inf the reference's own ladder adaptation produces NaN
    # temperatures (APT:494-496: exp((-Inf) - (-Inf))), so the ladder is held fixed until every rung has recovered
    _run_both(eng, orc, 150, 9, adaptTemps=False)
    for l in range(2):
        assert eng.logProbsLadder(l)[0] == -np.inf and np.isfinite(eng.logProbsLadder(l)[-1])
        assert np.all(np.isfinite(eng._get(nb.api.APT_LOGPROBTEMPS, l, eng.nTemps)))
    _run_both(eng, orc, 110, 10, reset=False)
    for l in range(2):
        lp = eng.logProbsLadder(l)
        assert lp.shape == (260,) and np.all(np.isfinite(lp[150:])) and np.all(np.isfinite(eng.TempsLadder(l)))
        np.testing.assert_allclose(lp, orc.logProbs(l), rtol=1e-9)
        np.testing.assert_allclose(eng.mvSamples.matrix(l), orc.samples(l), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(eng.TempsLadder(l), orc.Temps(l), rtol=1e-9)
        st = eng.rungStates(l)
        tot, _ = eng.calculate(st)
        np.testing.assert_allclose(eng._get(nb.api.APT_LOGPROBTEMPS, l, eng.nTemps), tot, rtol=1e-10)
