"""calculateWAIC (APT_functions.R:593-635) and logProb_* monitors (demo/APT_demo.R:180-181).
CPU: the oracle's restatement is pinned by an independent scipy computation; the monitor validity check
(mcmc_checkWAICmonitors, APT_functions.R:1313-1341) of the lowering step.  GPU: engine vs oracle through the C ABI."""
import math

import numpy as np
import pytest
from scipy import stats

import models
import nimbleapt_b200 as nb
from helpers import build_engine, build_oracle, run_pair


def _waic_from_matrix(lp):
    """WAIC from a samples x dataNodes matrix of log predictive densities, exactly as APT:623-630 writes it."""
    mx = lp.max(axis=0)
    lavg = np.sum(mx + np.log(np.mean(np.exp(lp - mx), axis=0)))
    pw = np.sum(np.var(lp, axis=0, ddof=1))
    return -2 * (lavg - pw), lavg, pw


def test_oracle_waic_against_scipy():
    spec = models.c2_mixture(n=20, K=8, tmax=200.0)
    orc = build_oracle(spec, 1)
    orc.run(600, thin=3)
    S = orc.samples(0)  # theta[1], theta[2]
    nb_post = math.ceil(100 / 3)  # nburninPostThinning, APT:605
    S = S[nb_post:]
    y = spec["data"]["y"]
    # data nodes in declaration order: y[i,1], y[i,2] interleaved (one loop, two declarations -> all y[,1] then all y[,2])
    lp1 = stats.norm.logpdf(y[None, :, 0], loc=np.abs(S[:, 0])[:, None], scale=1.0)
    lp2 = stats.norm.logpdf(y[None, :, 1], loc=S[:, 1][:, None], scale=1.0)
    want = _waic_from_matrix(np.concatenate([lp1, lp2], axis=1))
    got = orc.calculateWAIC(nburnin=100, thin=3)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    # fewer than two post-burn-in samples: -Inf (APT:607-610)
    assert orc.calculateWAIC(nburnin=599, thin=3)[0] == -np.inf


def test_oracle_logprob_monitors_sum_to_total():
    spec = models.demo_wrong()
    orc = build_oracle(spec, 1)
    orc.run(60, thin=1, thin2=1)
    assert orc.mon2_labels[:4] == ["logProb_k", "logProb_theta0", "logProb_sigma0", "logProb_y[1]"]
    np.testing.assert_allclose(orc.samples2(0).sum(axis=1), orc.logProbs(0), rtol=1e-12)


def _conf(spec, monitors, **kw):
    m = nb.nimbleModel(nb.nimbleCode(spec["code"]), constants=spec["constants"], data=spec["data"], inits=spec["inits"])
    conf = nb.configureMCMC(m, monitors=monitors, **kw)
    conf.removeSamplers()
    for s in spec["samplers"]:
        conf.addSampler(s["target"], type=s["type"], control=s["control"])
    return conf


def test_waic_monitor_check_of_the_lowering():
    spec = models.c4_glmm(G=20, per=10, K=4)
    # every parameter monitored: valid (compiles without a GPU)
    apt = nb.buildAPT(_conf(spec, ["b", "sigma", "u"], enableWAIC=True), Temps=spec["Temps"], compileOnly=True)
    assert "NDATA = 200" in apt.lowered_source()
    # b is a top-level parameter of the data nodes and is not monitored (APT:1323-1334)
    with pytest.raises(nb.AptError, match="following data nodes have un-monitored upstream parameters:\n y"):
        nb.buildAPT(_conf(spec, ["sigma", "u"]), Temps=spec["Temps"], compileOnly=True, enableWAIC=True)
    # u is downstream of the monitored sigma: the reference would simulate it from its prior (APT:618); refused
    with pytest.raises(nb.AptError, match="'u' is downstream of monitored nodes"):
        nb.buildAPT(_conf(spec, ["b", "sigma"], enableWAIC=True), Temps=spec["Temps"], compileOnly=True)
    orc = build_oracle(dict(spec, monitors=["sigma", "u"]), 1)
    with pytest.raises(ValueError, match="un-monitored upstream parameters: y"):
        orc.calculateWAIC()
    # no data nodes (APT:331)
    code = "{ a ~ dnorm(0, 1)\n b ~ dnorm(a, 1) }"
    m = nb.nimbleModel(code, inits=dict(a=0.0, b=0.0))
    conf = nb.configureMCMC(m, monitors=["a", "b"], enableWAIC=True)
    for t in ("a", "b"):
        conf.addSampler(t, type="sampler_RW_tempered", control=dict(temperPriors=True))
    with pytest.raises(nb.AptError, match="no data nodes were detected"):
        nb.buildAPT(conf, Temps=[1, 2], compileOnly=True)


WAIC_CASES = {
    "c1": lambda: (models.c1_vignette(nObs=100, K=5), {}),
    "c2": lambda: (models.c2_mixture(n=20, K=8, tmax=200.0), {}),
    "demo": lambda: (dict(models.demo_wrong(), monitors2=[]), {}),
    "c4": lambda: (dict(models.c4_glmm(G=20, per=10, K=4), monitors=["b", "sigma", "u"]), {}),
    "c3grid": lambda: (models.logistic(2000, 5, 4), dict(engine=2)),
    "mixed": lambda: (models.mixed_dists(), {}),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(WAIC_CASES))
def test_waic_matches_oracle(name):
    spec, kw = WAIC_CASES[name]()
    niter, thin, L = 330, 3, 2
    eng, orc, _ = run_pair(spec, niter, L=L, inject=False, thin=thin, record=False, build_kw=dict(kw, enableWAIC=True))
    for l in range(L):
        np.testing.assert_allclose(eng.mvSamples.matrix(l), orc.samples(l), rtol=1e-8, atol=1e-10)
        for nburnin in (0, 100):
            want = orc.calculateWAIC(nburnin=nburnin, ladder=l, thin=thin)
            got = eng.calculateWAIC(nburnin=nburnin, ladder=l, details=True)
            np.testing.assert_allclose(got, want, rtol=1e-9)
    assert eng.calculateWAIC(nburnin=niter - 1) == -np.inf  # APT:607-610
    # continuation appends rows (reset=FALSE): WAIC is over all of mvSamples (APT:606)
    eng.run(90, reset=False, thin=thin); orc.run(90, reset=False, thin=thin)
    np.testing.assert_allclose(eng.calculateWAIC(nburnin=30, details=True), orc.calculateWAIC(nburnin=30, thin=thin), rtol=1e-9)


@pytest.mark.gpu
def test_waic_disabled_and_pooled():
    spec = models.c2_mixture(n=20, K=8, tmax=200.0)
    eng = build_engine(spec, nLadders=3)
    eng.run(200, thin=2)
    assert math.isnan(eng.calculateWAIC())  # APT:596-599
    out = np.empty(3)
    assert nb._lib.lib().apt_waic(eng._h, 0, 0, nb.api._dp(out)) != 0
    assert "enableWAIC = TRUE" in nb._lib.lib().apt_last_error().decode()
    eng = build_engine(spec, nLadders=3, enableWAIC=True)
    eng.run(200, thin=2)
    # ladder = -1 pools the samples of all ladders: same as WAIC of the stacked sample matrix
    S = np.concatenate([eng.mvSamples.matrix(l)[10:] for l in range(3)])
    y = spec["data"]["y"]
    lp = np.concatenate([stats.norm.logpdf(y[None, :, 0], loc=np.abs(S[:, 0])[:, None]),
                         stats.norm.logpdf(y[None, :, 1], loc=S[:, 1][:, None])], axis=1)
    np.testing.assert_allclose(eng.calculateWAIC(nburnin=20, ladder=-1, details=True), _waic_from_matrix(lp), rtol=1e-10)
