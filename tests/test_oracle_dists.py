"""CPU: pins the oracle's Rmath-style log-densities (oracle/rmath_port.c) against scipy.stats — an independent
implementation — because the reference ships no golden vectors for this path."""
import numpy as np
import pytest
from scipy import special, stats

from oracle.oracle import lib

RTOL = 1e-12


def close(a, b, rtol=RTOL, atol=1e-13):
    if np.isinf(b) or np.isinf(a):
        return a == b
    return abs(a - b) <= atol + rtol * abs(b)


def test_dnorm_grid():
    L = lib()
    rng = np.random.default_rng(1)
    for _ in range(2000):
        x, mu, sd = rng.normal(0, 50), rng.normal(0, 50), np.exp(rng.normal(0, 3))
        assert close(L.apt_dnorm_log(x, mu, sd), stats.norm.logpdf(x, mu, sd))
    assert L.apt_dnorm_log(1.0, 1.0, 0.0) == np.inf and L.apt_dnorm_log(2.0, 1.0, 0.0) == -np.inf
    assert np.isnan(L.apt_dnorm_log(0.0, 0.0, -1.0))


def test_dunif_edges():
    L = lib()
    assert close(L.apt_dunif_log(0.3, 0.0, 5.0), -np.log(5.0))
    assert L.apt_dunif_log(5.0, 0.0, 5.0) == -np.log(5.0)  # closed support like Rmath
    assert L.apt_dunif_log(5.0000001, 0.0, 5.0) == -np.inf
    assert np.isnan(L.apt_dunif_log(0.0, 1.0, 1.0))


def test_dgamma_grid():
    L = lib()
    rng = np.random.default_rng(2)
    for _ in range(3000):
        shape, scale = np.exp(rng.normal(0, 2.5)), np.exp(rng.normal(0, 2))
        x = stats.gamma.rvs(shape, scale=scale, random_state=rng) if rng.random() < 0.7 else np.exp(rng.normal(0, 4))
        if x < 1e-290 or not np.isfinite(x):  # denormal x: x/scale itself is no longer a 53-bit quantity
            continue
        ref = stats.gamma.logpdf(x, shape, scale=scale)
        assert close(L.apt_dgamma_log(x, shape, scale), ref, rtol=2e-12), (x, shape, scale)
    assert L.apt_dgamma_log(-1.0, 2.0, 1.0) == -np.inf
    assert L.apt_dgamma_log(0.0, 0.5, 1.0) == np.inf and L.apt_dgamma_log(0.0, 2.0, 1.0) == -np.inf
    assert close(L.apt_dgamma_log(0.0, 1.0, 3.0), -np.log(3.0))


def test_dbeta_grid():
    L = lib()
    rng = np.random.default_rng(3)
    for _ in range(3000):
        a, b = np.exp(rng.normal(0, 2)), np.exp(rng.normal(0, 2))
        x = rng.beta(a, b)
        if x <= 0 or x >= 1:
            continue
        assert close(L.apt_dbeta_log(x, a, b), stats.beta.logpdf(x, a, b), rtol=5e-12, atol=5e-12), (x, a, b)
    assert L.apt_dbeta_log(1.5, 2.0, 2.0) == -np.inf
    assert L.apt_dbeta_log(0.0, 2.0, 2.0) == -np.inf and L.apt_dbeta_log(0.0, 0.5, 2.0) == np.inf


def test_dbinom_grid_and_edges():
    L = lib()
    rng = np.random.default_rng(4)
    for _ in range(3000):
        n = int(rng.integers(0, 500))
        p = rng.random() if rng.random() < 0.8 else rng.random() * 1e-6
        x = int(rng.integers(0, n + 1))
        assert close(L.apt_dbinom_log(float(x), float(n), p), stats.binom.logpmf(x, n, p), rtol=5e-12), (x, n, p)
    for p in (1e-12, 0.05, 0.5, 0.95, 1 - 1e-12):  # Bernoulli edge cases used by the logistic workloads
        assert close(L.apt_dbinom_log(0.0, 1.0, p), np.log1p(-p))
        assert close(L.apt_dbinom_log(1.0, 1.0, p), np.log(p))
    assert L.apt_dbinom_log(0.0, 5.0, 0.0) == 0.0 and L.apt_dbinom_log(1.0, 5.0, 0.0) == -np.inf
    assert L.apt_dbinom_log(5.0, 5.0, 1.0) == 0.0 and L.apt_dbinom_log(6.0, 5.0, 0.5) == -np.inf
    assert L.apt_dbinom_log(0.5, 5.0, 0.5) == -np.inf and np.isnan(L.apt_dbinom_log(1.0, 5.0, 1.5))


def test_dpois_grid_and_edges():
    L = lib()
    rng = np.random.default_rng(5)
    for _ in range(3000):
        lam = np.exp(rng.normal(1, 3))
        x = float(rng.poisson(min(lam, 1e6))) if rng.random() < 0.7 else float(rng.integers(0, 2000))
        assert close(L.apt_dpois_log(x, lam), stats.poisson.logpmf(x, lam), rtol=5e-12), (x, lam)
    assert L.apt_dpois_log(0.0, 0.0) == 0.0 and L.apt_dpois_log(3.0, 0.0) == -np.inf
    assert L.apt_dpois_log(-1.0, 2.0) == -np.inf and np.isnan(L.apt_dpois_log(1.0, -2.0))
    import mpmath as mp  # scipy's x log(lam) - lam - gammaln(x+1) cancels catastrophically here; use 40-digit arithmetic
    mp.mp.dps = 40
    ref = float(1e7 * mp.log(mp.mpf(1e7)) - mp.mpf(1e7) - mp.loggamma(mp.mpf(1e7) + 1))
    assert close(L.apt_dpois_log(1e7, 1e7), ref, rtol=1e-13)


def test_dexp_stirlerr_bd0_lbeta():
    L = lib()
    assert close(L.apt_dexp_log(2.0, 4.0), stats.expon.logpdf(2.0, scale=4.0))
    import mpmath as mp
    mp.mp.dps = 40
    for n in (0.5, 1.0, 7.5, 15.0, 16.0, 40.0, 90.0, 600.0, 3.3):
        m = mp.mpf(n)
        ref = float(mp.loggamma(m + 1) - (m + mp.mpf("0.5")) * mp.log(m) + m - mp.log(mp.sqrt(2 * mp.pi)))
        assert abs(L.apt_stirlerr(n) - ref) <= 2e-12 * abs(ref) + 1e-17, (n, L.apt_stirlerr(n), ref)
    for x, np_ in ((3.0, 3.1), (10.0, 2.0), (1.0, 0.95), (1e5, 1e5 + 7)):
        X, M = mp.mpf(x), mp.mpf(np_)
        assert close(L.apt_bd0(x, np_), float(X * mp.log(X / M) + M - X), rtol=1e-13, atol=1e-18)
    for a, b in ((0.5, 0.7), (3.0, 12.0), (20.0, 30.0), (1e-3, 400.0)):
        assert close(L.apt_lbeta(a, b), special.betaln(a, b), rtol=1e-12, atol=1e-12)


def test_dcat_unnormalised_and_edges():
    """nimble's dcat (demo/APT_demo.R:84): log(prob[x] / sum(prob)); -Inf outside 1..K and for non-integers."""
    import ctypes as C
    L = lib()
    L.apt_dcat_log.restype = C.c_double
    L.apt_dcat_log.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_int]
    p = np.array([0.2, 1.5, 0.0, 3.3, 0.7])
    pp = p.ctypes.data_as(C.POINTER(C.c_double))
    for x in range(1, 6):
        ref = np.log(p[x - 1] / p.sum()) if p[x - 1] > 0 else -np.inf
        got = L.apt_dcat_log(float(x), pp, 5)
        assert got == ref or close(got, ref, rtol=1e-15)
    for x in (0.0, 6.0, -1.0, 2.5):
        assert L.apt_dcat_log(x, pp, 5) == -np.inf
    assert np.isnan(L.apt_dcat_log(float("nan"), pp, 5))


def test_dmulti_and_rbinom_inversion():
    """nimble's dmulti (target of sampler_RW_multinomial_tempered, APT:1046) against scipy; rbinom by inversion of one
    uniform (include/aptb200.h) against the quantile function."""
    import ctypes as C
    from scipy import stats
    L = lib()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(3)
    for K in (2, 4, 9):
        p = rng.random(K) + 0.05
        for n in (0, 1, 7, 60):
            x = rng.multinomial(n, p / p.sum()).astype(float)
            got = L.apt_dmulti_log(x.ctypes.data_as(dp), p.ctypes.data_as(dp), K, float(n))  # unnormalised prob
            assert close(got, stats.multinomial.logpmf(x, n, p / p.sum()), rtol=1e-12, atol=1e-12)
            if n:
                assert L.apt_dmulti_log(x.ctypes.data_as(dp), p.ctypes.data_as(dp), K, float(n + 1)) == -np.inf
        x = np.array([1.5] + [0.0] * (K - 1))
        assert L.apt_dmulti_log(x.ctypes.data_as(dp), p.ctypes.data_as(dp), K, 1.5) == -np.inf
    p0 = np.array([0.0, 1.0, 1.0])
    x0 = np.array([0.0, 2.0, 3.0])
    assert close(L.apt_dmulti_log(x0.ctypes.data_as(dp), p0.ctypes.data_as(dp), 3, 5.0),
                 stats.multinomial.logpmf(x0, 5, p0 / 2), rtol=1e-12, atol=1e-12)
    for n in (1, 5, 37, 400, 1000):
        for pr in (0.003, 0.2, 0.5, 0.77, 0.999):
            u = rng.random(200)
            got = np.array([L.apt_rbinom_inv(float(n), pr, ui) for ui in u])
            # prob > 0.5: the walk runs on n - X ~ Binom(n, 1 - prob) (fewer steps) and n minus its quantile is returned
            q, walked = (1 - pr, n - got) if pr > 0.5 else (pr, got)
            ref = stats.binom.ppf(u, n, q) if pr <= 0.5 else n - stats.binom.ppf(u, n, q)
            cdf_lo = stats.binom.cdf(walked - 1, n, q)
            cdf_hi = stats.binom.cdf(walked, n, q)
            near_edge = (np.abs(u - cdf_lo) < 1e-9) | (np.abs(u - cdf_hi) < 1e-9)
            assert np.all((got == ref) | near_edge), (n, pr)
    assert L.apt_rbinom_inv(0.0, 0.3, 0.5) == 0.0 and L.apt_rbinom_inv(9.0, 1.0, 0.5) == 9.0 and L.apt_rbinom_inv(9.0, 0.0, 0.5) == 0.0


def test_dmnorm_prec_and_cov_forms_against_scipy():
    """dmnorm(mean, prec = ) / dmnorm(mean, cov = ) with constant matrices (the factor is taken once, like nimble does), n = 12:
    the oracle's whole-model log-probability against scipy.stats.multivariate_normal, and the lowering accepts the model."""
    from scipy.stats import multivariate_normal as mvn
    import models
    from helpers import build_engine, build_oracle
    spec = models.mvn_prior()
    orc = build_oracle(spec)
    c = spec["constants"]
    rng = np.random.default_rng(5)
    for _ in range(3):
        mu = rng.standard_normal(c["n"])
        want = mvn.logpdf(mu, c["m0"], np.linalg.inv(c["Q"])) + mvn.logpdf(spec["data"]["y"], mu, c["S"]).sum()
        got = orc.logprob(mu)[0]
        assert abs(got - want) <= 1e-11 * abs(want)
    src = build_engine(spec, compileOnly=True).lowered_source()
    assert "apt::dmnorm_chol_log<12>" in src
