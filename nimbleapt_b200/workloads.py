"""Workload definitions (SURVEY.md §8d / BASELINE.md §4): BUGS text + synthetic inputs for configs C1..C5, and
`build_engine`, which builds one of them through the package's public interface (nimbleModel -> configureMCMC ->
addSampler -> buildAPT, the call sequence of nimbleAPT-vignette.Rmd:127-137).  Shared by bench.py, __graft_entry__ and
the tests; product-side only (numpy + this package, never the CPU oracle).  Data seed 20261019 as proposed by the survey."""
import numpy as np

DATA_SEED = 20261019


def c1_vignette(nObs=100, K=8):
    """nimbleAPT/vignettes/nimbleAPT-vignette.Rmd:56-72,129-135 verbatim model; Temps=1:K, ULT=1000."""
    rng = np.random.default_rng(DATA_SEED)
    y = rng.standard_normal((nObs, 2)) + 3.0  # simulated from centroids=(-3,-3): abs() -> mean (3,3), cov I
    code = """{
    for (ii in 1:nObs) {
        y[ii,1:2] ~ dmnorm(mean=absCentroids[1:2], cholesky=cholCov[1:2,1:2], prec_param=0)
    }
    absCentroids[1:2] <- abs(centroids[1:2])
    for (ii in 1:2) {
        centroids[ii] ~ dnorm(0, sd=1E3)
    }
    }"""
    return dict(code=code, constants=dict(nObs=nObs, cholCov=np.eye(2)), data=dict(y=y),
                inits=dict(centroids=np.array([-3.0, -3.0])),
                samplers=[dict(target="centroids[1]", type="sampler_RW_tempered", control=dict(temperPriors=True)),
                          dict(target="centroids[2]", type="sampler_RW_tempered", control=dict(temperPriors=True))],
                monitors=["centroids"], Temps=np.arange(1, K + 1, dtype=float), ULT=1000.0)


def mvn_prior(n=12, nObs=15, K=6):
    """A 12-dimensional parameter with a dmnorm(mean, prec = constant matrix) prior, observed through dmnorm(mean, cov = ) rows
    (the parameterisations other than the vignette's cholesky = ; nimble takes chol() of a constant matrix once)."""
    rng = np.random.default_rng(DATA_SEED + 41)
    A = rng.standard_normal((n, n)) * 0.3
    prec = A @ A.T + np.eye(n) * 0.5
    B = rng.standard_normal((n, n)) * 0.2
    cov = B @ B.T + np.eye(n)
    truth = rng.standard_normal(n)
    y = truth + rng.multivariate_normal(np.zeros(n), cov, size=nObs)
    code = """{
    mu[1:n] ~ dmnorm(mean = m0[1:n], prec = Q[1:n, 1:n])
    for (i in 1:nObs) {
        y[i, 1:n] ~ dmnorm(mu[1:n], cov = S[1:n, 1:n])
    }
    }"""
    return dict(code=code, constants=dict(n=n, nObs=nObs, m0=np.zeros(n), Q=prec, S=cov), data=dict(y=y),
                inits=dict(mu=np.zeros(n)),
                samplers=[dict(target="mu[1:%d]" % n, type="sampler_RW_block_tempered", control=dict(temperPriors=True, scale=0.3)),
                          dict(target="mu[1]", type="sampler_RW_tempered", control=dict(temperPriors=True))],
                monitors=["mu"], Temps=np.exp(np.linspace(0.0, np.log(30.0), K)), ULT=1e6)


def c2_mixture(n=20, K=32, tmax=1000.0):
    """2-D bimodal posterior through an abs() link; exact posterior is a 2-component Gaussian mixture."""
    rng = np.random.default_rng(DATA_SEED + 2)
    y = np.stack([rng.standard_normal(n) + 3.0, rng.standard_normal(n)], axis=1)
    code = """{
    for (i in 1:n) {
        y[i,1] ~ dnorm(abs(theta[1]), sd=1)
        y[i,2] ~ dnorm(theta[2], sd=1)
    }
    for (j in 1:2) {
        theta[j] ~ dnorm(0, sd=100)
    }
    }"""
    temps = np.exp(np.linspace(0.0, np.log(tmax), K))
    return dict(code=code, constants=dict(n=n), data=dict(y=y), inits=dict(theta=np.array([3.0, 0.0])),
                samplers=[dict(target="theta[1]", type="sampler_RW_tempered", control=dict(temperPriors=True)),
                          dict(target="theta[2]", type="sampler_RW_tempered", control=dict(temperPriors=True))],
                monitors=["theta"], Temps=temps, ULT=1e6)


def logistic(N, p, K, tmax=50.0, seed_off=3, propcov=True):
    """C3 (N=1e6,p=50,K=16) / C5 (N=1024,p=8,K=64): Bayesian logistic regression, one block sampler."""
    rng = np.random.default_rng(DATA_SEED + seed_off)
    X = rng.standard_normal((N, p)); X[:, 0] = 1.0
    beta_true = rng.standard_normal(p) * 0.25
    eta = X @ beta_true
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-eta))).astype(float)
    code = """{
    for (i in 1:N) {
        y[i] ~ dbinom(prob=p[i], size=1)
        logit(p[i]) <- inprod(beta[1:P], X[i,1:P])
    }
    for (j in 1:P) {
        beta[j] ~ dnorm(0, sd=10)
    }
    }"""
    ctl = dict(temperPriors=True)
    if propcov:
        pr = 1.0 / (1.0 + np.exp(-eta))
        W = pr * (1 - pr)
        H = (X * W[:, None]).T @ X + np.eye(p) / 100.0
        cov = np.linalg.inv(H) * (2.38 ** 2 / p)
        ctl["propCov"] = (cov + cov.T) / 2
    temps = np.exp(np.linspace(0.0, np.log(tmax), K))
    return dict(code=code, constants=dict(N=N, P=p, X=X), data=dict(y=y), inits=dict(beta=np.zeros(p)),
                samplers=[dict(target="beta[1:%d]" % p, type="sampler_RW_block_tempered", control=ctl)],
                monitors=["beta"], Temps=temps, ULT=1e6)


def c3_logistic(N=1000000, p=50, K=16):
    return logistic(N, p, K, seed_off=3)


def c5_small_logistic(N=1024, p=8, K=64):
    return logistic(N, p, K, seed_off=5)


def c4_glmm(G=1000, per=100, K=24, tmax=30.0):
    """Hierarchical Poisson GLMM, mixed scalar / block / log-scale samplers."""
    rng = np.random.default_rng(DATA_SEED + 4)
    N = G * per
    g = np.repeat(np.arange(1, G + 1), per).astype(float)
    x = rng.standard_normal(N)
    u_true = rng.standard_normal(G) * 0.5
    lam = np.exp(0.5 + 0.3 * x + u_true[g.astype(int) - 1])
    y = rng.poisson(lam).astype(float)
    code = """{
    for (i in 1:N) {
        y[i] ~ dpois(lam[i])
        log(lam[i]) <- b[1] + b[2]*x[i] + u[g[i]]
    }
    for (k in 1:G) {
        u[k] ~ dnorm(0, sd=sigma)
    }
    sigma ~ dunif(0, 5)
    for (j in 1:2) {
        b[j] ~ dnorm(0, sd=10)
    }
    }"""
    samplers = [dict(target="u[%d]" % (k + 1), type="sampler_RW_tempered", control=dict(temperPriors=True))
                for k in range(G)]
    samplers.append(dict(target="b[1:2]", type="sampler_RW_block_tempered", control=dict(temperPriors=True)))
    samplers.append(dict(target="sigma", type="sampler_RW_tempered", control=dict(temperPriors=True, log=True)))
    temps = np.exp(np.linspace(0.0, np.log(tmax), K))
    return dict(code=code, constants=dict(N=N, G=G, x=x, g=g), data=dict(y=y),
                inits=dict(u=np.zeros(G), sigma=np.array(1.0), b=np.zeros(2)), samplers=samplers,
                monitors=["b", "sigma"], Temps=temps, ULT=1e6)


def mixed_dists(n=40, K=6):
    """Every distribution of the supported subset in one small model (gamma / beta / binomial(size>1) / exponential /
    uniform / Poisson / normal with precision parameterisation), scalar + block + log-scale + reflective samplers."""
    rng = np.random.default_rng(DATA_SEED + 9)
    size = rng.integers(3, 12, n).astype(float)
    yb = rng.binomial(size.astype(int), 0.35).astype(float)
    yp = rng.poisson(2.5, n).astype(float)
    yg = rng.gamma(3.0, 1.0 / 1.5, n)
    yn = rng.normal(0.7, 0.8, n)
    code = """{
    for (i in 1:n) {
        yb[i] ~ dbinom(prob=p, size=size[i])
        yp[i] ~ dpois(rate * expo[i])
        yg[i] ~ dgamma(shape, rate=grate)
        yn[i] ~ dnorm(mu, tau)
    }
    p ~ dbeta(2, 3)
    rate ~ dgamma(shape=2, rate=0.5)
    shape ~ dexp(0.2)
    grate ~ dunif(0.05, 20)
    mu ~ dnorm(0, 0.01)
    tau ~ dgamma(1, scale=2)
    }"""
    samplers = [dict(target="p", type="sampler_RW_tempered", control=dict(temperPriors=True, reflective=True, scale=0.1)),
                dict(target="rate", type="sampler_RW_tempered", control=dict(temperPriors=True, log=True, scale=0.3)),
                dict(target=["shape", "grate"], type="sampler_RW_block_tempered", control=dict(temperPriors=True, scale=0.3, adaptInterval=50)),
                dict(target="mu", type="sampler_RW_tempered", control=dict(temperPriors=False, scale=0.3)),
                dict(target="tau", type="sampler_RW_tempered", control=dict(temperPriors=True, log=True, scale=0.3))]
    temps = np.exp(np.linspace(0.0, np.log(20.0), K))
    return dict(code=code, constants=dict(n=n, size=size, expo=rng.random(n) + 0.5),
                data=dict(yb=yb, yp=yp, yg=yg, yn=yn),
                inits=dict(p=np.array(0.4), rate=np.array(2.0), shape=np.array(2.0), grate=np.array(1.0), mu=np.array(0.0), tau=np.array(1.0)),
                samplers=samplers, monitors=["p", "rate", "shape", "grate", "mu", "tau"], Temps=temps, ULT=1e6)


def demo_wrong(N=50, K=10, nTemps=7):
    """The 'wrong' model of nimbleAPT/demo/APT_demo.R:83-92 with its four tempered samplers (demo:173-177) and
    Temps = 1:7 (demo:196).  The demo's user functions k2theta / k2sigma (demo:42-71: theta0 + rfx_theta[k] and
    abs(sigma0 + rfx_sigma[k])) are written as index expressions on the sampled node k, which the BUGS subset
    supports; `pk`, `rfx_theta`, `rfx_sigma` are passed as constants (the demo passes them as inits of
    right-hand-side-only nodes, demo:108)."""
    rng = np.random.default_rng(DATA_SEED + 7)
    y = rng.standard_normal(N)  # right$simulate('y'): theta = 0, sigma0 = 1 (demo:99-101,125)
    rfx_theta, rfx_sigma = rng.standard_normal(K), rng.standard_normal(K)
    code = """{
    k      ~ dcat(pk[1:K])
    theta0 ~ dnorm(0, sd=1E6)
    theta  <- theta0 + rfx_theta[k]
    sigma0 ~ dexp(1)
    sigma  <- abs(abs(sigma0 + rfx_sigma[k]))
    for (i in 1:N) {
        y[i] ~ dnorm(mean=theta, sd=sigma)
    }
    }"""
    tp = dict(temperPriors=True)
    return dict(code=code, constants=dict(N=N, K=K, pk=np.full(K, 1.0 / K), rfx_theta=rfx_theta, rfx_sigma=rfx_sigma),
                data=dict(y=y), inits=dict(k=np.array(1.0), theta0=np.array(0.0), sigma0=np.array(1.0)),
                samplers=[dict(target="theta0", type="sampler_RW_tempered", control=tp),
                          dict(target="sigma0", type="sampler_RW_tempered", control=tp),
                          dict(target="k", type="sampler_slice_tempered", control=tp),
                          dict(target=["theta0", "sigma0"], type="sampler_RW_block_tempered", control=tp)],
                monitors=["k", "theta0", "sigma0"],
                monitors2=["logProb_k", "logProb_theta0", "logProb_sigma0", "logProb_y"],  # addMonitors2(allLogProbs), demo:180-181
                Temps=np.arange(1, nTemps + 1, dtype=float), ULT=1e6)


# the demo's two user-supplied functions and its 'wrong' model, verbatim (demo/APT_demo.R:42-71, 83-92)
K2THETA_R = """function(k = double(),              ## A categorical random variable in continuous form.
                   K = double(),              ## Upper limit of k.
                   randomEffects = double(1), ## A set of "randomEffects" which will be fixed during MCMC in order to generate poor mixing
                   theta0=double()            ## An intercept term
                   ) {
        ## Uses a categorical rv to apply a shift in the value of a parameter theta
        k <- ceiling(k)
        if (k < 1 | k > K)
            return(-Inf)
        theta <- theta0 + randomEffects[k]
        return(theta)
        returnType(double())
    }"""
K2SIGMA_R = """function(k = double(), K = double(), randomEffects = double(1), sigma0=double()) {
        ## Uses a categorical rv to determine a change in a standard deviation parameter theta
        k <- ceiling(k)
        if (k < 1 | k > K)
            return(-Inf)
        sigma <- abs(sigma0 + randomEffects[k])
        return(sigma)
        returnType(double())
    }"""
WRONG_CODE_R = """{
    k      ~ dcat(pk[1:K])
    theta0 ~ dnorm(0, sd=1E6)
    theta <- k2theta(k, K, rfx_theta[1:K], theta0)
    sigma0 ~ dexp(1)
    sigma <- abs(k2sigma(k, K, rfx_sigma[1:K], sigma0))
    for (i in 1:N)
        y[i] ~ dnorm(mean=theta, sd=sigma)
}"""


def demo_wrong_verbatim(N=50, K=10, nTemps=7):
    """demo_wrong() with the demo's own text: `wrongCode` and the nimbleFunctions k2theta / k2sigma as the demo has them."""
    spec = demo_wrong(N, K, nTemps)
    spec["code"] = WRONG_CODE_R
    spec["functions"] = dict(k2theta=K2THETA_R, k2sigma=K2SIGMA_R)
    return spec


def demo_wrong_k_posterior(spec):
    """Exact marginal posterior of k in demo_wrong(): theta0 has an (effectively) flat prior and integrates out
    analytically, sigma0 ~ dexp(1) is integrated on a grid."""
    y = spec["data"]["y"]
    N, rs = len(y), spec["constants"]["rfx_sigma"]
    s2 = float(np.sum((y - y.mean()) ** 2))
    g = np.linspace(1e-6, 30.0, 3000001)
    lw = []
    for kk in range(len(rs)):
        sg = np.abs(g + rs[kk])
        with np.errstate(divide="ignore", invalid="ignore"):
            lf = -g - (N - 1) * np.log(sg) - s2 / (2 * sg ** 2)
        lw.append(np.where(np.isfinite(lf), lf, -np.inf))
    lw = np.array(lw)
    pk = np.trapezoid(np.exp(lw - lw.max()), g, axis=1)
    return pk / pk.sum()


def multinomial_latent(Lc=4, N=30, K=5, adaptInterval=100):
    """Latent counts x[1:Lc] ~ dmulti(prob, N) observed through Poisson counts, one sampler_RW_multinomial_tempered
    (APT_functions.R:1005-1145) plus a scalar sampler on the rate multiplier.  Small enough for the exact posterior of
    x (all compositions of N into Lc parts) to be enumerated."""
    rng = np.random.default_rng(DATA_SEED + 31)
    prob = np.array([0.1, 0.2, 0.3, 0.4])[:Lc]
    prob = prob / prob.sum()
    y = np.array([2.0, 9.0, 6.0, 14.0])[:Lc]
    code = """{
    x[1:Lc] ~ dmulti(prob = prob[1:Lc], size = N)
    for (j in 1:Lc) {
        y[j] ~ dpois(0.5 + a * x[j])
    }
    a ~ dgamma(shape = 4, rate = 4)
    }"""
    x0 = np.array([N - 3.0 * (Lc - 1)] + [3.0] * (Lc - 1))
    return dict(code=code, constants=dict(Lc=Lc, N=N, prob=prob), data=dict(y=y),
                inits=dict(x=x0, a=np.array(1.0)),
                samplers=[dict(target="x[1:%d]" % Lc, type="sampler_RW_multinomial_tempered",
                               control=dict(useTempering=True, adaptInterval=adaptInterval)),
                          dict(target="a", type="sampler_RW_tempered", control=dict(temperPriors=True, log=True, scale=0.3))],
                monitors=["x", "a"], Temps=np.exp(np.linspace(0.0, np.log(8.0), K)), ULT=1e6)


def multinomial_latent_posterior(spec):
    """Exact posterior mean of x (and of a) in multinomial_latent(): every composition of N into Lc parts is enumerated,
    a ~ dgamma(4, 4) is integrated on a grid."""
    import itertools
    from scipy import special
    c, y = spec["constants"], spec["data"]["y"]
    Lc, N, prob = int(c["Lc"]), int(c["N"]), np.asarray(c["prob"], dtype=float)
    comps = np.array([np.diff(np.concatenate([[-1], cuts, [N + Lc - 1]])) - 1
                      for cuts in itertools.combinations(range(N + Lc - 1), Lc - 1)], dtype=float)
    lprior = special.gammaln(N + 1) + np.sum(comps * np.log(prob / prob.sum()) - special.gammaln(comps + 1), axis=1)
    a = np.linspace(1e-4, 12.0, 4001)
    lpa = 3 * np.log(a) - 4 * a                                      # dgamma(shape = 4, rate = 4) up to a constant
    lam = 0.5 + a[None, :, None] * comps[:, None, :]                 # [composition, a, j]
    ll = np.sum(y * np.log(lam) - lam, axis=2) + lpa[None, :]        # dpois up to a constant
    m = ll.max()
    wa = np.exp(ll - m)                                              # [composition, a]
    w = np.exp(lprior - lprior.max())[:, None] * wa
    Z = np.trapezoid(w.sum(axis=0), a)
    ex = np.array([np.trapezoid((w * comps[:, j][:, None]).sum(axis=0), a) for j in range(Lc)]) / Z
    ea = np.trapezoid(w.sum(axis=0) * a, a) / Z
    return ex, ea


def build_engine(spec, nLadders=1, **kw):
    """A workload spec (one of the functions above) through the drop-in interface; returns the built APT object."""
    from . import api as nb
    m = nb.nimbleModel(nb.nimbleCode(spec["code"]), constants=spec["constants"], data=spec["data"], inits=spec["inits"],
                       functions=spec.get("functions"))
    conf = nb.configureMCMC(m, monitors=spec["monitors"], monitors2=spec.get("monitors2"))
    conf.removeSamplers()
    for s in spec["samplers"]:
        conf.addSampler(s["target"], type=s["type"], control=s["control"])
    return nb.buildAPT(conf, Temps=spec["Temps"], ULT=spec["ULT"], nLadders=nLadders, **kw)


def initial_values(spec):
    """The flat parameter vector of spec["inits"] in declaration order (what apt_set_inits takes)."""
    return np.concatenate([np.ravel(spec["inits"][k], order="F") for k in spec["inits"]])
