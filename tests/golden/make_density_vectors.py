"""Generates tests/golden/density_vectors.json: log-densities of the supported BUGS distributions at fixed arguments,
evaluated with mpmath at 60 digits from their textbook definitions (no Rmath, no scipy, no code of this repository).
The reference ships no golden vectors for this path (parity unpinned); these are the committed,
independent known answers the oracle (and through it the CUDA engine) is pinned to.   python tests/golden/make_density_vectors.py"""
import json
import os
import mpmath as mp
import numpy as np

mp.mp.dps = 60
rng = np.random.default_rng(20261019)


def f(x):
    return float(x)


def dnorm(x, mu, sd):
    return -mp.log(sd) - mp.mpf("0.5") * mp.log(2 * mp.pi) - ((mp.mpf(x) - mu) / sd) ** 2 / 2


def dgamma(x, shape, scale):  # R's dgamma(x, shape, scale)
    x, shape, scale = mp.mpf(x), mp.mpf(shape), mp.mpf(scale)
    return (shape - 1) * mp.log(x) - x / scale - mp.loggamma(shape) - shape * mp.log(scale)


def dbeta(x, a, b):
    x, a, b = mp.mpf(x), mp.mpf(a), mp.mpf(b)
    return (a - 1) * mp.log(x) + (b - 1) * mp.log1p(-x) - (mp.loggamma(a) + mp.loggamma(b) - mp.loggamma(a + b))


def dbinom(x, p, n):  # canonical BUGS order (prob, size)
    x, p, n = mp.mpf(x), mp.mpf(p), mp.mpf(n)
    return mp.loggamma(n + 1) - mp.loggamma(x + 1) - mp.loggamma(n - x + 1) + x * mp.log(p) + (n - x) * mp.log1p(-p)


def dpois(x, lam):
    x, lam = mp.mpf(x), mp.mpf(lam)
    return x * mp.log(lam) - lam - mp.loggamma(x + 1)


def dexp(x, scale):  # R's dexp(x, 1/scale) = nimble's canonical scale form used by the lowering
    return -mp.mpf(x) / scale - mp.log(scale)


def dunif(x, a, b):
    return -mp.log(mp.mpf(b) - a)


out = {"dnorm": [], "dgamma": [], "dbeta": [], "dbinom": [], "dpois": [], "dexp": [], "dunif": []}
for _ in range(150):
    x, mu, sd = rng.normal(0, 50), rng.normal(0, 50), float(np.exp(rng.normal(0, 3)))
    out["dnorm"].append([x, mu, sd, f(dnorm(x, mu, sd))])
for _ in range(150):
    shape, scale = float(np.exp(rng.normal(0, 2.5))), float(np.exp(rng.normal(0, 2)))
    x = float(rng.gamma(shape, scale)) if rng.random() < 0.7 else float(np.exp(rng.normal(0, 4)))
    if x > 1e-280 and np.isfinite(x):
        out["dgamma"].append([x, shape, scale, f(dgamma(x, shape, scale))])
for _ in range(150):
    a, b = float(np.exp(rng.normal(0, 2))), float(np.exp(rng.normal(0, 2)))
    x = float(rng.beta(a, b))
    if 1e-300 < x < 1 - 1e-12:
        out["dbeta"].append([x, a, b, f(dbeta(x, a, b))])
for _ in range(150):
    n = int(rng.choice([1, 2, 5, 17, 100, 1000, 100000]))
    p = float(rng.uniform(1e-6, 1 - 1e-6)) if rng.random() < 0.8 else float(10 ** rng.uniform(-12, -1))
    x = int(min(n, rng.binomial(n, p) + rng.integers(0, 3)))
    out["dbinom"].append([x, p, n, f(dbinom(x, p, n))])
for _ in range(150):
    lam = float(np.exp(rng.uniform(-8, 16)))
    x = int(rng.poisson(min(lam, 1e15))) if rng.random() < 0.8 else int(rng.integers(0, 200))
    out["dpois"].append([x, lam, f(dpois(x, lam))])
for _ in range(60):
    scale, x = float(np.exp(rng.normal(0, 3))), float(np.exp(rng.normal(0, 3)))
    out["dexp"].append([x, scale, f(dexp(x, scale))])
for _ in range(30):
    a = float(rng.normal(0, 10)); b = a + float(np.exp(rng.normal(0, 2))); x = float(rng.uniform(a, b))
    out["dunif"].append([x, a, b, f(dunif(x, a, b))])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "density_vectors.json")
with open(path, "w") as fh:
    json.dump(out, fh)
print({k: len(v) for k, v in out.items()}, "->", path)
