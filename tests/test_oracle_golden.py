"""CPU: the oracle's log-densities against the committed known answers of tests/golden/density_vectors.json (mpmath,
60 digits, textbook definitions; generator: tests/golden/make_density_vectors.py).  The reference has no golden vectors
for this path, so these fixtures are what pins the oracle; the CUDA engine is pinned to the oracle."""
import json
import os

import numpy as np

from oracle.oracle import lib

HERE = os.path.dirname(os.path.abspath(__file__))


def _check(got, ref, rtol, atol, what):
    assert abs(got - ref) <= atol + rtol * abs(ref), (what, got, ref)


def test_density_vectors():
    L = lib()
    V = json.load(open(os.path.join(HERE, "golden", "density_vectors.json")))
    assert sum(len(v) for v in V.values()) > 800
    for x, mu, sd, ref in V["dnorm"]:
        _check(L.apt_dnorm_log(x, mu, sd), ref, 1e-13, 1e-13, ("dnorm", x, mu, sd))
    for x, shape, scale, ref in V["dgamma"]:
        _check(L.apt_dgamma_log(x, shape, scale), ref, 2e-12, 1e-13, ("dgamma", x, shape, scale))
    for x, a, b, ref in V["dbeta"]:
        _check(L.apt_dbeta_log(x, a, b), ref, 5e-12, 5e-12, ("dbeta", x, a, b))
    for x, p, n, ref in V["dbinom"]:
        _check(L.apt_dbinom_log(float(x), float(n), p), ref, 2e-12, 1e-12, ("dbinom", x, p, n))
    for x, lam, ref in V["dpois"]:
        _check(L.apt_dpois_log(float(x), lam), ref, 2e-12, 1e-12, ("dpois", x, lam))
    for x, scale, ref in V["dexp"]:
        _check(L.apt_dexp_log(x, scale), ref, 1e-13, 1e-13, ("dexp", x, scale))
    for x, a, b, ref in V["dunif"]:
        _check(L.apt_dunif_log(x, a, b), ref, 1e-14, 0.0, ("dunif", x, a, b))
