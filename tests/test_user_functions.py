"""User-supplied deterministic functions (nimbleFunction with a `run` function, demo/APT_demo.R:42-71): the demo's `wrongCode`
and its k2theta / k2sigma are taken verbatim and inlined by the product's lowering (C++) and, independently, by the oracle's
front end (Python).  CPU part: the two inliners agree with the hand-rewritten model on the log-probability, the lowered
source contains what the function bodies say, unsupported constructs are named.  The GPU parity of the verbatim model is
tests/test_gpu_parity.py[demo]."""
import ctypes as C

import numpy as np
import pytest

import models
import nimbleapt_b200 as nb
from helpers import build_engine, build_oracle
from nimbleapt_b200._lib import check, lib


def test_oracle_inlines_the_demo_functions_like_the_rewritten_model():
    a, b = build_oracle(models.demo_wrong()), build_oracle(models.demo_wrong_verbatim())
    for th in ([3.0, 0.5, 1.2], [10.0, -0.3, 0.4], [1.0, 2.0, 0.1], [11.0, 0.1, 1.0], [0.0, 0.1, 1.0], [7.0, -1.5, 3.0]):
        la, lb = a.logprob(np.array(th))[0], b.logprob(np.array(th))[0]
        assert la == lb or (np.isinf(la) and np.isinf(lb)), (th, la, lb)


def test_lowering_inlines_the_demo_functions():
    apt = build_engine(models.demo_wrong_verbatim(), compileOnly=True)
    src = apt.lowered_source()
    # k <- ceiling(k); if (k < 1 | k > K) return(-Inf); theta0 + randomEffects[k]
    assert "apt::dyn_index(ceil(th[0 + 0]), 10)" in src
    assert "((ceil(th[0 + 0]) < 1.0) || (ceil(th[0 + 0]) > 10.0)) ? (-CUDART_INF)" in src
    assert "apt::slice_update<Model, S2>" in src


FN_BRANCHES = """nimbleFunction(run = function(x = double(), w = double(1), n = double()) {
    s <- 0
    if (x > 0 & !(x >= 2)) {
        s <- w[1] * x
    } else {
        s <- w[2] - x
    }
    if (n == 3) return(s + w[n])
    else return(floor(s))
    returnType(double())
})"""


def _oracle_value(x):
    code = "{ a ~ dnorm(0, sd = 10)\n mu <- f(a, w[1:3], n)\n for (i in 1:2) y[i] ~ dnorm(mu, sd = 1) }"
    spec = dict(code=code, constants=dict(w=np.array([2.0, 5.0, 0.25]), n=3.0), data=dict(y=np.array([0.0, 0.0])),
                inits=dict(a=np.array(x)), samplers=[dict(target="a", type="sampler_RW_tempered", control=dict(temperPriors=True))],
                monitors=["a"], Temps=np.array([1.0, 2.0]), ULT=1e6, functions=dict(f=FN_BRANCHES))
    return spec, build_oracle(spec)


@pytest.mark.parametrize("x,mu", [(1.0, 2.0 * 1.0 + 0.25), (3.0, 5.0 - 3.0 + 0.25), (-1.5, 5.0 + 1.5 + 0.25)])
def test_if_else_assignments_and_logical_operators(x, mu):
    """Branches that assign (merged as a conditional), `&`, `!`, `==`, an early return inside if / else."""
    from scipy.stats import norm
    spec, orc = _oracle_value(x)
    want = norm.logpdf(x, 0, 10) + 2 * norm.logpdf(0.0, mu, 1)
    assert abs(orc.logprob(np.array([x]))[0] - want) < 1e-12
    src = build_engine(spec, compileOnly=True).lowered_source()
    assert "&&" in src and "? " in src  # the merged assignment is a lazily evaluated conditional in the generated code


def test_unsupported_function_bodies_are_named():
    h = C.c_void_p()
    check(lib().apt_model_create(b"{ a ~ dnorm(0, 1) }", C.byref(h)))
    try:
        for src, msg in (("function(x = double()) { for (i in 1:3) x <- x + 1; return(x) }", "unsupported statement|expected"),
                         ("function(x = double(2)) { return(x) }", "matrix parameters"),
                         ("function(x = double(1)) { x[1] <- 2; return(x[1]) }", "assignment to an element")):
            with pytest.raises(nb.AptError, match=msg):
                check(lib().apt_model_add_function(h, b"f", src.encode()))
    finally:
        lib().apt_model_free(h)
    with pytest.raises(nb.AptError, match="does not return a value|must return a value"):
        m = nb.nimbleModel(nb.nimbleCode("{ a ~ dnorm(0, 1)\n b <- g(a)\n y ~ dnorm(b, 1) }"), data=dict(y=0.0), inits=dict(a=0.0),
                           functions=dict(g="function(x = double()) { if (x > 0) return(x) }"))
        conf = nb.configureMCMC(m, monitors=["a"])
        conf.removeSamplers()
        conf.addSampler("a", type="sampler_RW_tempered", control=dict(temperPriors=True))
        nb.buildAPT(conf, Temps=[1.0, 2.0], compileOnly=True)
