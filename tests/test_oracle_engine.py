"""CPU: known-answer and invariant tests of the oracle itself (Philox vectors, reference-code invariants,
analytic posteriors), since the reference pins nothing (SURVEY.md §8c "What pins results instead")."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

import models
from helpers import build_oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def test_philox_known_answers(tmp_path):
    """Random123 kat_vectors for philox4x32-10."""
    src = tmp_path / "kat.c"
    src.write_text('#include <stdio.h>\n#include "philox.h"\nint main(){uint32_t c[3][4]={{0,0,0,0},{0xffffffff,0xffffffff,0xffffffff,0xffffffff},'
                   '{0x243f6a88,0x85a308d3,0x13198a2e,0x03707344}};uint32_t k[3][2]={{0,0},{0xffffffff,0xffffffff},{0xa4093822,0x299f31d0}};'
                   'for(int i=0;i<3;i++){uint32_t o[4];apt_philox4x32_10(c[i],k[i],o);printf("%08x %08x %08x %08x\\n",o[0],o[1],o[2],o[3]);}'
                   'printf("%.17g %.17g\\n", apt_u53(0,0), apt_u53(0xffffffffu,0xffffffffu));return 0;}')
    exe = tmp_path / "kat"
    subprocess.check_call(["gcc", "-O1", "-I", os.path.join(HERE, "..", "oracle"), str(src), "-o", str(exe), "-lm"])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    assert out[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert out[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert out[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"
    lo, hi = (float(x) for x in out[3].split())
    assert 0.0 < lo < 1e-15 and 1.0 - 1e-15 < hi < 1.0


def test_ladder_invariants_c1():
    """Temps[1] never changes (APT:494-496), ladder stays sorted, ULT clamp (APT:497-498), trace shapes (APT:386-393)."""
    spec = models.c1_vignette(K=5)
    o = build_oracle(spec)
    o.run(1500, thin=3)
    tt = o.tempTraj()
    assert tt.shape == (500, 5) and o.samples().shape == (500, 2) and o.logProbs().shape == (500,)
    assert np.all(tt[:, 0] == 1.0)
    assert np.all(np.diff(tt, axis=1) > 0)
    assert np.all(tt <= spec["ULT"] + 5)
    # continuation appends (APT:378-385); resetMV restarts the sample store
    o.run(300, reset=False, thin=3)
    assert o.samples().shape == (600, 2) and o.tempTraj().shape == (100, 5)
    o.run(300, reset=False, resetMV=True, thin=3)
    assert o.samples().shape == (100, 2)


def test_fixed_ladder_keeps_temps():
    spec = models.c2_mixture(K=6)
    o = build_oracle(spec)
    o.run(400, adaptTemps=False)
    assert np.array_equal(o.Temps(), np.sort(spec["Temps"]))
    acc = o.accCountSwap()
    assert np.all(np.diag(acc) == 0) and acc.sum() > 0


def test_c2_analytic_posterior():
    """C2's posterior is the exact mixture 1/2 N((-ybar1, ybar2), I/n) + 1/2 N((ybar1, ybar2), I/n) (flat-ish prior)."""
    spec = models.c2_mixture(n=20, K=8, tmax=200.0)
    L = 8
    o = build_oracle(spec, nLadders=L)
    o.run(6000, thin=2, nthreads=8)
    y = spec["data"]["y"]
    n = y.shape[0]
    yb = y.mean(0)
    s = np.concatenate([o.samples(l)[500:] for l in range(L)])
    shrink = 1.0 / (1.0 + 1.0 / (n * 100.0 ** 2))
    # |theta1| ~ N(ybar1, 1/n) folded (modes are far from 0), theta2 ~ N(ybar2, 1/n)
    per_ladder_abs = np.array([np.abs(o.samples(l)[500:, 0]).mean() for l in range(L)])
    per_ladder_t2 = np.array([o.samples(l)[500:, 1].mean() for l in range(L)])
    for est, truth in ((per_ladder_abs, yb[0] * shrink), (per_ladder_t2, yb[1] * shrink)):
        mcse = est.std(ddof=1) / np.sqrt(L)
        assert abs(est.mean() - truth) < 4 * mcse + 1e-3, (est.mean(), truth, mcse)
    assert abs(np.var(s[:, 1]) - 1.0 / n) < 0.15 / n
    frac_pos = np.array([(o.samples(l)[500:, 0] > 0).mean() for l in range(L)])
    assert abs(frac_pos.mean() - 0.5) < 4 * frac_pos.std(ddof=1) / np.sqrt(L) + 0.02  # both modes visited equally


def test_demo_model_slice_sampler_exact_posterior():
    """sampler_slice_tempered (APT:898-995) on the discrete node of the reference's demo model (demo/APT_demo.R:83-92):
    the marginal posterior of k is known exactly (models.demo_wrong_k_posterior)."""
    spec = models.demo_wrong()
    exact = models.demo_wrong_k_posterior(spec)
    L, niter, burn = 32, 10000, 1000
    o = build_oracle(spec, nLadders=L)
    o.run(niter, nthreads=8)
    freq = np.zeros((L, len(exact)))
    for l in range(L):
        k = o.samples(l)[burn:, 0]
        assert np.all(k == np.floor(k)) and k.min() >= 1 and k.max() <= len(exact)  # floor() of the slice draw, inside 1..K
        freq[l] = np.bincount(k.astype(int) - 1, minlength=len(exact)) / len(k)
    mean, mcse = freq.mean(axis=0), freq.std(axis=0, ddof=1) / np.sqrt(L)
    assert np.all(np.abs(mean - exact) <= 4 * mcse + 2e-3), (mean, exact, mcse)
    widths = np.array([o.scale(0, t, 2) for t in range(o.K)])
    assert np.all(widths > 0) and not np.allclose(widths, 1.0)  # adaptiveProcedure ran (APT:976-987)
    bad = dict(spec)
    bad["samplers"] = [dict(target=["theta0", "sigma0"], type="sampler_slice_tempered", control={})]
    with pytest.raises(ValueError, match="more than one target"):  # APT:928
        build_oracle(bad)


def test_multinomial_sampler_exact_posterior_and_reference_quirk():
    """sampler_RW_multinomial_tempered (APT:1005-1145) on a latent dmulti node whose posterior is enumerated exactly
    (models.multinomial_latent_posterior).  With the exact reverse proposal (test-only switch) the restatement
    reproduces the posterior; restated as the reference has it (APT:1093-1095: pRevSwap computed from x_to + 2 nSwap)
    the stationary law is measurably off, which is what parity with the reference means for this sampler."""
    from oracle.oracle import lib
    spec = models.multinomial_latent()
    ex, ea = models.multinomial_latent_posterior(spec)
    exact = np.append(ex, ea)
    L, niter, burn = 32, 10000, 1000
    z = {}
    for mode in (1, 0):
        lib().orc_set_multinomial_exact_reverse(mode)
        try:
            o = build_oracle(spec, nLadders=L)
            o.run(niter, nthreads=8)
        finally:
            lib().orc_set_multinomial_exact_reverse(0)
        means = np.zeros((L, 5))
        for l in range(L):
            S = o.samples(l)[burn:]
            x = S[:, :4]
            assert np.all(x == np.floor(x)) and x.min() >= 0 and np.all(x.sum(axis=1) == 30)  # moves keep sum(x) = Ntotal
            means[l] = S.mean(axis=0)
        mean, mcse = means.mean(axis=0), means.std(axis=0, ddof=1) / np.sqrt(L)
        z[mode] = (mean - exact) / mcse
        st = o.multinomialState(0, 0, 0, 4)
        u, EN = st[0], st[1 + 5 * 16:1 + 6 * 16].reshape(4, 4, order="F")
        assert 0 <= u <= np.pi and np.allclose(EN, EN.T) and EN.min() >= 1 and EN.max() <= 30 and EN.max() > 1  # APT:1108-1125
    assert np.all(np.abs(z[1]) <= 4), z[1]
    assert np.abs(z[0]).max() > 6, z[0]  # the reference's proposal ratio leaves a bias far outside Monte-Carlo error


def test_multinomial_sampler_setup_checks():
    spec = models.multinomial_latent()
    bad = dict(spec)
    bad["samplers"] = [dict(target="x[1:4]", type="sampler_RW_multinomial_tempered", control={})]
    with pytest.raises(ValueError, match="useTempering unspecified"):  # APT:1015
        build_oracle(bad)
    bad["samplers"] = [dict(target="a", type="sampler_RW_multinomial_tempered", control=dict(useTempering=True))]
    with pytest.raises(ValueError, match="multinomial distributions"):  # APT:1046
        build_oracle(bad)
    bad["samplers"] = [dict(target="x[1:4]", type="sampler_RW_multinomial_tempered", control=dict(useTempering=True, adaptInterval=50))]
    with pytest.raises(ValueError, match="adaptInterval < 100"):  # APT:1048
        build_oracle(bad)


def test_sampler_validation_errors():
    spec = models.c2_mixture(K=4)
    bad = dict(spec)
    bad["samplers"] = [dict(target="theta[1]", type="sampler_RW_tempered", control={})]
    with pytest.raises(ValueError, match="temperPriors unspecified"):  # APT:660
        build_oracle(bad)
    bad["samplers"] = [dict(target="theta[1:2]", type="sampler_RW_tempered", control=dict(temperPriors=True))]
    with pytest.raises(ValueError, match="more than one target"):  # APT:682
        build_oracle(bad)
