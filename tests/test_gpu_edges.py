"""-m gpu: edge cases of the run prologue and of the ladder, engine against oracle on identical inputs:
the smallest ladder (2 rungs), empty runs and runs shorter than the thinning interval (APT:386-393), the ULT clamp
firing on most rungs (APT:497-498), an unsorted Temps argument (APT:266), niter < 0 (APT:350)."""
import numpy as np
import pytest

import models
import nimbleapt_b200 as nb
from helpers import build_engine, build_oracle, run_pair

pytestmark = pytest.mark.gpu


def _same(eng, orc, L, niter):
    for l in range(L):
        np.testing.assert_allclose(eng.mvSamples.matrix(l), orc.samples(l), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(eng.TempsLadder(l), orc.Temps(l), rtol=1e-9)
        np.testing.assert_allclose(eng.tempTrajLadder(l), orc.tempTraj(l), rtol=1e-9)
        np.testing.assert_array_equal(eng.accCountSwapLadder(l), orc.accCountSwap(l))


def test_two_rung_ladder():
    spec = dict(models.c2_mixture(n=20, K=8, tmax=200.0))
    spec["Temps"] = np.array([1.0, 7.0])
    eng, orc, recs = run_pair(spec, 500, L=3, inject=False)
    for l in range(3):
        assert np.array_equal(eng.decisions(l, 500), recs[l][0])
        assert np.array_equal(eng.swapRecord(l, 500) >> 16, recs[l][1] >> 16)
    _same(eng, orc, 3, 500)
    assert eng.TempsLadder(0)[0] == 1.0 and eng.TempsLadder(0)[1] != 7.0  # rung 1 fixed (APT:494-496), rung 2 adapted


def test_empty_and_short_runs():
    spec = models.c2_mixture(n=20, K=8, tmax=200.0)
    eng = build_engine(spec, nLadders=2, record=True)
    orc = build_oracle(spec, nLadders=2)
    eng.run(0)
    orc.run(0)
    assert eng.mvSamples.matrix(0).shape == (0, 2) and eng.mvSamples.matrix(-1).shape == (2, 0, 2)
    assert eng.tempTrajLadder(0).shape == (0, 8) and eng.logProbsLadder(0).shape == (0,)
    eng.run(7, thin=10)  # shorter than the thinning interval: no row is written, the chains still advance
    orc.run(7, thin=10)
    assert eng.mvSamples.matrix(1).shape == orc.samples(1).shape == (0, 2)
    eng.run(13, reset=False, thin=10)  # continues; one row (iteration 10 of THIS run, APT:514)
    orc.run(13, reset=False, thin=10)
    assert eng.mvSamples.matrix(1).shape == (1, 2)
    _same(eng, orc, 2, 13)
    for l in range(2):
        np.testing.assert_allclose(eng.rungStates(l), np.stack([orc.rung_state(l, k) for k in range(8)]), rtol=1e-8, atol=1e-10)
    with pytest.raises(nb.AptError, match="cannot specify niter < 0"):  # APT:350
        eng.run(-1)
    with pytest.raises(nb.AptError, match="thin and thin2 must be >= 1"):
        eng.run(5, thin=0)


def test_ult_clamp_and_unsorted_temps():
    spec = dict(models.c2_mixture(n=20, K=8, tmax=200.0))
    spec["ULT"] = 4.0                                  # most of the ladder sits above ULT: T[ii+1] := ULT + ii (APT:497-498)
    spec["Temps"] = np.array(spec["Temps"])[::-1].copy()  # buildAPT sorts (APT:266)
    eng, orc, recs = run_pair(spec, 400, L=2, inject=False)
    _same(eng, orc, 2, 400)
    T = eng.TempsLadder(0)
    assert np.all(np.diff(T) > 0) and T[0] == 1.0
    assert T.max() <= 4.0 + 8 and np.any(T == 4.0 + np.arange(1, 8)[-1])  # the clamp fired on the top rung
    for l in range(2):
        assert np.array_equal(eng.decisions(l, 400), recs[l][0])
