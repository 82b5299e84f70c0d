import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, models
from nimbleapt_b200.workloads import build_engine
N = 1000000
spec = models.logistic(N, 50, 16, seed_off=3)
eng = build_engine(spec, engine=2)
eng.run(20)
for rep in range(8):
    eng.run(100, reset=False)
    tm = eng.getTimes()
    ms = tm['main_kernel_ms'] / tm['main_launches']
    st = eng.rungStates(0)
    X = spec["constants"]["X"][:2000]
    eta = X @ st.T
    print("iters %4d: ms/pass %.4f launches %d  Temps[-3:] %s scale[0,-1] %.3g %.3g  frac|eta|>8 per rung: %s" % (20 + 100 * (rep + 1), ms, tm['main_launches'], np.round(eng.Temps[-3:], 1), eng.scales(0)[0, 0], eng.scales(0)[-1, 0], np.round((np.abs(eta) > 8).mean(0)[[0, 8, 15]], 3)))
