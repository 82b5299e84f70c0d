"""The reference's vignette (nimbleAPT/vignettes/nimbleAPT-vignette.Rmd:56-141) through the drop-in interface:
a bimodal toy posterior (abs() of the centroids), two sampler_RW_tempered, buildAPT, run, summarise.

    python examples/vignette.py            (needs a B200; compiles the lowered model with nvcc on first use)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nimbleapt_b200 as nb

bugs = """
{
  for (ii in 1:nObs) {
    y[ii, 1:2] ~ dmnorm(mean = absCentroids[1:2], cholesky = cholCov[1:2, 1:2], prec_param = 0)
  }
  absCentroids[1:2] <- abs(centroids[1:2])
  for (ii in 1:2) {
    centroids[ii] ~ dnorm(0, sd = 1E3)
  }
}
"""
rng = np.random.default_rng(11)
nObs = 100
y = rng.standard_normal((nObs, 2)) + 3.0           # simulated from centroids = (-3, -3): |centroids| = (3, 3)

model = nb.nimbleModel(nb.nimbleCode(bugs), constants=dict(nObs=nObs, cholCov=np.eye(2)), data=dict(y=y),
                       inits=dict(centroids=[-3.0, -3.0]))
conf = nb.configureMCMC(model, nodes="centroids", monitors="centroids")
conf.removeSamplers()
conf.addSampler("centroids[1]", type="sampler_RW_tempered", control=dict(temperPriors=True))
conf.addSampler("centroids[2]", type="sampler_RW_tempered", control=dict(temperPriors=True))
aptC = nb.compileNimble(nb.buildAPT(conf, Temps=range(1, 8), ULT=1000))
aptC.run(niter=15000)
samples = nb.as_matrix(aptC.mvSamples)[-10000:]
print("final ladder      :", np.round(aptC.Temps, 2))
print("mean |centroids|  :", np.abs(samples).mean(0), " (ybar =", y.mean(0), ")")
for sx in (+1, -1):
    for sy in (+1, -1):
        print("mass of mode (%+d,%+d): %.3f   (1/4 each in the exact posterior)" % (sx, sy, np.mean((np.sign(samples[:, 0]) == sx) & (np.sign(samples[:, 1]) == sy))))
