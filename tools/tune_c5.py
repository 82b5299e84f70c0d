import sys, os, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, models
from nimbleapt_b200.workloads import build_engine
L = int(os.environ.get("C5_L", "2048"))
kw = {}
if os.environ.get("C5_W"): kw["lanesPerChain"] = int(os.environ["C5_W"])
if os.environ.get("C5_CTA"): kw["ctaThreads"] = int(os.environ["C5_CTA"])
t0 = time.time()
eng = build_engine(models.c5_small_logistic(), nLadders=L, **kw)
tb = time.time() - t0
eng.run(40, thin=10)
eng.run(200, reset=False, thin=10)
t = eng.getTimes()
print("MINB=%s W=%s CTA=%s L=%d: build %.0fs kernel %.1f ms -> %.3e updates/s (W=%d cta=%d)" % (os.environ.get("APTB200_LADDER_MINB"), os.environ.get("C5_W"), os.environ.get("C5_CTA"), L, tb,
      t["kernel_ms"], 200.0 * L * 64 / (t["kernel_ms"] * 1e-3), eng.info(13), eng.info(14)))
