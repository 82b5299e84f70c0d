"""Config 3 (logistic N = 1e6, p = 50, 16 rungs): ms per pass early and in the steady state, at 16 and 8 rungs per pass.
APTB200_DEBUG_TIMING=1 adds the cycle breakdown of the finishing CTA (stderr)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nimbleapt_b200 import workloads as w
from nimbleapt_b200.workloads import build_engine
spec = w.c3_logistic()
for tb in [int(x) for x in (sys.argv[1:] or ["16", "8"])]:
    eng = build_engine(spec, engine=2, rungsPerPass=tb)
    eng.run(20)
    eng.run(400, reset=False)
    t = eng.getTimes(); early = t["main_kernel_ms"] / t["main_launches"]
    eng.run(1200, reset=False)
    eng.run(400, reset=False)
    t = eng.getTimes(); steady = t["main_kernel_ms"] / t["main_launches"]
    by = 8.0 * 1e6 * 51
    print("rungs/pass %2d: early %.4f ms (%.3f of 6551.7 GB/s)  steady %.4f ms (%.3f)  updates/s steady %.0f" % (
        tb, early, by / early / 1e6 / 6551.7, steady, by / steady / 1e6 / 6551.7, 16 * 400 / (t["kernel_ms"] * 1e-3)), flush=True)
    del eng
