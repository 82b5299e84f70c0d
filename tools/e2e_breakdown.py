import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import bench
from nimbleapt_b200.workloads import build_engine
spec, L, thin = bench.workload_spec("c2")
eng = build_engine(spec, nLadders=L)
y = np.ascontiguousarray(spec["data"]["y"])
th0 = np.concatenate([np.ravel(spec["inits"][k], order="F") for k in spec["inits"]])
steps = 4000
for rep in range(3):
    t = [time.perf_counter()]
    eng.setData("y", y); t.append(time.perf_counter())
    eng.setInits(th0); t.append(time.perf_counter())
    eng.run(steps, thin=thin); t.append(time.perf_counter())
    sm = eng.mvSamples.matrix(-1); t.append(time.perf_counter())
    tm = eng.getTimes()
    print("rep", rep, "setData %.3f setInits %.3f run %.3f (kernel %.3f) fetch %.3f ms  total %.3f" % tuple(
        [1e3 * (t[i + 1] - t[i]) for i in range(2)] + [1e3 * (t[3] - t[2]), tm["kernel_ms"], 1e3 * (t[4] - t[3]), 1e3 * (t[4] - t[0])]), sm.shape)
