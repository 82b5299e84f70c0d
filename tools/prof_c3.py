import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, models
from nimbleapt_b200.workloads import build_engine
spec = models.logistic(1000000, 50, 16, seed_off=3)
eng = build_engine(spec, engine=2)
eng.run(10)
print(eng.getTimes())
