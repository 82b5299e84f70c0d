import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, models
from nimbleapt_b200.workloads import build_engine
eng = build_engine(models.c2_mixture(), nLadders=4096)
eng.run(200, thin=10)
print(eng.getTimes())
