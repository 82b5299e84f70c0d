import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, models
from nimbleapt_b200.workloads import build_engine
eng = build_engine(models.c5_small_logistic(), nLadders=2048)
eng.run(40, thin=10)
print(eng.getTimes())
