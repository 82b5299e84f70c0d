import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, models
from nimbleapt_b200.workloads import build_engine
eng = build_engine(models.c4_glmm(), nLadders=148)
eng.run(3)
print(eng.getTimes())
