"""ncu target for config 3's grid_pass_kernel: `python tools/prof_c3.py [iterations]` (default 10: cold chains; 1700 reaches
the adapted ladder, capture with `ncu -k regex:grid_run -s 6 -c 1`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nimbleapt_b200 import workloads as w
eng = w.build_engine(w.c3_logistic(), engine=2)
eng.run(int(sys.argv[1]) if len(sys.argv) > 1 else 10)
print(eng.getTimes())
