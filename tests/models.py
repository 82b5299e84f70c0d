"""The workload definitions live in the package (nimbleapt_b200/workloads.py); tests keep importing `models`."""
from nimbleapt_b200.workloads import *  # noqa: F401,F403
