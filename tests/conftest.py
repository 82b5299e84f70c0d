import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` under gpurun)")
    config.addinivalue_line("markers", "slow: full-size oracle parity (the CPU oracle takes up to a minute per case); part of `-m gpu`")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the front library and the oracle are built (cheap no-ops when they are)."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "nimbleapt_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
