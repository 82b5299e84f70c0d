"""CPU: the engine's softplus primitive (nimbleapt_b200/csrc/engine/apt_softplus.cuh), compiled for the host with
the identical arithmetic (fma for fma), against mpmath: log(1 + exp(-a)) on [0, 8] to < 2.3e-16 absolute."""
import ctypes as C
import os
import subprocess
import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_softplus_against_mpmath(tmp_path):
    src = tmp_path / "sp.c"
    src.write_text('#include "apt_softplus.cuh"\nvoid spv(const double* a, double* o, int n) { for (int i = 0; i < n; i++) o[i] = apt_softplus_neg(a[i]); }\n')
    so = tmp_path / "sp.so"
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c", "-I",
                           os.path.join(ROOT, "nimbleapt_b200", "csrc", "engine"), str(src), "-o", str(so), "-lm"])
    L = C.CDLL(str(so))
    L.spv.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    rng = np.random.default_rng(0)
    a = np.concatenate([np.linspace(0, 8, 4001), rng.random(6000) * 8, [0.0, 8.0, 1e-300, 1e-17, np.nextafter(8.0, 0)]])
    o = np.empty_like(a)
    L.spv(a.ctypes.data_as(C.POINTER(C.c_double)), o.ctypes.data_as(C.POINTER(C.c_double)), len(a))
    mp.mp.dps = 40
    err = max(abs(float(mp.mpf(float(o[i])) - mp.log1p(mp.exp(-mp.mpf(float(a[i])))))) for i in range(len(a)))
    assert err < 2.3e-16, err
    assert np.all(np.diff(o[:4001]) <= 1e-16)  # monotone non-increasing on the grid up to rounding
