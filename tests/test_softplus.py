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


def test_exp_and_log_building_blocks_against_mpmath(tmp_path):
    """apt_exp / apt_exp_neg / apt_log_pos (same arithmetic as on the device): relative error below 2 ulp."""
    src = tmp_path / "el.c"
    src.write_text('#include "apt_softplus.cuh"\n'
                   'void ev(const double* a, double* o, int n) { for (int i = 0; i < n; i++) o[i] = apt_exp(a[i]); }\n'
                   'void en(const double* a, double* o, int n) { for (int i = 0; i < n; i++) o[i] = apt_exp_neg(a[i]); }\n'
                   'void lg(const double* a, double* o, int n) { for (int i = 0; i < n; i++) o[i] = apt_log_pos(a[i]); }\n')
    so = tmp_path / "el.so"
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c", "-I",
                           os.path.join(ROOT, "nimbleapt_b200", "csrc", "engine"), str(src), "-o", str(so), "-lm"])
    L = C.CDLL(str(so))
    P = C.POINTER(C.c_double)
    rng = np.random.default_rng(1)
    mp.mp.dps = 40

    def run(fn, a):
        o = np.empty_like(a)
        fn(a.ctypes.data_as(P), o.ctypes.data_as(P), len(a))
        return o

    x = np.concatenate([rng.uniform(-700, 700, 3000), rng.uniform(-2, 2, 2000), [0.0, -700.0, 700.0, 1e-300, -1e-300, 705.0, -705.0, 745.0]])
    o = run(L.ev, x)
    rel = max(abs(float(mp.mpf(float(o[i])) / mp.exp(mp.mpf(float(x[i]))) - 1)) for i in range(len(x)) if abs(x[i]) <= 700)
    assert rel < 4.5e-16, rel
    assert o[-3] == np.exp(705.0) and o[-2] == np.exp(-705.0)  # beyond the fast range: the library path
    a = np.concatenate([rng.uniform(0, 700, 3000), [0.0, 700.0, 746.5, 800.0]])
    o = run(L.en, a)
    rel = max(abs(float(mp.mpf(float(o[i])) / mp.exp(-mp.mpf(float(a[i]))) - 1)) for i in range(len(a)) if a[i] <= 700)
    assert rel < 4.5e-16, rel
    assert o[-1] == 0.0 and o[-2] == 0.0
    v = np.concatenate([np.exp(rng.uniform(-700, 700, 4000)), rng.uniform(0.5, 2.0, 2000), [1.0, 2.0, 0.5, 1e-300, 1e300]])
    o = run(L.lg, v)
    # table + polynomial: relative accuracy away from 1, absolute accuracy ~1e-16 next to 1 (the rounded 1/c_i of the
    # table).  The engine takes logs of uniforms, of swap probabilities and of x / lambda with |x - lambda| >= 0.1 (x + lambda).
    lv = [float(mp.log(mp.mpf(float(t)))) for t in v]
    ae = [abs(float(mp.mpf(float(o[i])) - mp.log(mp.mpf(float(v[i]))))) for i in range(len(v))]
    assert max(ae[i] for i in range(len(v)) if abs(lv[i]) < 0.7) < 1.7e-16, max(ae[i] for i in range(len(v)) if abs(lv[i]) < 0.7)
    assert max(ae[i] / abs(lv[i]) for i in range(len(v)) if abs(lv[i]) >= 0.35) < 4.5e-16
    assert o[len(v) - 5] == 0.0  # log(1) exactly
